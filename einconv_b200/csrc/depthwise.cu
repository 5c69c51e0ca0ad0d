// K7-depthwise: direct stencil kernels for groups == in_channels (c_in_g == 1), 1-d and 2-d.
//
// This is what the reference's simplifier reaches by squeezing c_in / c_out out of the network
// (simplifications/opt.py:148-171; observed equation `ngbc,deb,fhc,gdf->ngeh`):
// each output channel reads a single input channel, so there is no reduction over channels to feed
// a tensor core with — the op is an HBM-bound stencil (2.25 flop/B for 3x3).  Forward
// (expressions/convNd_forward.py:48-55) and input VJP (expressions/convNd_input_vjp.py:53-60).
//
// A CTA owns one (b, c) input plane and a tile of output rows.  It stages the zero-padded input
// window of the tile in shared memory once and produces the tile for all `c_out_g` output channels
// of that group; taps are unrolled at compile time for the common kernel sizes, weights live in
// registers.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace einconv {

struct DwParams {
  int batch, channels, mult;   // B, C (= groups), c_out_g
  int in_h, in_w, out_h, out_w;
  int k_h, k_w, s_h, s_w, d_h, d_w, p_h, p_w;
  int tile_h, n_tiles;         // rows of the produced tensor per CTA
  int rows_cap, cols;          // staged window
  FastDiv w_div;               // / width of the produced tensor
};

// ---- forward: y[b, c*m + j, oh, ow] = bias + sum_{kh,kw} x[b, c, s*oh + d*kh - p, ..] * w[c*m + j, kh, kw]
template <typename T, int KH, int KW>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const T* __restrict__ x,
                                                     const T* __restrict__ w,
                                                     const T* __restrict__ bias, T* __restrict__ y,
                                                     const __grid_constant__ DwParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* win = reinterpret_cast<float*>(smem_raw);  // [rows][cols], zero padded
  const int k_h = KH > 0 ? KH : p.k_h, k_w = KW > 0 ? KW : p.k_w;

  const int tile = blockIdx.x % p.n_tiles;
  const int bc = blockIdx.x / p.n_tiles;
  const int b = bc / p.channels, c = bc - b * p.channels;
  const int oh0 = tile * p.tile_h;
  const int th = min(p.tile_h, p.out_h - oh0);
  const int rows = p.s_h * (th - 1) + p.d_h * (k_h - 1) + 1;
  const int cols = p.cols;

  const T* src = x + (long long)bc * p.in_h * p.in_w;
  const int r0 = p.s_h * oh0 - p.p_h;
  for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
    const int jr = idx / cols, jc = idx - jr * cols;
    const int r = r0 + jr, cc = jc - p.p_w;
    float val = 0.f;
    if (r >= 0 && r < p.in_h && cc >= 0 && cc < p.in_w) val = (float)Cvt<T>::load(src + r * p.in_w + cc);
    win[idx] = val;
  }
  __syncthreads();

  const int chunk = th * p.out_w;
  for (int j = 0; j < p.mult; ++j) {
    const int co = c * p.mult + j;
    const T* wc = w + (long long)co * k_h * k_w;
    float wreg[(KH > 0 ? KH * KW : 1)];
    if (KH > 0) {
#pragma unroll
      for (int t = 0; t < KH * KW; ++t) wreg[t] = (float)Cvt<T>::load(wc + t);
    }
    const float bv = bias ? (float)Cvt<T>::load(bias + co) : 0.f;
    T* dst = y + ((long long)b * p.channels * p.mult + co) * p.out_h * p.out_w + (long long)oh0 * p.out_w;
    for (int e = threadIdx.x; e < chunk; e += blockDim.x) {
      uint32_t ohl, ow;
      p.w_div.divmod((uint32_t)e, ohl, ow);
      const float* base = win + (p.s_h * (int)ohl) * cols + p.s_w * (int)ow;
      float acc = 0.f;
      if (KH > 0) {
#pragma unroll
        for (int kh = 0; kh < KH; ++kh)
#pragma unroll
          for (int kw = 0; kw < KW; ++kw)
            acc = fmaf(base[p.d_h * kh * cols + p.d_w * kw], wreg[kh * KW + kw], acc);
      } else {
        for (int kh = 0; kh < k_h; ++kh)
          for (int kw = 0; kw < k_w; ++kw)
            acc = fmaf(base[p.d_h * kh * cols + p.d_w * kw],
                       (float)Cvt<T>::load(wc + kh * k_w + kw), acc);
      }
      Cvt<T>::store(dst + e, (typename Cvt<T>::acc_t)(acc + bv));
    }
  }
}

// ---- input VJP: gx[b, c, ih, iw] = sum_j sum_{kh,kw : s | (i + p - d*k)} v[b, c*m + j, (ih+p-d*kh)/s, ..] * w[c*m + j, kh, kw]
// The CTA stages the cotangent rows its input-row tile can touch, for one output channel at a time.
template <typename T, int KH, int KW>
__global__ void __launch_bounds__(256) dw_input_vjp_kernel(const T* __restrict__ w,
                                                           const T* __restrict__ v,
                                                           T* __restrict__ gx,
                                                           const __grid_constant__ DwParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* win = reinterpret_cast<float*>(smem_raw);  // [rows][out_w + 2*halo], zero padded
  const int k_h = KH > 0 ? KH : p.k_h, k_w = KW > 0 ? KW : p.k_w;

  const int tile = blockIdx.x % p.n_tiles;
  const int bc = blockIdx.x / p.n_tiles;
  const int b = bc / p.channels, c = bc - b * p.channels;
  const int ih0 = tile * p.tile_h;
  const int th = min(p.tile_h, p.in_h - ih0);
  // cotangent rows that can be touched: oh in [floor((ih0 + p - d(K-1)) / s), (ih0 + th - 1 + p) / s]
  const int lo_num = ih0 + p.p_h - p.d_h * (k_h - 1);
  const int oh_lo = lo_num >= 0 ? lo_num / p.s_h : -((-lo_num + p.s_h - 1) / p.s_h);
  const int oh_hi = (ih0 + th - 1 + p.p_h) / p.s_h;
  const int rows = oh_hi - oh_lo + 1;
  // cotangent columns: ow in [ow_lo, ow_hi] for the full input width
  const int wl_num = p.p_w - p.d_w * (k_w - 1);
  const int ow_lo = wl_num >= 0 ? wl_num / p.s_w : -((-wl_num + p.s_w - 1) / p.s_w);
  const int cols = p.cols;

  const int chunk = th * p.in_w;

  // accumulate over the group's output channels; partial sums stay in registers across j by
  // looping e in the outer position would need storage, so loop j outermost with a smem pass each.
  T* dst = gx + (long long)bc * p.in_h * p.in_w + (long long)ih0 * p.in_w;
  for (int j = 0; j < p.mult; ++j) {
    const int co = c * p.mult + j;
    const T* srcv = v + ((long long)b * p.channels * p.mult + co) * p.out_h * p.out_w;
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
      const int jr = idx / cols, jc = idx - jr * cols;
      const int oh = oh_lo + jr, ow = ow_lo + jc;
      float val = 0.f;
      if (oh >= 0 && oh < p.out_h && ow >= 0 && ow < p.out_w)
        val = (float)Cvt<T>::load(srcv + oh * p.out_w + ow);
      win[idx] = val;
    }
    __syncthreads();
    const T* wc = w + (long long)co * k_h * k_w;
    float wreg[(KH > 0 ? KH * KW : 1)];
    if (KH > 0) {
#pragma unroll
      for (int t = 0; t < KH * KW; ++t) wreg[t] = (float)Cvt<T>::load(wc + t);
    }
    for (int e = threadIdx.x; e < chunk; e += blockDim.x) {
      uint32_t ihl, iw;
      p.w_div.divmod((uint32_t)e, ihl, iw);
      const int ih = ih0 + (int)ihl;
      float acc = 0.f;
      auto tap = [&](int kh, int kw, float wv) {
        const int th_ = ih + p.p_h - p.d_h * kh;
        const int tw_ = (int)iw + p.p_w - p.d_w * kw;
        // floor division is not needed: negative numerators are rejected
        const int oh = th_ / p.s_h, ow = tw_ / p.s_w;
        if (th_ >= 0 && tw_ >= 0 && oh * p.s_h == th_ && ow * p.s_w == tw_)
          acc = fmaf(win[(oh - oh_lo) * cols + (ow - ow_lo)], wv, acc);
      };
      if (KH > 0) {
#pragma unroll
        for (int kh = 0; kh < KH; ++kh)
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) tap(kh, kw, wreg[kh * KW + kw]);
      } else {
        for (int kh = 0; kh < k_h; ++kh)
          for (int kw = 0; kw < k_w; ++kw) tap(kh, kw, (float)Cvt<T>::load(wc + kh * k_w + kw));
      }
      if (j > 0) acc += (float)Cvt<T>::load(dst + e);  // running sum over the group's channels
      Cvt<T>::store(dst + e, (typename Cvt<T>::acc_t)acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path: register-tiled direct stencil, dilation 1, compile-time taps and stride.
//
//   out[b, q, oh, ow] = bias[q] + sum_{j<J} sum_{kh,kw} in[b, src(q,j), S*oh + kh - p_h, S*ow + kw - p_w] * w[wsrc(q,j), t(kh,kw)]
//
// forward:          q = output channel, J = 1, src = q / mult, wsrc = q, t = kh*KW + kw
// input VJP (S=1):  q = input channel,  J = mult, src = wsrc = q*mult + j, t flipped, p' = K-1-p
//                   (the adjoint of a stride-1 correlation is the correlation with the flipped kernel)
//
// A thread produces RV vertically adjacent outputs at one column, so consecutive lanes read
// consecutive addresses (coalesced, L1-resident re-reads) and each loaded row is reused by up to
// KH outputs from registers.  No shared memory, no barrier; x is read once from HBM and y written
// once: the kernel is HBM-bound (2.25 flop/B for 3x3).
// ---------------------------------------------------------------------------------------------
struct DwFastParams {
  int planes_out;          // B * (number of produced channels)
  int chan_out;            // produced channels per sample
  int mult;                // c_out_g
  int J;                   // planes summed per output plane
  int in_h, in_w, out_h, out_w, p_h, p_w;
  int n_vgroups;           // ceil(out_h / RV)
  int flip;                // 1: use taps in reversed order (input VJP)
  int in_chan;             // channels per sample of the read tensor
  FastDiv ow_div, vg_div, chan_div, mult_div;
  long long total;         // planes_out * n_vgroups * out_w threads
};

template <typename T, int KH, int KW, int S, int RV, bool FLIP>
__global__ void __launch_bounds__(256) dw_direct_kernel(const T* __restrict__ in,
                                                        const T* __restrict__ w,
                                                        const T* __restrict__ bias,
                                                        T* __restrict__ out,
                                                        const __grid_constant__ DwFastParams p) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (unsigned)p.total) return;
  uint32_t rest, ow, plane, vg, b, q;
  p.ow_div.divmod(tid, rest, ow);
  p.vg_div.divmod(rest, plane, vg);
  p.chan_div.divmod(plane, b, q);
  const int oh0 = (int)vg * RV;
  constexpr int ROWS = S * (RV - 1) + KH;
  const int in_w = p.in_w, in_h = p.in_h;
  const int r0 = S * oh0 - p.p_h;
  const int c0 = S * (int)ow - p.p_w;
  const bool interior = r0 >= 0 && r0 + ROWS <= in_h && c0 >= 0 && c0 + KW <= in_w;
  const int plane_elems = in_h * in_w;
  const int off0 = r0 * in_w + c0;  // may be negative; only dereferenced where in bounds

  float acc[RV];
#pragma unroll
  for (int r = 0; r < RV; ++r) acc[r] = 0.f;

  const int J = FLIP ? p.mult : 1;
  for (int j = 0; j < J; ++j) {
    int src;
    if (FLIP) {
      src = (int)q * p.mult + j;
    } else {
      src = (int)p.mult_div.quot(q);
    }
    const int wsrc = FLIP ? src : (int)q;
    float wreg[KH * KW];
#pragma unroll
    for (int t = 0; t < KH * KW; ++t)
      wreg[t] = (float)Cvt<T>::load(w + wsrc * (KH * KW) + (FLIP ? KH * KW - 1 - t : t));
    const T* base = in + ((long long)b * p.in_chan + src) * plane_elems + off0;

    auto accumulate = [&](int rr, const float* val) {
      // input row rr feeds output (rr - kh) / S for every kh with S | (rr - kh)
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
        if ((rr - kh) >= 0 && (rr - kh) % S == 0 && (rr - kh) / S < RV) {
#pragma unroll
          for (int kw = 0; kw < KW; ++kw)
            acc[(rr - kh) / S] = fmaf(val[kw], wreg[kh * KW + kw], acc[(rr - kh) / S]);
        }
      }
    };

    if (interior) {
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        float val[KW];
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) val[kw] = (float)Cvt<T>::load(base + rr * in_w + kw);
        accumulate(rr, val);
      }
    } else {
      bool col_ok[KW];
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) col_ok[kw] = (c0 + kw >= 0) && (c0 + kw < in_w);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const bool row_ok = (r0 + rr >= 0) && (r0 + rr < in_h);
        float val[KW];
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
          val[kw] = 0.f;
          if (row_ok && col_ok[kw]) val[kw] = (float)Cvt<T>::load(base + rr * in_w + kw);
        }
        accumulate(rr, val);
      }
    }
  }
  const float bv = bias ? (float)Cvt<T>::load(bias + q) : 0.f;
  const int out_w = p.out_w;
  T* dst = out + (long long)plane * p.out_h * out_w + oh0 * out_w + (int)ow;
#pragma unroll
  for (int r = 0; r < RV; ++r)
    if (oh0 + r < p.out_h) Cvt<T>::store(dst + r * out_w, (typename Cvt<T>::acc_t)(acc[r] + bv));
}

// ---------------------------------------------------------------------------------------------
// Warp-per-strip stencil for stride 1, channel multiplier 1 (the MobileNet case).  Lane l holds input
// column c_in0 + l of the current input row (ONE coalesced load per row and warp); the KW - 1
// neighbours come from shuffles, the KH rows of the window live in registers and slide down the
// plane.  Per output row: 1 load + (KW - 1) shuffles + KH * KW FMAs + 1 coalesced store per warp,
// versus KH * KW / 4 scalar loads per output of the thread-per-column kernel above (which was
// instruction-issue bound: 74% issue-active, 12% of DRAM).
// ---------------------------------------------------------------------------------------------
constexpr int kDwWarpRows = 16;  // input rows held in registers per warp

struct DwWarpParams {
  int planes, chan;        // B * C, C
  int in_h, in_w, out_h, out_w, p_h, p_w;
  int n_strips, strip_out; // output columns per strip = 33 - KW
  int n_rblocks, rows_per_block;
  long long total_warps;
  FastDiv strip_div, rblock_div, chan_div;
};

template <typename T, int KH, int KW, bool FLIP>
__global__ void __launch_bounds__(256) dw_warp_kernel(const T* __restrict__ in, const T* __restrict__ w,
                                                      const T* __restrict__ bias, T* __restrict__ out,
                                                      const __grid_constant__ DwWarpParams p) {
  const int lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wg >= p.total_warps) return;
  uint32_t rest, strip, plane, rblk, b, q;
  p.strip_div.divmod((uint32_t)wg, rest, strip);
  p.rblock_div.divmod(rest, plane, rblk);
  p.chan_div.divmod(plane, b, q);
  float wreg[KH * KW];
#pragma unroll
  for (int t = 0; t < KH * KW; ++t)
    wreg[t] = (float)Cvt<T>::load(w + (int)q * (KH * KW) + (FLIP ? KH * KW - 1 - t : t));
  const float bv = bias ? (float)Cvt<T>::load(bias + q) : 0.f;

  const int oh0 = (int)rblk * p.rows_per_block;
  const int oh1 = min(p.out_h, oh0 + p.rows_per_block);
  const int ow = (int)strip * p.strip_out + lane;
  const int ic = ow - p.p_w;
  const bool col_ok = ic >= 0 && ic < p.in_w;
  const T* base = in + (long long)plane * p.in_h * p.in_w + ic;
  T* dst = out + (long long)plane * p.out_h * p.out_w + ow;
  const bool store_ok = lane < p.strip_out && ow < p.out_w;

  // all NR input rows of this warp's block are requested before the first use (NR independent loads
  // in flight per lane: with one load per loop iteration the kernel was latency-bound at 25 us)
  constexpr int NR = kDwWarpRows;
  const int ih0 = oh0 - p.p_h;
  float v[NR][KW];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int ih = ih0 + i;
    v[i][0] = 0.f;
    if (col_ok && ih >= 0 && ih < p.in_h && i < (oh1 - oh0) + KH - 1)
      v[i][0] = (float)Cvt<T>::load(base + (long long)ih * p.in_w);
  }
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int kw = 1; kw < KW; ++kw) v[i][kw] = __shfl_down_sync(0xffffffffu, v[i][0], kw);
#pragma unroll
  for (int o = 0; o < NR - KH + 1; ++o) {
    if (oh0 + o < oh1) {
      float acc = bv;
#pragma unroll
      for (int kh = 0; kh < KH; ++kh)
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) acc = fmaf(v[o + kh][kw], wreg[kh * KW + kw], acc);
      if (store_ok) Cvt<T>::store(dst + (long long)(oh0 + o) * p.out_w, (typename Cvt<T>::acc_t)acc);
    }
  }
}

template <typename T, int KH, int KW>
static int launch_dw_warp(const void* in, const void* w, const void* bias, void* out, const DwFastParams& f,
                          cudaStream_t stream) {
  DwWarpParams p{};
  p.planes = f.planes_out;
  p.chan = f.chan_out;
  p.in_h = f.in_h; p.in_w = f.in_w; p.out_h = f.out_h; p.out_w = f.out_w; p.p_h = f.p_h; p.p_w = f.p_w;
  p.strip_out = 33 - KW;
  p.n_strips = ceil_div(p.out_w, p.strip_out);
  // a warp produces up to kDwWarpRows - KH + 1 output rows from kDwWarpRows input rows
  const long long base_warps = (long long)p.planes * p.n_strips;
  const int max_rows = kDwWarpRows - KH + 1;
  p.rows_per_block = ceil_div(p.out_h, ceil_div(p.out_h, max_rows));  // even split, <= max_rows
  p.n_rblocks = ceil_div(p.out_h, p.rows_per_block);
  p.total_warps = base_warps * p.n_rblocks;
  EINCONV_CHECK_ARG(p.total_warps < (1ll << 31), "depthwise: tensor too large");
  p.strip_div = FastDiv((uint32_t)p.n_strips);
  p.rblock_div = FastDiv((uint32_t)p.n_rblocks);
  p.chan_div = FastDiv((uint32_t)p.chan);
  const unsigned grid = (unsigned)((p.total_warps + 7) / 8);
  if (f.flip)
    dw_warp_kernel<T, KH, KW, true><<<grid, 256, 0, stream>>>((const T*)in, (const T*)w, (const T*)bias, (T*)out, p);
  else
    dw_warp_kernel<T, KH, KW, false><<<grid, 256, 0, stream>>>((const T*)in, (const T*)w, (const T*)bias, (T*)out, p);
  EINCONV_CHECK_LAUNCH("dw_warp_kernel");
  return EINCONV_OK;
}

// ---------------------------------------------------------------------------------------------
// Vectorised "same" stencil for fp32 rows that are a multiple of 4 wide (stride 1, 2 * p_w = KW - 1,
// channel multiplier 1: the MobileNet case).  A lane owns FOUR adjacent columns: one 16-byte load per
// input row, the (KW - 1) halo columns come from the two neighbouring lanes, KH rolling accumulator
// rows, one 16-byte store per output row.  W / 4 lanes cover a row, so a warp works on 32 / (W / 4)
// independent (plane, row block) units side by side.  ~11 instructions per output instead of ~36 in
// dw_warp_kernel, which was issue-bound (70 % issue-active at 0.4 of the HBM roofline).
// ---------------------------------------------------------------------------------------------
struct DwVecParams {
  int planes, chan;
  int in_h, in_w, out_h, p_h;      // out_w == in_w
  int lanes_per_row, slots;        // in_w / 4, 32 / lanes_per_row
  int n_rblocks, rows_per_block;
  long long total_units;           // planes * n_rblocks
  FastDiv rblock_div, chan_div, lane_div;
};

__device__ __forceinline__ float f4_get(const float4& v, int c) {
  return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

template <int KH, int KW, bool FLIP, int NR>
__global__ void __launch_bounds__(128) dw_vec4_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                      const __grid_constant__ DwVecParams p) {
  constexpr int PW = (KW - 1) / 2;  // NR: input rows held in registers per unit
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint32_t slot, cg;
  p.lane_div.divmod((uint32_t)lane, slot, cg);
  const long long unit = warp * p.slots + slot;
  const bool active = (int)slot < p.slots && unit < p.total_units;
  uint32_t plane = 0, rblk = 0, b = 0, q = 0;
  if (active) {
    p.rblock_div.divmod((uint32_t)unit, plane, rblk);
    p.chan_div.divmod(plane, b, q);
  }
  float wreg[KH * KW];
#pragma unroll
  for (int t = 0; t < KH * KW; ++t) wreg[t] = __ldg(w + (int)q * (KH * KW) + (FLIP ? KH * KW - 1 - t : t));
  const float bv = bias ? __ldg(bias + q) : 0.f;

  const int row_v = p.lanes_per_row;
  const int oh0 = (int)rblk * p.rows_per_block;
  const int rows = min(p.rows_per_block, p.out_h - oh0);
  const int ih0 = oh0 - p.p_h;
  const float4* base = reinterpret_cast<const float4*>(in + (long long)plane * p.in_h * p.in_w) + cg;
  float4* dst = reinterpret_cast<float4*>(out + (long long)plane * p.out_h * p.in_w) + cg;
  const bool first = cg == 0, last = (int)cg == row_v - 1;

  float4 v[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int ih = ih0 + i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && i < rows + KH - 1 && ih >= 0 && ih < p.in_h) v[i] = __ldg(base + (long long)ih * row_v);
  }
  float acc[KH][4];
#pragma unroll
  for (int r = 0; r < KH; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = bv;
  const int n_in = p.rows_per_block + KH - 1;  // uniform over the grid: shuffles stay convergent
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (i < n_in) {
      float e[4 + KW - 1];
#pragma unroll
      for (int c = 0; c < 4; ++c) e[PW + c] = f4_get(v[i], c);
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const float t = __shfl_up_sync(FULL, f4_get(v[i], 4 - PW + j), 1);
        e[j] = first ? 0.f : t;
      }
#pragma unroll
      for (int j = 0; j < KW - 1 - PW; ++j) {
        const float t = __shfl_down_sync(FULL, f4_get(v[i], j), 1);
        e[PW + 4 + j] = last ? 0.f : t;
      }
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
        const int o = i - kh;
        if (o >= 0 && o < NR - KH + 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int kw = 0; kw < KW; ++kw)
              acc[o % KH][c] = fmaf(e[c + kw], wreg[kh * KW + kw], acc[o % KH][c]);
        }
      }
      const int oc = i - KH + 1;
      if (oc >= 0) {
        if (active && oc < rows)
          dst[(long long)(oh0 + oc) * row_v] =
              make_float4(acc[oc % KH][0], acc[oc % KH][1], acc[oc % KH][2], acc[oc % KH][3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[oc % KH][c] = bv;
      }
    }
  }
}

template <int KH, int KW>
static int launch_dw_vec4(const void* in, const void* w, const void* bias, void* out, const DwFastParams& f,
                          cudaStream_t stream) {
  DwVecParams p{};
  p.planes = f.planes_out;
  p.chan = f.chan_out;
  p.in_h = f.in_h; p.in_w = f.in_w; p.out_h = f.out_h; p.p_h = f.p_h;
  p.lanes_per_row = f.in_w / 4;
  p.slots = 32 / p.lanes_per_row;
  const int max_rows = kDwWarpRows - KH + 1;
  p.n_rblocks = ceil_div(p.out_h, max_rows);
  // enough warps to fill the chip several times over: halve the row blocks while they stay >= 2 * KH
  auto warps_for = [&](int nrb) { return ((long long)p.planes * nrb + p.slots - 1) / p.slots; };
  while (warps_for(p.n_rblocks) < 148ll * 32 && ceil_div(p.out_h, p.n_rblocks * 2) >= 2 * KH) p.n_rblocks *= 2;
  // small problems (C3: 51 MB per direction) are latency-bound: blocks of 4 output rows use the variant
  // that holds KH + 3 input rows (56 registers, 8 CTAs per SM); measured 10.2 us against 10.8 (7-row blocks)
  // and 12.3 (14-row blocks).  Large ones keep the long blocks (less halo re-reading).
  if (p.out_h > 4 && warps_for(ceil_div(p.out_h, 4)) <= 148ll * 32 * 8) p.n_rblocks = ceil_div(p.out_h, 4);
  if (const char* e = getenv("EINCONV_DW_VEC_RBLOCKS")) {
    const int v = atoi(e);
    if (v >= p.n_rblocks && v <= p.out_h) p.n_rblocks = v;
  }
  p.rows_per_block = ceil_div(p.out_h, p.n_rblocks);
  p.n_rblocks = ceil_div(p.out_h, p.rows_per_block);
  p.total_units = (long long)p.planes * p.n_rblocks;
  EINCONV_CHECK_ARG(p.total_units < (1ll << 31), "depthwise: tensor too large");
  p.rblock_div = FastDiv((uint32_t)p.n_rblocks);
  p.chan_div = FastDiv((uint32_t)p.chan);
  p.lane_div = FastDiv((uint32_t)p.lanes_per_row);
  const long long warps = warps_for(p.n_rblocks);
  const unsigned grid = (unsigned)((warps + 3) / 4);
  // short row blocks: a variant that holds fewer input rows needs fewer registers (more resident warps)
  constexpr int NRS = KH + 7, NRT = KH + 3;
  const int n_in = p.rows_per_block + KH - 1;
  const bool big = getenv("EINCONV_DW_VEC_NR16") != nullptr;
  const bool tiny = !big && n_in <= NRT;
  const bool small = !big && !tiny && n_in <= NRS && NRS < kDwWarpRows;
  auto go = [&](auto kern) {
    kern<<<grid, 128, 0, stream>>>((const float*)in, (const float*)w, (const float*)bias, (float*)out, p);
  };
  if (f.flip) {
    if (tiny) go(dw_vec4_kernel<KH, KW, true, NRT>);
    else if (small) go(dw_vec4_kernel<KH, KW, true, NRS>);
    else go(dw_vec4_kernel<KH, KW, true, kDwWarpRows>);
  } else {
    if (tiny) go(dw_vec4_kernel<KH, KW, false, NRT>);
    else if (small) go(dw_vec4_kernel<KH, KW, false, NRS>);
    else go(dw_vec4_kernel<KH, KW, false, kDwWarpRows>);
  }
  EINCONV_CHECK_LAUNCH("dw_vec4_kernel");
  return EINCONV_OK;
}

template <typename T, int KH, int KW, int S>
static int launch_dw_direct(const void* in, const void* w, const void* bias, void* out,
                            DwFastParams& p, cudaStream_t stream) {
  constexpr int RV = 4;
  p.n_vgroups = ceil_div(p.out_h, RV);
  p.ow_div = FastDiv((uint32_t)p.out_w);
  p.vg_div = FastDiv((uint32_t)p.n_vgroups);
  p.chan_div = FastDiv((uint32_t)p.chan_out);
  p.mult_div = FastDiv((uint32_t)p.mult);
  p.total = (long long)p.planes_out * p.n_vgroups * p.out_w;
  EINCONV_CHECK_ARG(p.total < (1ll << 31) && (long long)p.in_h * p.in_w < (1ll << 30),
                    "depthwise: tensor too large");
  const unsigned grid = (unsigned)((p.total + 255) / 256);
  if (p.flip)
    dw_direct_kernel<T, KH, KW, S, RV, true><<<grid, 256, 0, stream>>>(
        (const T*)in, (const T*)w, (const T*)bias, (T*)out, p);
  else
    dw_direct_kernel<T, KH, KW, S, RV, false><<<grid, 256, 0, stream>>>(
        (const T*)in, (const T*)w, (const T*)bias, (T*)out, p);
  EINCONV_CHECK_LAUNCH("dw_direct_kernel");
  return EINCONV_OK;
}

// returns 1 if a fast variant was launched, 0 if none applies, <0 on error
template <typename T>
static int try_dw_direct(const void* in, const void* w, const void* bias, void* out,
                         DwFastParams& p, int k_h, int k_w, int s_h, int s_w, int d_h, int d_w,
                         cudaStream_t stream) {
  if (d_h != 1 || d_w != 1 || s_h != s_w) return 0;
  int st = 1;
  if constexpr (std::is_same_v<T, float>) {
    // "same" geometry along W, rows a multiple of 4 wide, 16-byte aligned tensors
    if (s_h == 1 && p.mult == 1 && k_h == k_w && 2 * p.p_w == k_w - 1 && p.out_w == p.in_w &&
        p.in_w % 4 == 0 && p.in_w <= 128 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0 &&
        !getenv("EINCONV_DW_DIRECT") && !getenv("EINCONV_DW_NO_VEC")) {
      if (k_h == 3) return launch_dw_vec4<3, 3>(in, w, bias, out, p, stream) == EINCONV_OK ? 1 : EINCONV_ERR_CUDA;
      if (k_h == 5) return launch_dw_vec4<5, 5>(in, w, bias, out, p, stream) == EINCONV_OK ? 1 : EINCONV_ERR_CUDA;
      if (k_h == 7) return launch_dw_vec4<7, 7>(in, w, bias, out, p, stream) == EINCONV_OK ? 1 : EINCONV_ERR_CUDA;
    }
  }
  if (s_h == 1 && p.mult == 1 && !getenv("EINCONV_DW_DIRECT")) {
    if (k_h == 3 && k_w == 3) st = launch_dw_warp<T, 3, 3>(in, w, bias, out, p, stream);
    else if (k_h == 5 && k_w == 5) st = launch_dw_warp<T, 5, 5>(in, w, bias, out, p, stream);
    else if (k_h == 7 && k_w == 7) st = launch_dw_warp<T, 7, 7>(in, w, bias, out, p, stream);
    else if (k_h == 1 && k_w == 3) st = launch_dw_warp<T, 1, 3>(in, w, bias, out, p, stream);
    else st = 1;
    if (st != 1) return st == EINCONV_OK ? 1 : st;
  }
  if (k_h == 3 && k_w == 3 && s_h == 1) st = launch_dw_direct<T, 3, 3, 1>(in, w, bias, out, p, stream);
  else if (k_h == 3 && k_w == 3 && s_h == 2) st = launch_dw_direct<T, 3, 3, 2>(in, w, bias, out, p, stream);
  else if (k_h == 5 && k_w == 5 && s_h == 1) st = launch_dw_direct<T, 5, 5, 1>(in, w, bias, out, p, stream);
  else if (k_h == 1 && k_w == 3 && s_h == 1) st = launch_dw_direct<T, 1, 3, 1>(in, w, bias, out, p, stream);
  else return 0;
  return st == EINCONV_OK ? 1 : st;
}

// ---------------------------------------------------------------------------------------------
bool depthwise_supported(const DevGeom& g, int dtype) {
  return g.c_in_g == 1 && g.groups > 1 && g.n_dims <= 2 && dtype != EINCONV_F64 && g.out_total > 0 &&
         g.in_total > 0;
}

static void fill_dw(DwParams& p, const DevGeom& g) {
  const int N = g.n_dims;
  auto get = [&](const int* a, int which, int dflt) {  // which: 0 = H, 1 = W
    const int d = N - 2 + which;
    return d >= 0 ? a[d] : dflt;
  };
  p.batch = g.batch; p.channels = g.groups; p.mult = g.c_out_g;
  p.in_h = get(g.in, 0, 1);   p.in_w = get(g.in, 1, 1);
  p.out_h = get(g.out, 0, 1); p.out_w = get(g.out, 1, 1);
  p.k_h = get(g.k, 0, 1);     p.k_w = get(g.k, 1, 1);
  p.s_h = get(g.stride, 0, 1); p.s_w = get(g.stride, 1, 1);
  p.d_h = get(g.dil, 0, 1);   p.d_w = get(g.dil, 1, 1);
  p.p_h = get(g.pad, 0, 0);   p.p_w = get(g.pad, 1, 0);
}

template <typename F>
static int dispatch_taps(int k_h, int k_w, F&& f) {
  if (k_h == 3 && k_w == 3) return f(std::integral_constant<int, 3>{}, std::integral_constant<int, 3>{});
  if (k_h == 5 && k_w == 5) return f(std::integral_constant<int, 5>{}, std::integral_constant<int, 5>{});
  if (k_h == 1 && k_w == 3) return f(std::integral_constant<int, 1>{}, std::integral_constant<int, 3>{});
  return f(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
}

int dw_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                const DevGeom& g, cudaStream_t stream) {
  DwParams p{};
  fill_dw(p, g);
  if (dtype != EINCONV_F64) {
    DwFastParams f{};
    f.planes_out = p.batch * p.channels * p.mult; f.chan_out = p.channels * p.mult;
    f.mult = p.mult; f.J = 1; f.flip = 0; f.in_chan = p.channels;
    f.in_h = p.in_h; f.in_w = p.in_w; f.out_h = p.out_h; f.out_w = p.out_w; f.p_h = p.p_h; f.p_w = p.p_w;
    int fast = dispatch_dtype(dtype, [&](auto* tag) -> int {
      using T = std::remove_pointer_t<decltype(tag)>;
      if constexpr (std::is_same_v<T, double>) return 0;
      else return try_dw_direct<T>(x, w, bias, y, f, p.k_h, p.k_w, p.s_h, p.s_w, p.d_h, p.d_w, stream);
    });
    if (fast != 0) return fast > 0 ? EINCONV_OK : fast;
  }
  p.cols = p.s_w * (p.out_w - 1) + p.d_w * (p.k_w - 1) + 1;
  auto rows_for = [&](int th) { return p.s_h * (th - 1) + p.d_h * (p.k_h - 1) + 1; };
  int th = p.out_h;
  while (th > 1 && (long long)rows_for(th) * p.cols * 4 > 32 * 1024) th = (th + 1) / 2;
  EINCONV_CHECK_ARG((long long)rows_for(th) * p.cols * 4 <= 200 * 1024, "depthwise: row too wide");
  p.tile_h = th; p.n_tiles = ceil_div(p.out_h, th); p.rows_cap = rows_for(th);
  p.w_div = FastDiv((uint32_t)p.out_w);
  const size_t smem = (size_t)p.rows_cap * p.cols * 4;
  const long long grid = (long long)p.batch * p.channels * p.n_tiles;
  EINCONV_CHECK_ARG(grid < (1ll << 31), "depthwise: grid too large");
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    if constexpr (std::is_same_v<T, double>) {
      set_error("depthwise: fp64 unsupported");
      return EINCONV_ERR_UNSUPPORTED;
    } else {
      return dispatch_taps(p.k_h, p.k_w, [&](auto kh, auto kw) -> int {
        auto kern = dw_fwd_kernel<T, decltype(kh)::value, decltype(kw)::value>;
        if (smem > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dw_fwd)");
        }
        kern<<<(unsigned)grid, 256, smem, stream>>>((const T*)x, (const T*)w, (const T*)bias, (T*)y, p);
        EINCONV_CHECK_LAUNCH("dw_fwd_kernel");
        return EINCONV_OK;
      });
    }
  });
}

int dw_conv_input_vjp(const void* w, const void* v, void* gx, int dtype, const DevGeom& g,
                      cudaStream_t stream) {
  DwParams p{};
  fill_dw(p, g);
  // stride 1: the adjoint is the same stencil with flipped taps and padding K-1-p (>= 0 required)
  if (dtype != EINCONV_F64 && p.s_h == 1 && p.s_w == 1 && p.k_h - 1 - p.p_h >= 0 &&
      p.k_w - 1 - p.p_w >= 0) {
    DwFastParams f{};
    f.planes_out = p.batch * p.channels; f.chan_out = p.channels;
    f.mult = p.mult; f.J = p.mult; f.flip = 1; f.in_chan = p.channels * p.mult;
    f.in_h = p.out_h; f.in_w = p.out_w; f.out_h = p.in_h; f.out_w = p.in_w;
    f.p_h = p.k_h - 1 - p.p_h; f.p_w = p.k_w - 1 - p.p_w;
    int fast = dispatch_dtype(dtype, [&](auto* tag) -> int {
      using T = std::remove_pointer_t<decltype(tag)>;
      if constexpr (std::is_same_v<T, double>) return 0;
      else return try_dw_direct<T>(v, w, nullptr, gx, f, p.k_h, p.k_w, 1, 1, p.d_h, p.d_w, stream);
    });
    if (fast != 0) return fast > 0 ? EINCONV_OK : fast;
  }
  // staged cotangent window: all columns any input column can touch
  auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
  const int ow_lo = fdiv(p.p_w - p.d_w * (p.k_w - 1), p.s_w);
  const int ow_hi = (p.in_w - 1 + p.p_w) / p.s_w;
  p.cols = ow_hi - ow_lo + 1;
  auto rows_for = [&](int th) {
    // worst case over tile origins
    return (th - 1 + p.d_h * (p.k_h - 1)) / p.s_h + 2;
  };
  int th = p.in_h;
  while (th > 1 && (long long)rows_for(th) * p.cols * 4 > 32 * 1024) th = (th + 1) / 2;
  EINCONV_CHECK_ARG((long long)rows_for(th) * p.cols * 4 <= 200 * 1024, "depthwise: row too wide");
  p.tile_h = th; p.n_tiles = ceil_div(p.in_h, th); p.rows_cap = rows_for(th);
  p.w_div = FastDiv((uint32_t)p.in_w);
  const size_t smem = (size_t)p.rows_cap * p.cols * 4;
  const long long grid = (long long)p.batch * p.channels * p.n_tiles;
  EINCONV_CHECK_ARG(grid < (1ll << 31), "depthwise: grid too large");
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    if constexpr (std::is_same_v<T, double>) {
      set_error("depthwise: fp64 unsupported");
      return EINCONV_ERR_UNSUPPORTED;
    } else {
      return dispatch_taps(p.k_h, p.k_w, [&](auto kh, auto kw) -> int {
        auto kern = dw_input_vjp_kernel<T, decltype(kh)::value, decltype(kw)::value>;
        if (smem > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dw_input_vjp)");
        }
        kern<<<(unsigned)grid, 256, smem, stream>>>((const T*)w, (const T*)v, (T*)gx, p);
        EINCONV_CHECK_LAUNCH("dw_input_vjp_kernel");
        return EINCONV_OK;
      });
    }
  });
}

}  // namespace einconv
