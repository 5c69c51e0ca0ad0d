// Weight VJP (expressions/convNd_weight_vjp.py:51-58) on tcgen05 from channels-last padded copies,
// stride 1, any N / dilation / padding / groups, 16-bit dtypes.
//
//   gW[g*Cog + co, ci, k..] = sum_{b, o..} V[b, g*Cog + co, o..] * X[b, g*Cg + ci, o.. + dil*k.. - pad..]
//
// Same row-shift trick as conv_tma.cu: with xp[rows, channels] the zero-padded channels-last copy of X
// and vp[rows, channels] the cotangent laid on the SAME padded grid (zero where o_d >= O_d), flattened
// padded coordinate r, the contraction is  gW_t[ci, co] = sum_r xp[r + shift(t), ci] * vp[r, co]:
// a GEMM whose reduction runs over ROWS, so both operands are MN-major (channels contiguous).
// One staged slab of xp rows [r0 + group_shift, + KB + span) serves every tap of a tap group: the
// A operand of a tap PAIR (t1, t2) is a single M = 128 MN-major descriptor whose two 64-channel
// blocks are LBO = (shift(t2) - shift(t1)) rows apart, starting shift(t1) rows into the slab; the
// B operand is the vp tile [KB rows, BN channels].  ceil(tpg / 2) accumulators of BN columns live in
// TMEM for the whole K range of the CTA (split-K over CTAs), then go to fp32 partials which a small
// finish kernel sums, casts and stores as gW[co, ci, t].
//
// Compared with tc_tma.cu (one TMA tile per tap and k-block, L2 -> SM traffic = prod(K) * |X|) the
// slab is fetched once per tap group: (1 + span / KB) * |X| per group.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "dispatch.cuh"
#include "tc_ptx.cuh"

namespace einconv {
namespace wslab {

using namespace ptx;

constexpr int kThreads = 192;
constexpr int kRowBytes = 128;  // 64 channels of a 16-bit type: one swizzle-128B row
constexpr int kCB = 64;         // channels per block
constexpr int kMaxTpg = 64;
constexpr int kMaxGroups = 128;
constexpr int kSmemBudget = 216 * 1024;
constexpr int kBoxRows = 64;

struct Params {
  // work item = (conv group, ci block, tap group, co tile) x K split
  int groups, n_cblocks, n_tgroups, n_cotiles, k_splits;
  int tpg, n_pairs, taps;
  int kb_rows;            // rows of vp per k-block
  int slab_rows;          // rows of xp per k-block (kb_rows + span, rounded to kBoxRows)
  int slab_bytes, vtile_bytes, stage_bytes, n_stages;
  int bn, n_vblocks;      // BN columns = n_vblocks TMA boxes of 64 channels
  int n_kblocks;          // ceil(rows_total / kb_rows)
  int cg, cog;            // channels per group
  int group_shift[kMaxGroups];
  uint32_t pair_off16[kMaxTpg / 2];  // (row offset of the pair's first tap inside the slab) * 128 / 16
  uint32_t pair_lbo16[kMaxTpg / 2];  // (row distance between the pair's taps) * 128 / 16   (0: single tap)
  // accumulator block -> tap of the group (-1: unused): [MMA][M half: first / second unit][N block]
  signed char tap_tbl[kMaxTpg / 2][2][2];
  // quad mode (BN = 64 only): the B operand is N = 128 with its two 64-column blocks `quad_rows` ROWS
  // apart in a vp slab that starts quad_rows before the k-block, so one MMA pairs each of its two xp
  // units with vp[r - quad_rows] (tap shift + quad_rows, N block 0) and vp[r] (N block 1): four taps
  // per M = 128, N = 128 instruction, which is tensor-bound where N = 64 is shared-memory-read bound.
  int quad_rows;
  int v_rows;             // rows of vp staged per 64-channel block (kb_rows, or kb_rows + 64 in quad mode)
  int acc_cols;           // TMEM columns per MMA accumulator (bn, or 128 in quad mode)
  float* partial;         // [k_splits][groups * cog * cg * taps]
  long long gw_elems;
};

// MN-major, 128-byte swizzle (cute::UMMA canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO))):
// 64 16-bit channels per row, 8-row groups SBO = 1024 bytes apart, 64-channel blocks LBO apart.
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t start16, uint32_t lbo16) {
  uint64_t d = 0;
  d |= (uint64_t)(start16 & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// both operands MN-major: a_major @15, b_major @16
__host__ __device__ constexpr uint32_t make_idesc_mn(int fmt, int M, int N) {
  return make_idesc(fmt, M, N) | (1u << 15) | (1u << 16);
}

struct Bars {
  uint64_t full[4], empty[4];
  uint64_t done;
  uint32_t tmem_base;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_slab_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_v,
                  const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;
  Bars* bars = reinterpret_cast<Bars*>(ring + (size_t)p.n_stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int FMT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

  // decode the work item
  int item = blockIdx.x;
  const int split = item % p.k_splits;
  item /= p.k_splits;
  const int cotile = item % p.n_cotiles;
  item /= p.n_cotiles;
  const int tg = item % p.n_tgroups;
  item /= p.n_tgroups;
  const int cblock = item % p.n_cblocks;
  const int g = item / p.n_cblocks;
  const int kb_per = (p.n_kblocks + p.k_splits - 1) / p.k_splits;
  const int kb_begin = split * kb_per;
  const int kb_end = min(p.n_kblocks, kb_begin + kb_per);
  const int n_kb = max(0, kb_end - kb_begin);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_barrier_init();
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_v);
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.n_pairs * p.acc_cols)) tmem_cols <<= 1;
  if (warp == 1) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int st = 0;
    uint32_t ph = 0;
    const uint32_t stage_tx = (uint32_t)(p.slab_rows * kRowBytes + p.n_vblocks * p.v_rows * kRowBytes);
    const int col_x = g * p.cg + cblock * kCB;
    const int col_v = g * p.cog + cotile * p.bn;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&bars->empty[st], ph ^ 1);
      {
        // one box per lane: a single issuing thread sustains only ~one 8 KB box per 550 cycles
        if (lane == 0) mbar_expect_tx(&bars->full[st], stage_tx);
        __syncwarp();
        unsigned char* stage = ring + (size_t)st * p.stage_bytes;
        const int r0 = kb * p.kb_rows;
        const int xrow0 = r0 + p.group_shift[tg];
        const int x_boxes = p.slab_rows / kBoxRows, v_boxes_per_block = p.v_rows / kBoxRows;
        const int n_boxes = x_boxes + p.n_vblocks * v_boxes_per_block;
        unsigned char* vt = stage + p.slab_bytes;
        for (int bx = lane; bx < n_boxes; bx += 32) {
          if (bx < x_boxes) {
            tma_load_2d(stage + (size_t)bx * kBoxRows * kRowBytes, &map_x, &bars->full[st], col_x,
                        xrow0 + bx * kBoxRows);
          } else {
            const int v = bx - x_boxes, nb = v / v_boxes_per_block, r = (v - nb * v_boxes_per_block) * kBoxRows;
            // quad mode: the slab starts quad_rows early (negative rows of the first k-block zero-fill)
            tma_load_2d(vt + (size_t)(nb * p.v_rows + r) * kRowBytes, &map_v, &bars->full[st], col_v + nb * kCB,
                        r0 - p.quad_rows + r);
          }
        }
      }
      __syncwarp();
      if (++st == p.n_stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_mn(FMT, 128, p.acc_cols);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, slab16 = (uint32_t)p.slab_bytes >> 4;
    // distance between the 64-column blocks of B: the next channel block of vp, or (quad mode) the same
    // channels quad_rows further down the slab
    const uint32_t vlbo16 = (uint32_t)((p.quad_rows ? p.quad_rows : p.v_rows) * kRowBytes) >> 4;
    const int k_steps = p.kb_rows / 16;
    int st = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(&bars->full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a16 = ring16 + (uint32_t)st * stage16;
        const uint64_t b_base = make_desc_mnmajor(a16 + slab16, vlbo16);
        for (int pr = 0; pr < p.n_pairs; ++pr) {
          const uint64_t a_base = make_desc_mnmajor(a16 + p.pair_off16[pr], p.pair_lbo16[pr]);
          const uint32_t d_tmem = tmem_base + (uint32_t)(pr * p.acc_cols);
          // 16 rows (two 8-row groups of 1024 bytes) per MMA: + 2048 bytes = + 128 in the address field
          for (int k = 0; k < k_steps; ++k)
            umma<false>(d_tmem, a_base + (uint64_t)(k * 128), b_base + (uint64_t)(k * 128), idesc,
                        (uint32_t)(kb | k));
        }
        umma_commit(&bars->empty[st]);
        if (kb + 1 == n_kb) umma_commit(&bars->done);
      }
      __syncwarp();
      if (++st == p.n_stages) { st = 0; ph ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;
    if (n_kb > 0) {
      mbar_wait_relaxed(&bars->done, 0);
      tc_fence_after();
    }
    // accumulator row = unit * 64 + ci_local  (unit 0 / 1 = first / second tap of the pair)
    const int row = quarter * 32 + lane;
    const int unit = row >> 6;
    const int ci = cblock * kCB + (row & 63);
    float* part = p.partial + (long long)split * p.gw_elems;
    for (int pr = 0; pr < p.n_pairs; ++pr) {
      for (int c0 = 0; c0 < p.acc_cols; c0 += 16) {
        const int nblock = p.quad_rows ? (c0 >> 6) : 0;
        const int tl = p.tap_tbl[pr][unit][nblock];  // tap within the group
        const bool tap_ok = tl >= 0 && ci < p.cg;
        const int t = tg * p.tpg + tl;
        const int cbase = p.quad_rows ? (c0 & 63) : c0;
        uint32_t r[16];
        if (n_kb > 0) {
          tmem_ld16_nowait(tmem_base + (uint32_t)(pr * p.acc_cols + c0) + ((uint32_t)(quarter * 32) << 16), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        if (tap_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = cotile * p.bn + cbase + j;
            if (co < p.cog)
              part[(((long long)g * p.cog + co) * p.cg + ci) * p.taps + t] = __uint_as_float(r[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// absmax != nullptr: the operands were scaled fp16 copies of fp32 tensors; un-scale by 1 / (s_x * s_v)
template <typename T>
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ partial, T* __restrict__ gw,
                                                           long long n, int splits,
                                                           const unsigned* __restrict__ absmax) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(long long)k * n + idx];
  if (absmax) s *= 1.f / (operand_scale(__ldg(absmax)) * operand_scale(__ldg(absmax + 1)));
  Cvt<T>::store(gw + idx, s);
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
struct Plan {
  bool ok = false;
  int n_dims;
  long long P[kMaxDims], pitch[kMaxDims], ptot, rows_total;
  int cxp, cvp;
  int taps, tpg, n_tgroups, n_pairs;
  int quad_rows = 0, v_rows = 0;  // quad mode: see Params
  int kb_rows, slab_rows, slab_bytes, vtile_bytes, stage_bytes, n_stages;
  int bn, n_vblocks, n_cotiles, n_cblocks, n_kblocks, k_splits;
  long long xp_bytes, vp_bytes, partial_bytes, gw_elems;
  size_t smem;
};

static Plan make_plan(const DevGeom& g, int dtype) {
  Plan pl;
  // fp32 tensors (TF32 mode) run on power-of-two-scaled fp16 operand copies: same 10-bit mantissa
  if (dtype != EINCONV_BF16 && dtype != EINCONV_F16 && dtype != EINCONV_F32) return pl;
  if (getenv("EINCONV_NO_WGRAD_SLAB")) return pl;
  if (dtype == EINCONV_F32 && getenv("EINCONV_WGRAD_TF32")) return pl;
  const int N = g.n_dims;
  pl.n_dims = N;
  for (int d = 0; d < N; ++d)
    if (g.stride[d] != 1) return pl;
  if (g.groups > 1 && (g.c_in_g % 8 != 0 || g.c_out_g % 8 != 0)) return pl;  // 16-byte aligned TMA columns
  if (g.batch == 0 || g.c_in_g == 0 || g.c_out_g == 0) return pl;
  pl.ptot = 1;
  for (int d = N - 1; d >= 0; --d) {
    pl.P[d] = std::max<long long>(g.in[d] + g.pad[d], (long long)(g.out[d] - 1) + (long long)g.dil[d] * (g.k[d] - 1) + 1);
    pl.pitch[d] = pl.ptot;
    pl.ptot *= pl.P[d];
    if (pl.ptot >= (1ll << 31)) return pl;
  }
  pl.rows_total = pl.ptot * g.batch;
  long long max_shift = 0;
  pl.taps = g.k_total;
  for (int d = 0; d < N; ++d) max_shift += (long long)g.dil[d] * (g.k[d] - 1) * pl.pitch[d];
  if (pl.rows_total + max_shift + 4096 >= (1ll << 31)) return pl;
  pl.cxp = (g.groups * g.c_in_g + 7) / 8 * 8;
  pl.cvp = (g.groups * g.c_out_g + 7) / 8 * 8;

  // N tile: whole 64-channel boxes of the cotangent (or one partial box)
  if (g.c_out_g <= 64) {
    pl.bn = (g.c_out_g + 15) / 16 * 16;
    pl.n_vblocks = 1;
    pl.n_cotiles = 1;
  } else {
    const int blocks = (g.c_out_g + 63) / 64;
    pl.n_cotiles = (blocks + 3) / 4;                       // at most 256 columns per tile
    pl.n_vblocks = (blocks + pl.n_cotiles - 1) / pl.n_cotiles;
    pl.bn = pl.n_vblocks * 64;
  }
  pl.n_cblocks = (g.c_in_g + kCB - 1) / kCB;

  // tap groups (trailing `level` dims share a slab) and k-block size: as many taps per slab as TMEM
  // (pairs * BN <= 512 columns) and shared memory (two stages) allow; prefer larger k-blocks.
  for (int level = N; level >= 0 && !pl.ok; --level) {
    long long tpg = 1, span = 0;
    for (int d = N - level; d < N; ++d) {
      tpg *= g.k[d];
      span += (long long)g.dil[d] * (g.k[d] - 1) * pl.pitch[d];
    }
    if (tpg > kMaxTpg || pl.taps / tpg > kMaxGroups) continue;
    const int pairs = (int)((tpg + 1) / 2);
    // quad mode: units of up to two taps adjacent along the innermost dimension, two units per MMA;
    // worth it when its N = 128 MMAs (64 cycles each) beat the N = 64 pairs (48 cycles, smem-bound)
    const int kw = level >= 1 ? g.k[N - 1] : 1;
    const int units = (int)(tpg / kw) * ((kw + 1) / 2);
    const int quads = (units + 1) / 2;
    const long long quad_rows = (long long)g.dil[N - 1];
    const bool quad_ok = pl.bn == 64 && pl.n_vblocks == 1 && kw >= 2 && quads * 128 <= 512 && quad_rows <= kBoxRows &&
                         quads * 64 < pairs * 48 && !getenv("EINCONV_WGRAD_NO_QUAD");
    if (!quad_ok && (long long)pairs * pl.bn > 512) continue;
    for (int kb : {256, 128, 64}) {
      const long long rows = (kb + span + kBoxRows - 1) / kBoxRows * kBoxRows;
      for (int quad = quad_ok ? 1 : 0; quad >= 0 && !pl.ok; --quad) {
        if (!quad && (long long)pairs * pl.bn > 512) continue;
        const long long v_rows = quad ? kb + kBoxRows : kb;
        const long long stage = rows * kRowBytes + (long long)pl.n_vblocks * v_rows * kRowBytes;
        if (2 * stage > kSmemBudget) continue;
        pl.tpg = (int)tpg;
        pl.n_pairs = quad ? quads : pairs;
        pl.quad_rows = quad ? (int)quad_rows : 0;
        pl.v_rows = (int)v_rows;
        pl.n_tgroups = pl.taps / pl.tpg;
        pl.kb_rows = kb;
        pl.slab_rows = (int)rows;
        pl.slab_bytes = (int)(rows * kRowBytes);
        pl.vtile_bytes = (int)(pl.n_vblocks * v_rows * kRowBytes);
        pl.stage_bytes = (int)stage;
        pl.n_stages = (int)std::min<long long>(4, kSmemBudget / stage);
        pl.ok = true;
      }
      if (pl.ok) break;
    }
  }
  if (!pl.ok) return pl;
  // quad mode reads vp[r - quad_rows]: r must run quad_rows past the last row
  pl.n_kblocks = (int)((pl.rows_total + pl.quad_rows + pl.kb_rows - 1) / pl.kb_rows);
  const long long types = (long long)g.groups * pl.n_cblocks * pl.n_tgroups * pl.n_cotiles;
  // split K to fill whole waves of the 148 SMs (one CTA per SM)
  {
    const long long cap = std::max<long long>(1, std::min<long long>(pl.n_kblocks / 4, 256));
    pl.gw_elems = (long long)g.groups * g.c_out_g * g.c_in_g * pl.taps;
    // relative cost of one more partial copy (written once, read once by the finish kernel at ~4 TB/s)
    // against the main loop at a nominal 400 TFLOP/s
    const double t_main = 2.0 * (double)pl.rows_total * (double)pl.gw_elems / 400e12;
    const double w_partial = (2.0 * (double)pl.gw_elems * 4.0 / 4e12) / std::max(t_main, 1e-9);
    double best = 1e30;
    int best_s = 1;
    for (long long s = 1; s <= cap; ++s) {
      const long long ctas = types * s;
      const long long rounds = (ctas + num_sms() - 1) / num_sms();
      const double t = (double)rounds * num_sms() / (double)ctas + w_partial * (double)s;
      if (t < best) {
        best = t;
        best_s = (int)s;
      }
    }
    pl.k_splits = best_s;
  }
  if (types * pl.k_splits >= (1ll << 31)) {
    pl.ok = false;
    return pl;
  }
  pl.gw_elems = (long long)g.groups * g.c_out_g * g.c_in_g * pl.taps;
  pl.xp_bytes = (((pl.rows_total + max_shift + 2 * kBoxRows) * pl.cxp * 2) + 1023) / 1024 * 1024;
  pl.vp_bytes = ((pl.rows_total * pl.cvp * 2) + 1023) / 1024 * 1024;
  pl.partial_bytes = ((long long)pl.k_splits * pl.gw_elems * 4 + 1023) / 1024 * 1024;
  pl.smem = (size_t)pl.n_stages * pl.stage_bytes + sizeof(Bars) + 64;
  return pl;
}

bool supported(const DevGeom& g, int dtype) { return make_plan(g, dtype).ok; }

int64_t workspace_bytes(const DevGeom& g, int dtype) {
  const Plan pl = make_plan(g, dtype);
  if (!pl.ok) return 0;
  return pl.xp_bytes + pl.vp_bytes + pl.partial_bytes + 2048;
}

// T: operand type of the kernel (bf16 / fp16); fp32 inputs are converted to scaled fp16 while packing
template <typename T>
static int run_t(const Plan& pl, const void* x, const void* v, void* gw, int dtype, const DevGeom& g,
                 void* workspace, cudaStream_t stream) {
  uintptr_t base = ((uintptr_t)workspace + 1023) & ~(uintptr_t)1023;
  unsigned* absmax = reinterpret_cast<unsigned*>(base);  // [0] max|x|, [1] max|v| (fp32 inputs only)
  base += 1024;
  T* xp = reinterpret_cast<T*>(base);
  T* vp = reinterpret_cast<T*>(base + pl.xp_bytes);
  float* partial = reinterpret_cast<float*>(base + pl.xp_bytes + pl.vp_bytes);
  const int N = g.n_dims;
  const int op_dtype = dtype == EINCONV_F32 ? EINCONV_F16 : dtype;
  int zero_pad[kMaxDims] = {0};
  // xp: X on the padded grid; rows beyond B * Ptot are never written and are excluded from the tensor
  // map (TMA zero-fills them).  vp: V on the same grid, zero where o_d >= O_d.
  // (running the two transposes concurrently on a forked stream was measured and gains nothing: they
  // are limited by DRAM efficiency, not by latency)
  int st;
  if (dtype == EINCONV_F32) {
    cudaError_t e = cudaMemsetAsync(absmax, 0, 8, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(absmax)");
    st = launch_absmax((const float*)x, (long long)g.batch * g.groups * g.c_in_g * g.in_total, absmax, stream);
    if (st) return st;
    st = launch_absmax((const float*)v, (long long)g.batch * g.groups * g.c_out_g * g.out_total, absmax + 1, stream);
    if (st) return st;
    st = ctma::pack_channels_last_f32_to_f16((const float*)x, xp, N, g.batch, g.groups * g.c_in_g, pl.cxp, g.in, pl.P,
                                             g.pad, absmax, stream);
    if (st) return st;
    st = ctma::pack_channels_last_f32_to_f16((const float*)v, vp, N, g.batch, g.groups * g.c_out_g, pl.cvp, g.out,
                                             pl.P, zero_pad, absmax + 1, stream);
    if (st) return st;
  } else {
    st = ctma::pack_channels_last(x, xp, dtype, N, g.batch, g.groups * g.c_in_g, pl.cxp, g.in, pl.P, g.pad, stream);
    if (st) return st;
    st = ctma::pack_channels_last(v, vp, dtype, N, g.batch, g.groups * g.c_out_g, pl.cvp, g.out, pl.P, zero_pad, stream);
    if (st) return st;
  }

  CUtensorMap map_x, map_v;
  st = ctma::encode_2d(&map_x, op_dtype, xp, pl.rows_total, pl.cxp, kCB, kBoxRows, 128);
  if (st) return st;
  st = ctma::encode_2d(&map_v, op_dtype, vp, pl.rows_total, pl.cvp, kCB, kBoxRows, 128);
  if (st) return st;

  Params p{};
  p.groups = g.groups;
  p.n_cblocks = pl.n_cblocks;
  p.n_tgroups = pl.n_tgroups;
  p.n_cotiles = pl.n_cotiles;
  p.k_splits = pl.k_splits;
  p.tpg = pl.tpg;
  p.n_pairs = pl.n_pairs;
  p.taps = pl.taps;
  p.kb_rows = pl.kb_rows;
  p.slab_rows = pl.slab_rows;
  p.slab_bytes = pl.slab_bytes;
  p.vtile_bytes = pl.vtile_bytes;
  p.stage_bytes = pl.stage_bytes;
  p.n_stages = pl.n_stages;
  p.bn = pl.bn;
  p.n_vblocks = pl.n_vblocks;
  p.n_kblocks = pl.n_kblocks;
  p.cg = g.c_in_g;
  p.cog = g.c_out_g;
  long long tap_shift[kMaxTpg];
  for (int t = 0; t < pl.taps; ++t) {
    long long shift = 0;
    int rem = t;
    for (int d = N - 1; d >= 0; --d) {
      const int kd = rem % g.k[d];
      rem /= g.k[d];
      shift += (long long)g.dil[d] * kd * pl.pitch[d];
    }
    if (t % pl.tpg == 0) p.group_shift[t / pl.tpg] = (int)shift;
    if (t < pl.tpg) tap_shift[t] = shift;
  }
  p.quad_rows = pl.quad_rows;
  p.v_rows = pl.v_rows;
  p.acc_cols = pl.quad_rows ? 128 : pl.bn;
  // units: the M = 64 halves of the MMAs.  Pair mode: one tap each.  Quad mode: the tap at an even
  // position of the innermost kernel dimension (N block 1, vp[r]) and its right neighbour, if any
  // (N block 0, vp[r - quad_rows]).
  int unit_tap[kMaxTpg][2], n_units = 0;
  if (pl.quad_rows) {
    const int kw = g.k[N - 1];
    for (int row = 0; row < pl.tpg / kw; ++row)
      for (int i = 0; 2 * i < kw; ++i) {
        unit_tap[n_units][1] = row * kw + 2 * i;
        unit_tap[n_units][0] = 2 * i + 1 < kw ? row * kw + 2 * i + 1 : -1;
        ++n_units;
      }
  } else {
    for (int t = 0; t < pl.tpg; ++t) {
      unit_tap[n_units][0] = t;
      unit_tap[n_units][1] = -1;
      ++n_units;
    }
  }
  EINCONV_CHECK_ARG((n_units + 1) / 2 == pl.n_pairs, "wgrad_slab: inconsistent plan");
  for (int pr = 0; pr < pl.n_pairs; ++pr) {
    const int u1 = 2 * pr, u2 = 2 * pr + 1 < n_units ? 2 * pr + 1 : -1;
    const int first = pl.quad_rows ? 1 : 0;  // the N block that holds the unit's un-shifted tap
    const long long s1 = tap_shift[unit_tap[u1][first]];
    const long long s2 = u2 >= 0 ? tap_shift[unit_tap[u2][first]] : s1;
    p.pair_off16[pr] = (uint32_t)((s1 * kRowBytes) >> 4);
    p.pair_lbo16[pr] = (uint32_t)(((s2 - s1) * kRowBytes) >> 4);
    EINCONV_CHECK_ARG(s2 >= s1 && (s2 - s1) * kRowBytes < (1 << 18),
                      "wgrad_slab: tap distance exceeds the descriptor range");
    for (int nb = 0; nb < 2; ++nb) {
      p.tap_tbl[pr][0][nb] = (signed char)unit_tap[u1][nb];
      p.tap_tbl[pr][1][nb] = (signed char)(u2 >= 0 ? unit_tap[u2][nb] : -1);
    }
  }
  p.partial = partial;
  p.gw_elems = pl.gw_elems;

  auto kern = wgrad_slab_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad_slab)");
  const long long grid = (long long)g.groups * pl.n_cblocks * pl.n_tgroups * pl.n_cotiles * pl.k_splits;
  kern<<<(unsigned)grid, kThreads, pl.smem, stream>>>(map_x, map_v, p);
  EINCONV_CHECK_LAUNCH("wgrad_slab_kernel");
  const unsigned fin_blocks = (unsigned)((pl.gw_elems + 255) / 256);
  if (dtype == EINCONV_F32)
    wgrad_finish_kernel<float><<<fin_blocks, 256, 0, stream>>>(partial, (float*)gw, pl.gw_elems, pl.k_splits, absmax);
  else
    wgrad_finish_kernel<T><<<fin_blocks, 256, 0, stream>>>(partial, (T*)gw, pl.gw_elems, pl.k_splits, nullptr);
  EINCONV_CHECK_LAUNCH("wgrad_finish_kernel");
  return EINCONV_OK;
}

int run(const void* x, const void* v, void* gw, int dtype, const DevGeom& g, void* workspace,
        int64_t workspace_bytes_given, cudaStream_t stream) {
  const Plan pl = make_plan(g, dtype);
  if (!pl.ok) {
    set_error("wgrad_slab: unsupported geometry");
    return EINCONV_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes_given < pl.xp_bytes + pl.vp_bytes + pl.partial_bytes + 2048) {
    set_error("wgrad_slab: workspace too small");
    return EINCONV_ERR_WORKSPACE;
  }
  if (dtype == EINCONV_BF16) return run_t<__nv_bfloat16>(pl, x, v, gw, dtype, g, workspace, stream);
  return run_t<__half>(pl, x, v, gw, dtype, g, workspace, stream);  // fp16 and (scaled) fp32
}

}  // namespace wslab
}  // namespace einconv
