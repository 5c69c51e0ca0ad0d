// K1: unfoldNd (im2col) and its adjoint fold (col2im) for arbitrary N, stride, dilation, padding.
//
// Replaces the einsum of einconv/expressions/convNd_unfold.py:42-53 (three one-hot GEMMs in the
// reference) by a gather:  out[b, c, k.., o..] = x[b, c, s*o + d*k - p] or 0.
//
// Work decomposition (input-stationary): a CTA owns `cb` consecutive (b, c) planes, ONE virtual
// coordinate of every outer spatial dimension (all but the last two; "virtual" = shifted into the
// zero-padded frame so that padding planes are visited too) and a tile of output rows of the
// second-to-last dimension.  It stages exactly the input pixels those outputs can touch into
// shared memory once (zero-filled where the virtual coordinate is padding, compacted by
// gcd(stride, dilation) so never-read rows / columns are neither loaded nor stored), then streams
// every kernel tap that maps onto this plane to global memory as contiguous runs: for a fixed tap
// the tile's outputs are `th * O_w` consecutive elements of `out`, written by consecutive lanes.
// x is therefore read once from L2/HBM (plus row halos) while `out` (prod(k) times larger) is
// written with full-line coalesced stores — the kernel is bound by HBM write bandwidth.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "dispatch.cuh"
#include "tc_ptx.cuh"

namespace einconv {

constexpr int kMaxOuterTaps = 256;  // prod(k) over the outer dims handled by one CTA

struct UnfoldParams {
  DevGeom g;  // n_dims >= 2 (a dummy unit dimension is inserted in front of a 1-d problem)
  long long xs_b, xs_c;       // element strides of x: batch, channel
  long long xs[kMaxDims];     // element strides of the spatial dims
  int channels;               // C = groups * c_in_g
  int n_outer;                // n_dims - 2
  int vstep[kMaxDims];        // gcd(stride, dilation) of outer dim d: only every vstep-th virtual
                              // coordinate can be reached by a tap
  int vcount[kMaxDims];       // number of visited virtual coordinates of outer dim d
  int v_total;                // prod(vcount[0..n_outer))
  int k_outer_total;          // prod(k[0..n_outer))
  int k_inner_total;          // k[H] * k[W]
  int tile_h, n_tiles;        // output rows per tile, tiles per plane
  int cb;                     // (b, c) planes per CTA
  int gh, gw;                 // gcd(stride, dilation) of the two inner dims
  int sh, dh, sw, dw;         // stride/g and dilation/g of the inner dims
  int rows_cap;               // staged rows for a full tile
  int cols;                   // staged columns
  int owv;                    // O_w / V: vectors per output row
  int lpr_shift;              // log2(lanes per output row); rows per warp pass = 32 >> lpr_shift
  FastDiv kw_div;             // / K_w
  FastDiv kinner_div;         // / (K_h * K_w)
  FastDiv tiles_div;          // / n_tiles
  FastDiv vtotal_div;         // / v_total
  FastDiv chan_div;           // / channels
  FastDiv vc_div[kMaxDims];   // / vcount[d]
  FastDiv ko_div[kMaxDims];   // / k[d]       (outer dims)
  FastDiv so_div[kMaxDims];   // / stride[d]  (outer dims)
};

static __host__ __device__ inline int igcd(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

template <typename T, int V> struct alignas(sizeof(T) * V) VecOf { T v[V]; };

// streaming (evict-first) vector store: `out` is written once and never re-read by this kernel
template <typename T, int V>
__device__ __forceinline__ void store_stream(VecOf<T, V>* dst, const VecOf<T, V>& val) {
  constexpr int bytes = sizeof(T) * V;
  if constexpr (bytes == 16) {
    __stcs(reinterpret_cast<uint4*>(dst), *reinterpret_cast<const uint4*>(&val));
  } else if constexpr (bytes == 8) {
    __stcs(reinterpret_cast<uint2*>(dst), *reinterpret_cast<const uint2*>(&val));
  } else if constexpr (bytes == 4) {
    __stcs(reinterpret_cast<unsigned int*>(dst), *reinterpret_cast<const unsigned int*>(&val));
  } else {
    *dst = val;
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 8) unfold_plane_kernel(const T* __restrict__ x,
                                                           T* __restrict__ out,
                                                           const __grid_constant__ UnfoldParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* plane = reinterpret_cast<T*>(smem_raw);
  using VecT = VecOf<T, V>;
  // CTA-uniform bookkeeping, computed once by warp 0
  __shared__ int s_combo_kq[kMaxOuterTaps];     // outer tap combinations that land on this plane
  __shared__ int s_combo_oflat[kMaxOuterTaps];  // ... and their flattened outer output index
  __shared__ int s_n_combos;
  __shared__ int s_plane_real;
  __shared__ long long s_x_outer_off;

  const DevGeom& g = p.g;
  const int H = g.n_dims - 2, W = g.n_dims - 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;

  // ---- CTA coordinates: tile fastest, then virtual outer coordinate, then plane group
  uint32_t rest, tile, vflat, bcg;
  p.tiles_div.divmod(blockIdx.x, rest, tile);
  p.vtotal_div.divmod(rest, bcg, vflat);

  if (warp == 0) {
    // virtual outer coordinates of this CTA (input frame; outside [0, in) means padding)
    int v[kMaxDims];
    bool plane_real = true;
    long long x_off = 0;
    uint32_t rem = vflat;
    for (int d = p.n_outer - 1; d >= 0; --d) {
      uint32_t q, r;
      p.vc_div[d].divmod(rem, q, r);
      rem = q;
      const int vd = (int)r * p.vstep[d] - g.pad[d];
      v[d] = vd;
      if (vd < 0 || vd >= g.in[d]) plane_real = false;
      else x_off += (long long)vd * p.xs[d];
    }
    // outer tap combinations kq (lane-strided) that map this plane to a valid output coordinate
    int count = 0;
    for (int kq0 = 0; kq0 < p.k_outer_total; kq0 += 32) {
      const int kq = kq0 + lane;
      bool ok = kq < p.k_outer_total;
      int oflat = 0, omul = 1;
      uint32_t krem = ok ? (uint32_t)kq : 0u;
      for (int d = p.n_outer - 1; d >= 0; --d) {
        uint32_t q, kd;
        p.ko_div[d].divmod(krem, q, kd);
        krem = q;
        const int t = v[d] + g.pad[d] - g.dil[d] * (int)kd;
        uint32_t o = 0, r = 0;
        if (t >= 0) p.so_div[d].divmod((uint32_t)t, o, r);
        if (t < 0 || r != 0 || (int)o >= g.out[d]) ok = false;
        oflat += (int)o * omul;
        omul *= g.out[d];
      }
      const unsigned mask = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int slot = count + __popc(mask & ((1u << lane) - 1u));
        s_combo_kq[slot] = kq;
        s_combo_oflat[slot] = oflat;
      }
      count += __popc(mask);
    }
    if (lane == 0) {
      s_n_combos = count;
      s_plane_real = plane_real ? 1 : 0;
      s_x_outer_off = x_off;
    }
  }
  __syncthreads();
  const int n_combos = s_n_combos;
  if (n_combos == 0) return;  // no tap reaches this plane
  const bool plane_real = s_plane_real != 0;
  const long long x_outer_off = s_x_outer_off;

  const int oh0 = (int)tile * p.tile_h;
  const int th = min(p.tile_h, g.out[H] - oh0);
  const int rows = (p.sh * (th - 1) + p.dh * (g.k[H] - 1)) + 1;  // staged rows (compacted by gh)
  const int cols = p.cols;
  const int plane_elems = p.rows_cap * cols;
  const int bc0 = (int)bcg * p.cb;
  const int n_planes = min(p.cb, g.batch * p.channels - bc0);

  // ---- stage: plane[pl][jr][jc] = x[b, c, v.., r0 + jr*gh, -p_w + jc*gw] or 0
  {
    const int r0 = g.stride[H] * oh0 - g.pad[H];
    const int gh = p.gh, gw = p.gw, pad_w = g.pad[W], in_w = g.in[W], in_h = g.in[H];
    const long long xs_h = p.xs[H];
    const int xs_w = (int)p.xs[W];  // |stride| of the innermost dim fits in 32 bits (checked on host)
    for (int pl = 0; pl < n_planes; ++pl) {
      uint32_t b, c;
      p.chan_div.divmod((uint32_t)(bc0 + pl), b, c);
      const T* src_plane = x + (long long)b * p.xs_b + (long long)c * p.xs_c + x_outer_off;
      T* dst_plane = plane + pl * plane_elems;
#pragma unroll 4
      for (int jr = warp; jr < rows; jr += n_warps) {
        const int r = r0 + jr * gh;
        const bool row_real = plane_real && r >= 0 && r < in_h;
        const T* src = src_plane + (long long)r * xs_h;
        T* dst = dst_plane + jr * cols;
        for (int jc = lane; jc < cols; jc += 32) {
          const int cc = jc * gw - pad_w;
          T val = T(0);
          if (row_real && cc >= 0 && cc < in_w) val = __ldg(src + cc * xs_w);
          dst[jc] = val;
        }
      }
    }
  }
  __syncthreads();

  // ---- emit.  For a fixed (plane, outer tap combo, inner tap) the tile's outputs are th * O_w
  // consecutive elements of `out`.  Lanes are arranged as (32 >> lpr_shift) rows x (1 << lpr_shift)
  // vector columns, so the walk needs no division: consecutive lanes write consecutive vectors of
  // V elements (V | O_w: a vector never straddles an output row).  Everything the row loop needs
  // is held in registers; per-tap offsets come from a small shared table.
  __shared__ int s_tap_tab[256];
  const int k_inner = p.k_inner_total;
  int* tap_tab = (k_inner <= 256) ? s_tap_tab : nullptr;
  if (tap_tab) {
    for (int t = threadIdx.x; t < k_inner; t += blockDim.x) {
      uint32_t kh, kw;
      p.kw_div.divmod((uint32_t)t, kh, kw);
      tap_tab[t] = p.dh * (int)kh * cols + p.dw * (int)kw;
    }
    __syncthreads();
  }

  const int lpr_shift = p.lpr_shift;
  const int rpw = 32 >> lpr_shift;
  const int lr = lane >> lpr_shift;
  const int lc = lane & ((1 << lpr_shift) - 1);
  const int owv = p.owv;
  const int sw = p.sw;
  const int n_cp = n_combos * n_planes;
  // split a (plane, tap) run into row segments only if the CTA would otherwise have fewer than two
  // units per warp, keeping >= 3 warp passes per unit so its set-up stays amortised
  int seg_shift = 0;
  {
    const int passes = (th + rpw - 1) / rpw;
    while (((n_cp * k_inner) << seg_shift) < 2 * n_warps && (passes >> (seg_shift + 1)) >= 3) ++seg_shift;
  }
  const int seg_mask = (1 << seg_shift) - 1;
  const long long out_plane = (long long)g.k_total * g.out_total;  // elements per (b, c)
  const long long out_total = g.out_total;
  const int inner_out = g.out[H] * g.out[W];
  const int row_pitch = p.sh * cols;
  const int rows_per_seg = (th + seg_mask) >> seg_shift;
  const int units_per_cp = k_inner << seg_shift;        // units per (combo, plane)
  const int n_units = n_cp * units_per_cp;
  const int src_step = rpw * row_pitch;
  const int dst_step = rpw * owv;
  const int lane_src = lr * row_pitch + lc * (sw * V);
  const int lane_dst = lr * owv + lc;
  const bool lane_on = lc < owv;

  // warp-uniform walk over units; (cp, u) advance incrementally, no division in the loop
  int cp = 0, u = warp;
  int cur_cp = -1;
  const T* src_cp = plane;
  T* dst_cp = out;
  for (int unit = warp; unit < n_units; unit += n_warps, u += n_warps) {
    while (u >= units_per_cp) {
      u -= units_per_cp;
      ++cp;
    }
    if (cp != cur_cp) {
      cur_cp = cp;
      const int combo = cp / n_planes;
      const int pl = cp - combo * n_planes;
      // out[bc][kq * k_inner + tap][o_outer_flat * inner_out + (oh0 + ohl) * O_w + ow]
      dst_cp = out + (long long)(bc0 + pl) * out_plane +
               (long long)s_combo_kq[combo] * k_inner * out_total +
               (long long)s_combo_oflat[combo] * inner_out + (long long)oh0 * g.out[W];
      src_cp = plane + pl * plane_elems;
    }
    const int tap = u >> seg_shift;
    const int seg = u & seg_mask;
    const int row_begin = seg * rows_per_seg;
    const int n_rows = min(rows_per_seg, th - row_begin);
    int tap_off;
    if (tap_tab) {
      tap_off = tap_tab[tap];
    } else {
      uint32_t kh, kw;
      p.kw_div.divmod((uint32_t)tap, kh, kw);
      tap_off = p.dh * (int)kh * cols + p.dw * (int)kw;
    }
    const T* src = src_cp + tap_off + row_begin * row_pitch + lane_src;
    VecT* dst = reinterpret_cast<VecT*>(dst_cp + (long long)tap * out_total) +
                (row_begin * owv + lane_dst);
    if (owv <= 32) {
      // one pass per row group: the common case
      if (lane_on) {
#pragma unroll 2
        for (int r = lr; r < n_rows; r += rpw) {
          VecT val;
#pragma unroll
          for (int j = 0; j < V; ++j) val.v[j] = src[j * sw];
          store_stream<T, V>(dst, val);
          src += src_step;
          dst += dst_step;
        }
      }
    } else {
      for (int r = 0; r < n_rows; ++r) {  // lpr == 32, rpw == 1
        for (int o = 0; lc + o < owv; o += 32) {  // src / dst already point at column lc
          VecT val;
#pragma unroll
          for (int j = 0; j < V; ++j) val.v[j] = src[o * (sw * V) + j * sw];
          store_stream<T, V>(dst + o, val);
        }
        src += src_step;
        dst += dst_step;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Whole-plane variant.  When the zero-padded, gcd-compacted N-d window of one (b, c) plane fits in
// shared memory, a CTA stages it once and then streams every kernel tap as ONE contiguous run of
// P = prod(out) elements:  out[bc, t, j] = plane[src(j) + tap_off(t)], where src(j) (the staged
// offset of output position j for tap 0) is tap-independent and lives in registers for the whole
// tap loop.  The inner loop is one shared load, one add and one coalesced streaming store per
// element and has no per-row or per-run set-up, which is what made the tiled kernel above
// instruction-issue bound on small planes (ncu: 38 M warp instructions for C2).
// ---------------------------------------------------------------------------------------------
#ifndef EINCONV_FLAT_THREADS
#define EINCONV_FLAT_THREADS 256
#endif
#ifndef EINCONV_FLAT_MINB
#define EINCONV_FLAT_MINB 4
#endif
#ifndef EINCONV_FLAT_JPT
#define EINCONV_FLAT_JPT 8
#endif
constexpr int kFlatThreads = EINCONV_FLAT_THREADS;
constexpr int kFlatMinBlocks = EINCONV_FLAT_MINB;  // resident CTAs per SM the register budget allows
constexpr int kFlatJpt = EINCONV_FLAT_JPT;         // outputs per thread per chunk (<= 16)
constexpr int kFlatMaxTaps = 2048;

template <int ND>
struct FlatParams {
  long long xs_b, xs_c;
  int xs[ND];        // element strides of the spatial dims (|.| < 2^31, plane span < 2^31)
  int channels;
  int in[ND], pad[ND], gstep[ND];
  int ext[ND];       // staged extent per dim, compact coordinates: virtual v = jc * gstep - pad
  int so[ND];        // staged offset per unit of o_d : (stride / g) * pitch
  int ko[ND];        // staged offset per unit of k_d : (dil / g) * pitch
  int out[ND], k[ND];
  int P, K, box;
  int tap_chunks, taps_per_chunk;
  int planes_per_cta;
  int pair;  // 16-bit elements, even P, one plane per CTA: two outputs per 32-bit store
  // planes too large for shared memory are cut along the outermost dimension: a CTA stages the window of
  // `tile0` output coordinates of dim 0 (all of the other dims) and emits P_loc = rows * rest outputs per tap
  int n_tiles0, tile0, rest, tile_shift;
  FastDiv tile_div;
  // TMA staging (16-byte aligned contiguous planes): one cp.async.bulk.tensor of rank
  // ND + 1 with element strides gstep (the innermost dimension cannot be strided and is staged as is)
  // replaces the software staging loop; `lead` elements at the start of the staged rows are alignment
  // slack, `tma_c[d]` the start coordinate of dim d for tile 0.
  int use_tma, lead;
  int tma_c[ND];
  // gcd(stride, dilation) > 1 along the innermost dimension: the copy lands rows of `tma_ext` elements,
  // of which every tma_g-th (from `lead`) is kept; the CTA compacts them in place to rows of ext[ND-1]
  // before the tap loop (stride-g shared loads would be g-way bank conflicts for every output).
  // box_t = elements of the staged buffer (== box unless compacting).
  int tma_compact, tma_ext, tma_g, box_t;
  FastDiv extc_div;
  long long n_planes;
  FastDiv ext_div[ND], out_div[ND], k_div[ND], chan_div, chunk_div, p_div;
};

// streaming (evict-first) store; st.global / .cg / .wt were measured and make no difference here
template <typename T>
__device__ __forceinline__ void store_stream1(T* dst, T val) {
  using U = std::conditional_t<sizeof(T) == 8, unsigned long long,
                               std::conditional_t<sizeof(T) == 4, unsigned int, unsigned short>>;
  __stcs(reinterpret_cast<U*>(dst), (U)val);
}

// One chunk of NI x 256 consecutive outputs for every tap of the CTA: slots 0..NI-2 are valid for
// all threads, slot NI-1 only where `last_on` (ragged end of the plane; its offset is 0 otherwise).
// shared-memory load from a 32-bit shared-window byte address
template <typename T>
__device__ __forceinline__ T lds_addr(uint32_t addr) {
  if constexpr (sizeof(T) == 8) {
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return (T)v;
  } else if constexpr (sizeof(T) == 4) {
    unsigned int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return (T)v;
  } else {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (T)v;
  }
}

template <typename T, int NI>
__device__ __forceinline__ void flat_emit(const T* __restrict__ plane, const int* __restrict__ tap_off,
                                          int n_taps, T* __restrict__ dst, int P,
                                          const int (&off)[kFlatJpt], bool last_on) {
  // byte addresses in the shared window, slot offsets folded in once per chunk: the tap loop is one
  // integer add + one shared load + one store per element
  uint32_t a[NI];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(plane));
#pragma unroll
  for (int i = 0; i < NI; ++i) a[i] = base + (uint32_t)off[i] * (uint32_t)sizeof(T);
#pragma unroll 2
  for (int t = 0; t < n_taps; ++t, dst += P) {
    const uint32_t tb = (uint32_t)tap_off[t] * (uint32_t)sizeof(T);
    T v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) v[i] = lds_addr<T>(a[i] + tb);
#pragma unroll
    for (int i = 0; i < NI - 1; ++i) store_stream1<T>(dst + i * kFlatThreads, v[i]);
    if (last_on) store_stream1<T>(dst + (NI - 1) * kFlatThreads, v[NI - 1]);
  }
}

// Several small planes per CTA: a slot is (plane, position); its destination is no longer a fixed
// stride from the thread's first slot, so it is kept per slot like the source offset.
template <typename T, int NI>
__device__ __forceinline__ void flat_emit_multi(const T* __restrict__ plane, const int* __restrict__ tap_off,
                                                int n_taps, T* __restrict__ dst, int P,
                                                const int (&soff)[kFlatJpt], const int (&doff)[kFlatJpt],
                                                bool last_on) {
  uint32_t a[NI];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(plane));
#pragma unroll
  for (int i = 0; i < NI; ++i) a[i] = base + (uint32_t)soff[i] * (uint32_t)sizeof(T);
#pragma unroll 2
  for (int t = 0; t < n_taps; ++t, dst += P) {
    const uint32_t tb = (uint32_t)tap_off[t] * (uint32_t)sizeof(T);
    T v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) v[i] = lds_addr<T>(a[i] + tb);
#pragma unroll
    for (int i = 0; i < NI - 1; ++i) store_stream1<T>(dst + doff[i], v[i]);
    if (last_on) store_stream1<T>(dst + doff[NI - 1], v[NI - 1]);
  }
}

// 16-bit elements, even P: a slot is a PAIR of consecutive outputs, written with one 32-bit store
// (128 bytes per warp instruction instead of 64); the two sources are independent shared loads.
template <int NI>
__device__ __forceinline__ void flat_emit_pair(const uint16_t* __restrict__ plane, const int* __restrict__ tap_off,
                                               int n_taps, uint16_t* __restrict__ dst, int P,
                                               const int (&offa)[kFlatJpt], const int (&offb)[kFlatJpt],
                                               bool last_on) {
  uint32_t aa[NI], ab[NI];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(plane));
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    aa[i] = base + 2u * (uint32_t)offa[i];
    ab[i] = base + 2u * (uint32_t)offb[i];
  }
#pragma unroll 2
  for (int t = 0; t < n_taps; ++t, dst += P) {
    const uint32_t tb = 2u * (uint32_t)tap_off[t];
    uint32_t v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i)
      v[i] = (uint32_t)lds_addr<uint16_t>(aa[i] + tb) | ((uint32_t)lds_addr<uint16_t>(ab[i] + tb) << 16);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
    for (int i = 0; i < NI - 1; ++i) __stcs(d32 + i * kFlatThreads, v[i]);
    if (last_on) __stcs(d32 + (NI - 1) * kFlatThreads, v[NI - 1]);
  }
}

template <typename T, int ND>
__global__ void __launch_bounds__(kFlatThreads, kFlatMinBlocks)
unfold_flat_kernel(const T* __restrict__ x, T* __restrict__ out,
                   const __grid_constant__ FlatParams<ND> p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* plane_buf = reinterpret_cast<T*>(smem_raw);
  // first staged element that maps to window coordinate 0 (after the in-place compaction: element 0)
  const T* plane = plane_buf + (p.tma_compact ? 0 : p.lead);
  int* tap_off = reinterpret_cast<int*>(smem_raw + (((size_t)p.box_t * p.planes_per_cta * sizeof(T) + 15) & ~(size_t)15));
  uint64_t* tma_bar = reinterpret_cast<uint64_t*>(tap_off + ((p.taps_per_chunk + 1) & ~1));
  const int tid = threadIdx.x;

  uint32_t bcg, chunk, tile = 0;
  p.chunk_div.divmod(blockIdx.x, bcg, chunk);
  if (p.n_tiles0 > 1) {
    uint32_t q;
    p.tile_div.divmod(bcg, q, tile);
    bcg = q;
  }
  // this CTA's share of the plane: P_loc outputs per tap, starting at dst_tile_off
  const int rows0 = min(p.tile0, p.out[0] - (int)tile * p.tile0);
  const int P_loc = rows0 * p.rest;
  const int dst_tile_off = (int)tile * p.tile0 * p.rest;
  const int cc_shift0 = (int)tile * p.tile_shift;
  const int t0 = (int)chunk * p.taps_per_chunk;
  const int t1 = min(p.K, t0 + p.taps_per_chunk);
  const int n_taps = t1 - t0;

  if (p.use_tma) {
    // the whole window of this CTA in one bulk tensor copy; it lands while the threads build their tables
    if (tid == 0) {
      ptx::mbar_init(tma_bar, 1);
      ptx::fence_barrier_init();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      ptx::mbar_expect_tx(tma_bar, (uint32_t)(p.box_t * (int)sizeof(T)));
      const int c0 = p.tma_c[0] + cc_shift0 * p.gstep[0];
      const uint32_t dst = ptx::smem_u32(plane_buf), bar = ptx::smem_u32(tma_bar);
      const uint64_t map = reinterpret_cast<uint64_t>(&tmap);
      const int bcp = (int)bcg;
      if constexpr (ND == 1)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                     "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(bcp) : "memory");
      else if constexpr (ND == 2)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                     "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(p.tma_c[1]), "r"(c0),
                     "r"(bcp) : "memory");
      else if constexpr (ND == 3)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                     "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(p.tma_c[2]),
                     "r"(p.tma_c[1]), "r"(c0), "r"(bcp) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                     "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(p.tma_c[3]),
                     "r"(p.tma_c[2]), "r"(p.tma_c[1]), "r"(c0), "r"(bcp) : "memory");
    }
  }

  // tap table: staged offset of tap t relative to tap 0
  for (int t = t0 + tid; t < t1; t += kFlatThreads) {
    uint32_t rem = (uint32_t)t;
    int off = 0;
#pragma unroll
    for (int d = ND - 1; d >= 0; --d) {
      uint32_t q, kd;
      p.k_div[d].divmod(rem, q, kd);
      rem = q;
      off += (int)kd * p.ko[d];
    }
    tap_off[t - t0] = off;
  }

  constexpr int kChunk = kFlatThreads * kFlatJpt;
  // staged offsets of this thread's output slots for the chunk starting at jb (only the used slots)
  auto slot_offsets = [&](int jb, int (&off)[kFlatJpt], int& n_i, bool& last_on) {
    const int left = P_loc - jb;
    n_i = min(kFlatJpt, (left + kFlatThreads - 1) / kFlatThreads);
    last_on = (n_i - 1) * kFlatThreads + tid < left;
#pragma unroll
    for (int i = 0; i < kFlatJpt; ++i) {
      off[i] = 0;
      if (i < n_i) {
        const int j = jb + i * kFlatThreads + tid;
        uint32_t rem = (uint32_t)(j < P_loc ? j : 0);
        int o = 0;
#pragma unroll
        for (int d = ND - 1; d >= 0; --d) {
          uint32_t q, od;
          p.out_div[d].divmod(rem, q, od);
          rem = q;
          o += (int)od * p.so[d];
        }
        off[i] = o;
      }
    }
  };
  // slots 0 .. n_i-2 are valid for every thread, slot n_i-1 only where last_on
  auto emit_chunk = [&](T* dst, int n_i, const int (&off)[kFlatJpt], bool last_on) {
    switch (n_i) {
#define EINCONV_FLAT_CASE(NI)                                                         \
  case NI:                                                                            \
    if constexpr (NI <= kFlatJpt)                                                     \
      flat_emit<T, NI>(plane, tap_off, n_taps, dst, p.P, off, last_on);               \
    break;
      EINCONV_FLAT_CASE(16)
      EINCONV_FLAT_CASE(15)
      EINCONV_FLAT_CASE(14)
      EINCONV_FLAT_CASE(13)
      EINCONV_FLAT_CASE(12)
      EINCONV_FLAT_CASE(11)
      EINCONV_FLAT_CASE(10)
      EINCONV_FLAT_CASE(9)
      EINCONV_FLAT_CASE(8)
      EINCONV_FLAT_CASE(7)
      EINCONV_FLAT_CASE(6)
      EINCONV_FLAT_CASE(5)
      EINCONV_FLAT_CASE(4)
      EINCONV_FLAT_CASE(3)
      EINCONV_FLAT_CASE(2)
#undef EINCONV_FLAT_CASE
      default: flat_emit<T, 1>(plane, tap_off, n_taps, dst, p.P, off, last_on); break;
    }
  };

  // small planes (one chunk): the slot offsets are the same for every plane of this CTA
  const bool single = P_loc <= kChunk;
  int off0[kFlatJpt], n_i0 = 0;
  bool last_on0 = false;
  if (single && p.planes_per_cta == 1) slot_offsets(0, off0, n_i0, last_on0);

  // staging walk: coordinates of element `tid` and of a stride of kFlatThreads elements
  int cc0[ND], est[ND];
  {
    uint32_t r0 = (uint32_t)tid, r1 = (uint32_t)kFlatThreads;
#pragma unroll
    for (int d = ND - 1; d >= 1; --d) {
      uint32_t q, r;
      p.ext_div[d].divmod(r0, q, r);
      cc0[d] = (int)r;
      r0 = q;
      p.ext_div[d].divmod(r1, q, r);
      est[d] = (int)r;
      r1 = q;
    }
    cc0[0] = (int)r0;  // >= ext[0] only for e >= box, which the loop excludes
    est[0] = (int)r1;
  }

  // stage one plane: plane_dst[c_0..c_{ND-1}] = x[b, c, c_d * gstep_d - pad_d] or 0
  auto stage_plane = [&](long long bc, T* plane_dst) {
    uint32_t b, c;
    p.chan_div.divmod((uint32_t)bc, b, c);
    const T* src = x + (long long)b * p.xs_b + (long long)c * p.xs_c;
    int cc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) cc[d] = cc0[d];
#ifndef EINCONV_FLAT_U
#define EINCONV_FLAT_U 8
#endif
    constexpr int U = EINCONV_FLAT_U;  // independent loads in flight per thread
    for (int e0 = tid; e0 < p.box; e0 += kFlatThreads * U) {
      T vals[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bool ok = e0 + u * kFlatThreads < p.box;
        int off = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const int v = (cc[d] + (d == 0 ? cc_shift0 : 0)) * p.gstep[d] - p.pad[d];
          ok = ok && (unsigned)v < (unsigned)p.in[d];
          off += v * p.xs[d];
        }
        vals[u] = T(0);
        if (ok) vals[u] = __ldg(src + off);
        int carry = 0;
#pragma unroll
        for (int d = ND - 1; d >= 1; --d) {
          const int nc = cc[d] + est[d] + carry;
          carry = nc >= p.ext[d] ? 1 : 0;
          cc[d] = nc - (carry ? p.ext[d] : 0);
        }
        cc[0] += est[0] + carry;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (e0 + u * kFlatThreads < p.box) plane_dst[e0 + u * kFlatThreads] = vals[u];
    }
  };

  if (p.planes_per_cta > 1) {
    // ---- several tiny planes per CTA (e.g. 14x14): all of them are staged at once and their outputs
    // emitted as one pool of (plane, position) slots, so that a CTA pays one load latency and one
    // barrier for tens of kilobytes of output instead of one per 7 KB plane
    const long long bc0 = (long long)bcg * p.planes_per_cta;
    const int n_pl = (int)min((long long)p.planes_per_cta, p.n_planes - bc0);
    for (int pl = 0; pl < n_pl; ++pl) stage_plane(bc0 + pl, plane_buf + pl * p.box);
    __syncthreads();
    const int total = n_pl * p.P;  // <= kFlatThreads * kFlatJpt by construction on the host
    const int n_i = (total + kFlatThreads - 1) / kFlatThreads;
    const bool last_on = (n_i - 1) * kFlatThreads + tid < total;
    int soff[kFlatJpt], doff[kFlatJpt];
    const int plane_out = p.K * p.P;
#pragma unroll
    for (int i = 0; i < kFlatJpt; ++i) {
      soff[i] = 0;
      doff[i] = 0;
      const int sl = i * kFlatThreads + tid;
      if (i < n_i && sl < total) {
        uint32_t pl, jj;
        p.p_div.divmod((uint32_t)sl, pl, jj);
        uint32_t rem = jj;
        int o = 0;
#pragma unroll
        for (int d = ND - 1; d >= 0; --d) {
          uint32_t q, od;
          p.out_div[d].divmod(rem, q, od);
          rem = q;
          o += (int)od * p.so[d];
        }
        soff[i] = (int)pl * p.box + o;
        doff[i] = (int)pl * plane_out + (int)jj;
      }
    }
    T* dst = out + (bc0 * p.K + t0) * (long long)p.P;
    switch (n_i) {
#define EINCONV_FLAT_CASE(NI)                                                              \
  case NI:                                                                                 \
    if constexpr (NI <= kFlatJpt)                                                          \
      flat_emit_multi<T, NI>(plane, tap_off, n_taps, dst, p.P, soff, doff, last_on);       \
    break;
      EINCONV_FLAT_CASE(16)
      EINCONV_FLAT_CASE(15)
      EINCONV_FLAT_CASE(14)
      EINCONV_FLAT_CASE(13)
      EINCONV_FLAT_CASE(12)
      EINCONV_FLAT_CASE(11)
      EINCONV_FLAT_CASE(10)
      EINCONV_FLAT_CASE(9)
      EINCONV_FLAT_CASE(8)
      EINCONV_FLAT_CASE(7)
      EINCONV_FLAT_CASE(6)
      EINCONV_FLAT_CASE(5)
      EINCONV_FLAT_CASE(4)
      EINCONV_FLAT_CASE(3)
      EINCONV_FLAT_CASE(2)
#undef EINCONV_FLAT_CASE
      default: flat_emit_multi<T, 1>(plane, tap_off, n_taps, dst, p.P, soff, doff, last_on); break;
    }
    return;
  }

  // ---- one plane per CTA
  const long long bc = bcg;
  bool pairs = false;
  if constexpr (sizeof(T) == 2) pairs = p.pair != 0;
  // slot offsets of the first chunk
  int off_first[kFlatJpt], n_i_first = n_i0;
  bool last_on_first = last_on0;
  auto first_offsets = [&]() {
    if (single) {
#pragma unroll
      for (int i = 0; i < kFlatJpt; ++i) off_first[i] = off0[i];
    } else {
      slot_offsets(0, off_first, n_i_first, last_on_first);
    }
  };
  if (p.use_tma) {
    if (!pairs) first_offsets();  // computed while the copy is in flight
    __syncthreads();              // tap table written, barrier initialised
    ptx::mbar_wait(tma_bar, 0);
    if (p.tma_compact) {
      // in place, in batches of 16 elements per thread held in registers: the destination index of an
      // element never exceeds its source index, so batches in increasing order never overwrite unread data
      constexpr int CU = 16;
      const int dr = kFlatThreads / p.ext[ND - 1], dc = kFlatThreads - dr * p.ext[ND - 1];
      for (int e0 = 0; e0 < p.box; e0 += kFlatThreads * CU) {
        uint32_t row, col;
        p.extc_div.divmod((uint32_t)(e0 + tid), row, col);
        T v[CU];
#pragma unroll
        for (int u = 0; u < CU; ++u) {
          if (e0 + tid + u * kFlatThreads < p.box) v[u] = plane_buf[(int)row * p.tma_ext + p.lead + (int)col * p.tma_g];
          col += dc;
          row += dr;
          if ((int)col >= p.ext[ND - 1]) {
            col -= p.ext[ND - 1];
            row += 1;
          }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CU; ++u)
          if (e0 + tid + u * kFlatThreads < p.box) plane_buf[e0 + tid + u * kFlatThreads] = v[u];
        __syncthreads();
      }
    }
  } else {
    stage_plane(bc, plane_buf);
    __syncthreads();
    if (!pairs) first_offsets();
  }
  if constexpr (sizeof(T) == 2) {
    if (p.pair) {
      // pairs of consecutive outputs per slot (P even, rows 4-byte aligned)
      const int P2 = P_loc >> 1;
      uint16_t* dst_pairs =
          reinterpret_cast<uint16_t*>(out) + (bc * p.K + t0) * (long long)p.P + dst_tile_off + 2 * tid;
      for (int jb = 0; jb < P2; jb += kChunk) {
        const int left = P2 - jb;
        const int n_i = min(kFlatJpt, (left + kFlatThreads - 1) / kFlatThreads);
        const bool last_on = (n_i - 1) * kFlatThreads + tid < left;
        int offa[kFlatJpt], offb[kFlatJpt];
#pragma unroll
        for (int i = 0; i < kFlatJpt; ++i) {
          offa[i] = 0;
          offb[i] = 0;
          if (i < n_i) {
            const int s2 = jb + i * kFlatThreads + tid;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = 2 * s2 + h;
              uint32_t rem = (uint32_t)(j < P_loc ? j : 0);
              int o = 0;
#pragma unroll
              for (int d = ND - 1; d >= 0; --d) {
                uint32_t q, od;
                p.out_div[d].divmod(rem, q, od);
                rem = q;
                o += (int)od * p.so[d];
              }
              if (h == 0) offa[i] = o; else offb[i] = o;
            }
          }
        }
        const uint16_t* pl16 = reinterpret_cast<const uint16_t*>(plane);
        uint16_t* d = dst_pairs + 2 * jb;
        switch (n_i) {
#define EINCONV_FLAT_CASE(NI)                                                          \
  case NI:                                                                             \
    if constexpr (NI <= kFlatJpt)                                                      \
      flat_emit_pair<NI>(pl16, tap_off, n_taps, d, p.P, offa, offb, last_on);          \
    break;
          EINCONV_FLAT_CASE(8)
          EINCONV_FLAT_CASE(7)
          EINCONV_FLAT_CASE(6)
          EINCONV_FLAT_CASE(5)
          EINCONV_FLAT_CASE(4)
          EINCONV_FLAT_CASE(3)
          EINCONV_FLAT_CASE(2)
#undef EINCONV_FLAT_CASE
          default: flat_emit_pair<1>(pl16, tap_off, n_taps, d, p.P, offa, offb, last_on); break;
        }
      }
      return;
    }
  }
  T* dst_plane = out + (bc * p.K + t0) * (long long)p.P + dst_tile_off + tid;
  emit_chunk(dst_plane, n_i_first, off_first, last_on_first);
  for (int jb = kChunk; jb < P_loc; jb += kChunk) {
    slot_offsets(jb, off_first, n_i_first, last_on_first);
    emit_chunk(dst_plane + jb, n_i_first, off_first, last_on_first);
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent, double-buffered variant of the whole-plane kernel (4- and 8-byte elements, planes whose
// outputs fit one chunk of slots).  The one-CTA-per-plane kernel above starts all its CTAs at once, so
// the staging reads of the whole grid happen before the first output byte is written, and its 512-1024
// equal-sized CTAs quantise badly on 148 SMs.  Here the (plane, tap) pairs are one list, cut into equal
// contiguous ranges for 148 x 4 resident CTAs (balance to within one tap); a CTA stages its next plane
// with cp.async (zero-fill for padding) into the second buffer while it emits the taps of the current
// one, so staging overlaps the stores everywhere except at the very start.
// ---------------------------------------------------------------------------------------------
constexpr int kPersJpt = 16;  // slots per thread: P <= 4096 outputs per tap

template <typename T>
__device__ __forceinline__ void cp_async_zfill(T* smem_dst, const T* gmem_src, bool valid) {
  const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const int src_size = valid ? (int)sizeof(T) : 0;  // 0: nothing is read, the destination is zero-filled
  if constexpr (sizeof(T) == 4)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_size) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_size) : "memory");
}

template <typename T, int NI>
__device__ __forceinline__ void pers_emit(const T* __restrict__ plane, const int* __restrict__ tap_off, int n_taps,
                                          T* __restrict__ dst, int P, const int (&off)[kPersJpt], bool last_on) {
#pragma unroll 2
  for (int t = 0; t < n_taps; ++t, dst += P) {
    const T* s = plane + tap_off[t];
    T v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) v[i] = s[off[i]];
#pragma unroll
    for (int i = 0; i < NI - 1; ++i) store_stream1<T>(dst + i * kFlatThreads, v[i]);
    if (last_on) store_stream1<T>(dst + (NI - 1) * kFlatThreads, v[NI - 1]);
  }
}

template <typename T, int ND>
__global__ void __launch_bounds__(kFlatThreads, 4)
unfold_flat_persistent_kernel(const T* __restrict__ x, T* __restrict__ out,
                              const __grid_constant__ FlatParams<ND> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t buf_bytes = ((size_t)p.box * sizeof(T) + 15) & ~(size_t)15;
  auto buf = [&](int i) { return reinterpret_cast<T*>(smem_raw + (size_t)i * buf_bytes); };
  int* tap_off = reinterpret_cast<int*>(smem_raw + 2 * buf_bytes);
  const int tid = threadIdx.x;

  // my contiguous range of (plane, tap) items
  const long long total = p.n_planes * p.K;
  const long long beg = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
  if (beg >= end) return;

  for (int t = tid; t < p.K; t += kFlatThreads) {
    uint32_t rem = (uint32_t)t;
    int off = 0;
#pragma unroll
    for (int d = ND - 1; d >= 0; --d) {
      uint32_t q, kd;
      p.k_div[d].divmod(rem, q, kd);
      rem = q;
      off += (int)kd * p.ko[d];
    }
    tap_off[t] = off;
  }

  // slot offsets (one chunk: P <= kFlatThreads * kPersJpt), shared by every plane and tap
  const int n_i = (p.P + kFlatThreads - 1) / kFlatThreads;
  const bool last_on = (n_i - 1) * kFlatThreads + tid < p.P;
  int off[kPersJpt];
#pragma unroll
  for (int i = 0; i < kPersJpt; ++i) {
    off[i] = 0;
    if (i < n_i) {
      const int j = i * kFlatThreads + tid;
      uint32_t rem = (uint32_t)(j < p.P ? j : 0);
      int o = 0;
#pragma unroll
      for (int d = ND - 1; d >= 0; --d) {
        uint32_t q, od;
        p.out_div[d].divmod(rem, q, od);
        rem = q;
        o += (int)od * p.so[d];
      }
      off[i] = o;
    }
  }

  // staging walk: coordinates of element `tid` and of a stride of kFlatThreads elements
  int cc0[ND], est[ND];
  {
    uint32_t r0 = (uint32_t)tid, r1 = (uint32_t)kFlatThreads;
#pragma unroll
    for (int d = ND - 1; d >= 1; --d) {
      uint32_t q, r;
      p.ext_div[d].divmod(r0, q, r);
      cc0[d] = (int)r;
      r0 = q;
      p.ext_div[d].divmod(r1, q, r);
      est[d] = (int)r;
      r1 = q;
    }
    cc0[0] = (int)r0;
    est[0] = (int)r1;
  }
  auto prefetch_plane = [&](long long bc, T* dst) {
    uint32_t b, c;
    p.chan_div.divmod((uint32_t)bc, b, c);
    const T* src = x + (long long)b * p.xs_b + (long long)c * p.xs_c;
    int cc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) cc[d] = cc0[d];
    for (int e = tid; e < p.box; e += kFlatThreads) {
      bool ok = true;
      int o = 0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int v = cc[d] * p.gstep[d] - p.pad[d];
        ok = ok && (unsigned)v < (unsigned)p.in[d];
        o += v * p.xs[d];
      }
      cp_async_zfill<T>(dst + e, src + (ok ? o : 0), ok);
      int carry = 0;
#pragma unroll
      for (int d = ND - 1; d >= 1; --d) {
        const int nc = cc[d] + est[d] + carry;
        carry = nc >= p.ext[d] ? 1 : 0;
        cc[d] = nc - (carry ? p.ext[d] : 0);
      }
      cc[0] += est[0] + carry;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const long long p_first = beg / p.K, p_last = (end - 1) / p.K;
  int cur = 0;
  prefetch_plane(p_first, buf(0));
  for (long long pl = p_first; pl <= p_last; ++pl) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // bufs[cur] is complete; everybody has finished reading bufs[cur ^ 1] (and tap_off is written)
    if (pl < p_last) prefetch_plane(pl + 1, buf(cur ^ 1));
    const long long lo = max(beg, pl * p.K), hi = min(end, (pl + 1) * p.K);
    const int t_lo = (int)(lo - pl * p.K), n_taps = (int)(hi - lo);
    T* dst = out + lo * (long long)p.P + tid;
    const T* plane = buf(cur);
    switch (n_i) {
#define EINCONV_PERS_CASE(NI) \
  case NI: pers_emit<T, NI>(plane, tap_off + t_lo, n_taps, dst, p.P, off, last_on); break;
      EINCONV_PERS_CASE(16)
      EINCONV_PERS_CASE(15)
      EINCONV_PERS_CASE(14)
      EINCONV_PERS_CASE(13)
      EINCONV_PERS_CASE(12)
      EINCONV_PERS_CASE(11)
      EINCONV_PERS_CASE(10)
      EINCONV_PERS_CASE(9)
      EINCONV_PERS_CASE(8)
      EINCONV_PERS_CASE(7)
      EINCONV_PERS_CASE(6)
      EINCONV_PERS_CASE(5)
      EINCONV_PERS_CASE(4)
      EINCONV_PERS_CASE(3)
      EINCONV_PERS_CASE(2)
#undef EINCONV_PERS_CASE
      default: pers_emit<T, 1>(plane, tap_off + t_lo, n_taps, dst, p.P, off, last_on); break;
    }
    cur ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// fold: gx[bc, i..] = sum_k gu[bc, k.., o..(i,k)]   (gather form of the adjoint)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fold_kernel(const T* __restrict__ gu, T* __restrict__ gx,
                                                   const __grid_constant__ DevGeom g,
                                                   long long total) {
  using acc_t = typename Cvt<T>::acc_t;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long bc = idx / g.in_total;
  uint32_t rem = (uint32_t)(idx - bc * g.in_total);
  int i[kMaxDims];
  for (int d = g.n_dims - 1; d >= 0; --d) {
    uint32_t q, r;
    g.in_div[d].divmod(rem, q, r);
    i[d] = (int)r;
    rem = q;
  }
  const T* src = gu + bc * (long long)g.k_total * g.out_total;
  acc_t acc = acc_t(0);
  for (int kf = 0; kf < g.k_total; ++kf) {
    uint32_t krem = (uint32_t)kf;
    bool ok = true;
    int oflat = 0, omul = 1;
    for (int d = g.n_dims - 1; d >= 0; --d) {
      uint32_t q, kd;
      g.k_div[d].divmod(krem, q, kd);
      krem = q;
      const int t = i[d] + g.pad[d] - g.dil[d] * (int)kd;
      if (t < 0 || t % g.stride[d] != 0 || t / g.stride[d] >= g.out[d]) ok = false;
      else {
        oflat += (t / g.stride[d]) * omul;
        omul *= g.out[d];
      }
    }
    if (ok) acc += Cvt<T>::load(src + (long long)kf * g.out_total + oflat);
  }
  Cvt<T>::store(gx + idx, acc);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
static int launch_unfold_t(const void* x, void* out, const UnfoldParams& p, size_t smem,
                           unsigned grid, cudaStream_t stream) {
  auto kern = unfold_plane_kernel<T, V>;
  if (smem > 32 * 1024) {  // static tables + dynamic staging may exceed the 48 KB default
    cudaError_t err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(unfold)");
  }
  kern<<<grid, 256, smem, stream>>>((const T*)x, (T*)out, p);
  EINCONV_CHECK_LAUNCH("unfold_plane_kernel");
  return EINCONV_OK;
}

// EINCONV_UNFOLD_TILED=1 forces the tiled kernel (tests exercise both variants on every geometry)
static bool unfold_force_tiled() {
  const char* e = getenv("EINCONV_UNFOLD_TILED");
  return e && e[0] == '1';
}

// Whole-plane kernel: returns 1 if it took the job, 0 if the geometry does not qualify, < 0 on error.
template <typename T, int ND, bool ALLOW_TMA = true>
static int launch_unfold_flat_t(const void* x, void* out, const DevGeom& g0, const long long* xs,
                                int C, cudaStream_t stream) {
  FlatParams<ND> p{};
  const int N0 = g0.n_dims, lead = ND - N0;
  p.xs_b = xs[0];
  p.xs_c = xs[1];
  p.channels = C;
  int stride[ND], dil[ND];
  for (int d = 0; d < ND; ++d) {
    const bool real = d >= lead;
    const int s = d - lead;
    p.in[d] = real ? g0.in[s] : 1;
    p.k[d] = real ? g0.k[s] : 1;
    p.out[d] = real ? g0.out[s] : 1;
    p.pad[d] = real ? g0.pad[s] : 0;
    stride[d] = real ? g0.stride[s] : 1;
    dil[d] = real ? g0.dil[s] : 1;
    const long long xsd = real ? xs[s + 2] : 0;
    if (xsd <= -(1ll << 31) || xsd >= (1ll << 31)) return 0;
    p.xs[d] = (int)xsd;
    p.gstep[d] = igcd(stride[d], dil[d]);
  }
  // TMA staging: planes contiguous and 16-byte aligned in every outer stride, window
  // extents within the 256-element box limit, and planes big enough that neither the persistent nor the
  // pooled small-plane mode applies.  Costs: the innermost dimension is staged uncompacted.
  bool tma = false;
  constexpr long long ES = (long long)sizeof(T);
  constexpr int A16 = 16 / (int)sizeof(T);  // elements per 16 bytes
  if constexpr (ND <= 4) {
    tma = ALLOW_TMA && !getenv("EINCONV_UNFOLD_NO_TMA") && ((uintptr_t)x % 16) == 0 && p.xs[ND - 1] == 1 &&
          (p.xs_c * ES) % 16 == 0 && p.xs_b == (long long)C * p.xs_c && p.xs_c > 0 &&
          (long long)g0.k_total * g0.out_total * ES > 64 * 1024 && g0.out_total * 2 > kFlatThreads * kFlatJpt &&
          (long long)g0.batch * C < (1ll << 31);
    for (int d = 0; d < ND - 1 && tma; ++d)
      if (p.in[d] > 1 && (p.xs[d] <= 0 || ((long long)p.xs[d] * ES) % 16 != 0)) tma = false;
  }
  // innermost gcd > 1: the rows stay as copied (stride-g shared loads in the tap loop).  Compacting them in
  // shared memory after the copy (EINCONV_UNFOLD_COMPACT=1) removes the g-way bank conflicts but was
  // measured neutral on C2 (49.8 vs 48.9 us): the tap loop is not bound by the shared-memory pipe.
  const int g_last = p.gstep[ND - 1];
  const bool compact = tma && g_last > 1 && getenv("EINCONV_UNFOLD_COMPACT") != nullptr;
  if (tma && !compact) p.gstep[ND - 1] = 1;
  long long box = 1, box_t = 1, span = 0;
  p.n_tiles0 = 1;
  p.tile0 = std::max(1, p.out[0]);
  p.rest = 1;
  for (int d = 1; d < ND; ++d) p.rest *= std::max(1, p.out[d]);
  for (int d = ND - 1; d >= 0; --d) {
    long long ext = (long long)(stride[d] / p.gstep[d]) * (p.out[d] - 1) +
                    (long long)(dil[d] / p.gstep[d]) * (p.k[d] - 1) + 1;
    long long ext_t = ext;  // staged extent of this dim in the copy's layout
    if (tma && d == ND - 1) {
      // innermost start coordinate must be 16-byte aligned: start `lead` elements early, whole 16-byte rows
      const int start = -(((p.pad[d] + A16 - 1) / A16) * A16);
      p.lead = -p.pad[d] - start;
      p.tma_c[d] = start;
      if (compact) {
        ext_t = ((ext - 1) * g_last + 1 + p.lead + A16 - 1) / A16 * A16;
        p.tma_ext = (int)ext_t;
        p.tma_g = g_last;
      } else {
        ext = (ext + p.lead + A16 - 1) / A16 * A16;
        ext_t = ext;
      }
    } else if (tma) {
      p.tma_c[d] = -p.pad[d];
    }
    if (d == 0 && !getenv("EINCONV_UNFOLD_NO_TILE0")) {
      // cut along the outermost dimension so that the staged window fits ~40 KB (one plane of the
      // uncompacted TMA layout may take up to 50 KB: still four CTAs per SM)
      const long long budget = ((tma ? 50 : 40) * 1024) / (long long)sizeof(T);
      const long long sg = stride[0] / p.gstep[0], halo = (long long)(dil[0] / p.gstep[0]) * (p.k[0] - 1) + 1;
      if (ext * box_t > budget) {
        const long long rows = (budget / std::max<long long>(1, box_t) - halo) / sg + 1;  // ext(tile0) <= budget / box
        if (rows < 1) return 0;
        p.tile0 = (int)std::min<long long>(rows, p.out[0]);
        p.n_tiles0 = ceil_div(p.out[0], p.tile0);
        p.tile0 = ceil_div(p.out[0], p.n_tiles0);  // even split
        ext = sg * (p.tile0 - 1) + halo;
        ext_t = ext;
      }
      p.tile_shift = (int)(sg * p.tile0);
    }
    if (ext_t * box_t > (1 << 20)) return 0;
    p.ext[d] = (int)ext;
    p.so[d] = (int)((stride[d] / p.gstep[d]) * box);
    p.ko[d] = (int)((dil[d] / p.gstep[d]) * box);
    box *= ext;
    box_t *= ext_t;
    span += (long long)(p.in[d] - 1) * (p.xs[d] < 0 ? -(long long)p.xs[d] : (long long)p.xs[d]);
  }
  if (span >= (1ll << 31)) return 0;
  if ((long long)p.tile0 * p.rest >= (1ll << 30)) return 0;
  if (tma) {
    bool fits = (compact ? p.tma_ext : p.ext[ND - 1]) <= 256;
    for (int d = 0; d < ND - 1; ++d) fits = fits && (long long)p.ext[d] * p.gstep[d] <= 256 && p.gstep[d] <= 8;
    if (!fits) return launch_unfold_flat_t<T, ND, false>(x, out, g0, xs, C, stream);
  }
  p.use_tma = tma ? 1 : 0;
  p.tma_compact = compact ? 1 : 0;
  p.extc_div = FastDiv((uint32_t)std::max(1, p.ext[ND - 1]));
  p.tile_div = FastDiv((uint32_t)p.n_tiles0);
  p.box = (int)box;
  p.box_t = (int)box_t;
  p.P = g0.out_total;
  p.K = g0.k_total;
  const long long n_planes = (long long)g0.batch * C;

  for (int d = 0; d < ND; ++d) {
    p.ext_div[d] = FastDiv((uint32_t)p.ext[d]);
    p.out_div[d] = FastDiv((uint32_t)p.out[d]);
    p.k_div[d] = FastDiv((uint32_t)p.k[d]);
  }
  p.chan_div = FastDiv((uint32_t)C);
  p.n_planes = n_planes;

  // persistent double-buffered variant: 4/8-byte elements, one chunk of slots, two staged windows in
  // shared memory at 4 CTAs per SM, enough (plane, tap) items to balance, and 4/8-byte aligned sources
  if constexpr (sizeof(T) >= 4) {
    const size_t buf_bytes = ((size_t)p.box * sizeof(T) + 15) & ~(size_t)15;
    const size_t smem_p = 2 * buf_bytes + (size_t)p.K * 4;
    const long long items = n_planes * p.K;
    const int grid_p = 4 * num_sms();
    bool aligned = ((uintptr_t)x % sizeof(T)) == 0;
    // measured: wins for small planes (14x14: 33 -> 27 us for 58 MB), loses ~5% on C2-sized planes, where
    // one CTA per plane already amortises its set-up (EINCONV_UNFOLD_PERSISTENT=1 forces it)
    const bool small_plane = (long long)p.K * p.P * (long long)sizeof(T) <= 64 * 1024;
    if (!tma && p.n_tiles0 == 1 && !getenv("EINCONV_UNFOLD_NO_PERSISTENT") &&
        (small_plane || getenv("EINCONV_UNFOLD_PERSISTENT")) && aligned &&
        p.P > 0 && p.P <= kFlatThreads * kPersJpt &&
        smem_p <= 56 * 1024 && items >= 8ll * grid_p && p.K <= kFlatMaxTaps && n_planes < (1ll << 31)) {
      auto kern = unfold_flat_persistent_kernel<T, ND>;
      if (smem_p > 32 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(unfold_flat_persistent)");
      }
      kern<<<grid_p, kFlatThreads, smem_p, stream>>>((const T*)x, (T*)out, p);
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess) return cuda_fail(err, "unfold_flat_persistent_kernel");
      count_launch();
      return 1;
    }
  }

  // CTAs per plane: split the taps while there are fewer CTAs than ~4 per SM
  int chunks = 1;
  while (n_planes * p.n_tiles0 * chunks < 4ll * num_sms() && chunks * 2 <= p.K) chunks *= 2;
  if (const char* e = getenv("EINCONV_UNFOLD_CHUNKS")) chunks = std::max(1, std::min(p.K, atoi(e)));
  p.taps_per_chunk = ceil_div(p.K, chunks);
  p.tap_chunks = ceil_div(p.K, p.taps_per_chunk);
  if (p.taps_per_chunk > kFlatMaxTaps) return 0;
  size_t smem = (((size_t)p.box_t * sizeof(T) + 15) & ~(size_t)15) + (size_t)p.taps_per_chunk * 4 + 16;
  if (smem > (size_t)(tma ? 52 : 48) * 1024) {  // keep >= 4 CTAs per SM resident
    if (tma) return launch_unfold_flat_t<T, ND, false>(x, out, g0, xs, C, stream);
    return 0;
  }
  // planes per CTA: small planes only; all their (plane, position) slots must fit the thread's slot
  // registers and their staged windows the shared-memory budget; aim at >= 64 KB of output per CTA
  // while keeping at least ~8 CTAs per SM in the grid
  p.n_planes = n_planes;
  p.planes_per_cta = 1;
  p.p_div = FastDiv((uint32_t)std::max(1, p.P));
  if (p.n_tiles0 == 1 && p.P > 0 && p.P * 2 <= kFlatThreads * kFlatJpt && !getenv("EINCONV_UNFOLD_ONE_PLANE")) {
    const long long per_plane_bytes = (long long)p.taps_per_chunk * p.P * (long long)sizeof(T);
    long long pc = std::max<long long>(1, std::min<long long>(32, (64 * 1024 + per_plane_bytes - 1) / per_plane_bytes));
    pc = std::min<long long>(pc, (kFlatThreads * kFlatJpt) / p.P);
    pc = std::min<long long>(pc, (long long)(40 * 1024) / std::max<long long>(1, (long long)p.box * (long long)sizeof(T)));
    while (pc > 1 && ((n_planes + pc - 1) / pc) * p.tap_chunks < 8ll * num_sms()) pc /= 2;
    p.planes_per_cta = (int)std::max<long long>(1, pc);
  }
  p.pair = (sizeof(T) == 2 && p.planes_per_cta == 1 && p.P % 2 == 0 && (p.n_tiles0 == 1 || p.rest % 2 == 0) &&
            ((uintptr_t)out % 4) == 0 &&
            kFlatJpt <= 8 && !getenv("EINCONV_UNFOLD_NO_PAIRS"))
               ? 1
               : 0;
  smem = (((size_t)p.box_t * p.planes_per_cta * sizeof(T) + 15) & ~(size_t)15) + (size_t)p.taps_per_chunk * 4 + 16;
  const long long grid = ((n_planes + p.planes_per_cta - 1) / p.planes_per_cta) * p.n_tiles0 * p.tap_chunks;
  if (grid >= (1ll << 31)) return 0;
  // Too little output per staged element: the tiled kernel amortises its staging better.
  for (int d = 0; d < ND; ++d) {
    p.ext_div[d] = FastDiv((uint32_t)p.ext[d]);
    p.out_div[d] = FastDiv((uint32_t)p.out[d]);
    p.k_div[d] = FastDiv((uint32_t)p.k[d]);
  }
  p.chan_div = FastDiv((uint32_t)C);
  p.chunk_div = FastDiv((uint32_t)p.tap_chunks);

  auto kern = unfold_flat_kernel<T, ND>;
  if (smem > 32 * 1024) {
    cudaError_t err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(unfold_flat)");
  }
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if constexpr (ND <= 4) if (tma) {
    unsigned long long dims[5], strides[4];
    unsigned box_dims[5], est[5];
    for (int i = 0; i < ND; ++i) {
      const int d = ND - 1 - i;
      dims[i] = (unsigned long long)p.in[d];
      box_dims[i] = (unsigned)(i == 0 ? (compact ? p.tma_ext : p.ext[d]) : p.ext[d] * p.gstep[d]);
      est[i] = (unsigned)(i == 0 ? 1 : p.gstep[d]);
      // byte stride of dim d (dims of extent 1 take any valid stride)
      if (i > 0) strides[i - 1] = (unsigned long long)(p.in[d] > 1 ? (long long)p.xs[d] : p.xs_c) * ES;
    }
    dims[ND] = (unsigned long long)n_planes;
    box_dims[ND] = 1;
    est[ND] = 1;
    strides[ND - 1] = (unsigned long long)(p.xs_c * ES);
    // a map the driver rejects is not an error of the call: the software staging loop handles every layout
    if (ctma::encode_tiled_raw(&tmap, x, (int)sizeof(T), ND + 1, dims, strides, box_dims, est) != EINCONV_OK)
      return launch_unfold_flat_t<T, ND, false>(x, out, g0, xs, C, stream);
  }
  kern<<<(unsigned)grid, kFlatThreads, smem, stream>>>((const T*)x, (T*)out, p, tmap);
  {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "unfold_flat_kernel");
    count_launch();
  }
  return 1;
}

template <typename T>
static int launch_unfold_flat(const void* x, void* out, const DevGeom& g0, const long long* xs,
                              int C, cudaStream_t stream) {
  switch (g0.n_dims) {
    case 1: return launch_unfold_flat_t<T, 1>(x, out, g0, xs, C, stream);
    case 2: return launch_unfold_flat_t<T, 2>(x, out, g0, xs, C, stream);
    case 3: return launch_unfold_flat_t<T, 3>(x, out, g0, xs, C, stream);
    case 4: return launch_unfold_flat_t<T, 4>(x, out, g0, xs, C, stream);
    default: return launch_unfold_flat_t<T, kMaxDims>(x, out, g0, xs, C, stream);
  }
}

int launch_unfold(const void* x, const int64_t* x_strides, void* out, int dtype,
                  const einconv_geom* geom, cudaStream_t stream) {
  DevGeom g0;
  int st = make_dev_geom(geom, &g0, /*need_c_out=*/false);
  if (st) return st;
  const int64_t esz = elem_size(dtype);
  EINCONV_CHECK_ARG(esz > 0, "unfold: unknown dtype %d", dtype);

  UnfoldParams p{};
  const int N0 = g0.n_dims;
  const int C = g0.groups * g0.c_in_g;
  p.channels = C;

  // contiguous default strides
  long long cs[kMaxDims + 2];
  {
    long long s = 1;
    for (int d = N0 - 1; d >= 0; --d) {
      cs[d + 2] = s;
      s *= g0.in[d];
    }
    cs[1] = s;
    cs[0] = s * C;
  }
  const long long* xs = x_strides ? (const long long*)x_strides : cs;
  p.xs_b = xs[0];
  p.xs_c = xs[1];

  if ((long long)g0.batch * C * g0.k_total * (long long)g0.out_total == 0) return EINCONV_OK;
  if (!unfold_force_tiled()) {
    int took;
    switch (esz) {
      case 2: took = launch_unfold_flat<uint16_t>(x, out, g0, xs, C, stream); break;
      case 4: took = launch_unfold_flat<uint32_t>(x, out, g0, xs, C, stream); break;
      default: took = launch_unfold_flat<unsigned long long>(x, out, g0, xs, C, stream); break;
    }
    if (took != 0) return took < 0 ? took : EINCONV_OK;
  }

  // normalise to >= 2 spatial dims by prepending a unit dimension
  const int shift = (N0 == 1) ? 1 : 0;
  EINCONV_CHECK_ARG(N0 + shift <= kMaxDims, "unfold: too many dimensions (%d)", N0);
  DevGeom& g = p.g;
  g = g0;
  g.n_dims = N0 + shift;
  if (shift) {
    g.in[0] = 1; g.k[0] = 1; g.stride[0] = 1; g.dil[0] = 1; g.pad[0] = 0; g.out[0] = 1;
    p.xs[0] = 0;
    for (int d = 0; d < N0; ++d) {
      g.in[d + 1] = g0.in[d]; g.k[d + 1] = g0.k[d]; g.stride[d + 1] = g0.stride[d];
      g.dil[d + 1] = g0.dil[d]; g.pad[d + 1] = g0.pad[d]; g.out[d + 1] = g0.out[d];
      p.xs[d + 1] = xs[d + 2];
    }
  } else {
    for (int d = 0; d < N0; ++d) p.xs[d] = xs[d + 2];
  }
  const int N = g.n_dims, H = N - 2, W = N - 1;

  const long long out_elems = (long long)g.batch * C * g.k_total * (long long)g.out_total;
  if (out_elems == 0) return EINCONV_OK;  // empty output: nothing to write

  p.n_outer = N - 2;
  long long v_total = 1, k_outer = 1;
  for (int d = 0; d < p.n_outer; ++d) {
    // virtual coordinate v (input frame shifted by the padding) is touched by a tap only if
    // gcd(stride, dilation) | v, so only those planes get a CTA
    const int vrange = g.stride[d] * (g.out[d] - 1) + g.dil[d] * (g.k[d] - 1) + 1;
    p.vstep[d] = igcd(g.stride[d], g.dil[d]);
    p.vcount[d] = (vrange - 1) / p.vstep[d] + 1;
    v_total *= p.vcount[d];
    k_outer *= g.k[d];
  }
  EINCONV_CHECK_ARG(v_total < (1ll << 30), "unfold: geometry too large");
  EINCONV_CHECK_ARG(p.xs[W] > -(1ll << 31) && p.xs[W] < (1ll << 31), "unfold: innermost stride too large");
  EINCONV_CHECK_ARG(k_outer <= kMaxOuterTaps,
                    "unfold: more than %d taps over the leading dimensions", kMaxOuterTaps);
  p.v_total = (int)v_total;
  p.k_outer_total = (int)k_outer;
  p.k_inner_total = g.k[H] * g.k[W];

  p.gh = igcd(g.stride[H], g.dil[H]);
  p.gw = igcd(g.stride[W], g.dil[W]);
  p.sh = g.stride[H] / p.gh; p.dh = g.dil[H] / p.gh;
  p.sw = g.stride[W] / p.gw; p.dw = g.dil[W] / p.gw;
  p.cols = p.sw * (g.out[W] - 1) + p.dw * (g.k[W] - 1) + 1;
  p.kw_div = FastDiv((uint32_t)g.k[W]);
  p.chan_div = FastDiv((uint32_t)C);
  for (int d = 0; d < p.n_outer; ++d) {
    p.vc_div[d] = FastDiv((uint32_t)p.vcount[d]);
    p.ko_div[d] = FastDiv((uint32_t)g.k[d]);
    p.so_div[d] = FastDiv((uint32_t)g.stride[d]);
  }

  // vector width of the output stores: V | O_w keeps a vector inside one output row; every run
  // starts at a multiple of O_w elements from `out`, which must itself be aligned.
  int V = 1;
  for (int cand : {4, 2}) {
    if (g.out[W] % cand == 0 && cand * esz <= 16 && ((uintptr_t)out % (cand * esz)) == 0) {
      V = cand;
      break;
    }
  }
  p.owv = g.out[W] / V;
  p.lpr_shift = 0;
  while ((1 << p.lpr_shift) < p.owv && p.lpr_shift < 5) ++p.lpr_shift;

  // Work per CTA: ~8-16k output elements per valid outer tap, so that there are many more CTAs
  // than the 148 x 8 resident slots and the tail stays short.  Row tiles only when one plane is
  // larger than that (or its staging window exceeds the shared-memory budget).
  const long long budget = 40 * 1024;
  auto rows_for = [&](int th) { return (long long)p.sh * (th - 1) + (long long)p.dh * (g.k[H] - 1) + 1; };
  const long long per_row_out = (long long)p.k_inner_total * g.out[W];
  int th = g.out[H];
  if (per_row_out * th > 16384) th = (int)std::max<long long>(1, 12288 / per_row_out);
  while (th > 1 && rows_for(th) * p.cols * esz > budget) th = (th + 1) / 2;
  EINCONV_CHECK_ARG(rows_for(th) * p.cols * esz <= 200 * 1024,
                    "unfold: one row tile does not fit in shared memory");
  p.tile_h = th;
  p.n_tiles = ceil_div(g.out[H], th);
  p.rows_cap = (int)rows_for(th);

  long long per_plane_out = per_row_out * th;
  long long per_plane_smem = (long long)p.rows_cap * p.cols * esz;
  long long cb = std::max<long long>(1, 12288 / per_plane_out);
  cb = std::min<long long>(cb, std::max<long long>(1, (40 * 1024) / per_plane_smem));
  cb = std::max<long long>(1, std::min<long long>(cb, (long long)g.batch * C));
  cb = std::min<long long>(cb, 64);
  // Prefer more, smaller CTAs (shorter tail) while there are fewer than ~4 waves of the
  // 148 x 8 resident slots: fewer planes per CTA.
  const long long want_ctas = 4ll * num_sms() * 8;
  auto n_ctas = [&](long long cb_, int th_) {
    return (long long)ceil_div((long long)g.batch * C, cb_) * p.v_total * ceil_div(g.out[H], th_);
  };
  while (cb > 1 && n_ctas(cb, th) < want_ctas) cb = (cb + 1) / 2;
  p.tile_h = th;
  p.n_tiles = ceil_div(g.out[H], th);
  p.rows_cap = (int)rows_for(th);
  per_plane_out = per_row_out * th;
  per_plane_smem = (long long)p.rows_cap * p.cols * esz;
  p.cb = (int)cb;
  const int n_bc_groups = ceil_div((long long)g.batch * C, cb);
  p.kinner_div = FastDiv((uint32_t)p.k_inner_total);
  p.tiles_div = FastDiv((uint32_t)p.n_tiles);
  p.vtotal_div = FastDiv((uint32_t)p.v_total);
  const long long grid = (long long)n_bc_groups * p.v_total * p.n_tiles;
  EINCONV_CHECK_ARG(grid > 0 && grid < (1ll << 31), "unfold: grid too large");
  const size_t smem = (size_t)(cb * per_plane_smem);

#define EINCONV_UNFOLD_CASE(TYPE)                                                              \
  switch (V) {                                                                                 \
    case 4: if constexpr (sizeof(TYPE) * 4 <= 16)                                              \
              return launch_unfold_t<TYPE, 4>(x, out, p, smem, (unsigned)grid, stream);        \
            [[fallthrough]];                                                                   \
    case 2: return launch_unfold_t<TYPE, 2>(x, out, p, smem, (unsigned)grid, stream);          \
    default: return launch_unfold_t<TYPE, 1>(x, out, p, smem, (unsigned)grid, stream);         \
  }
  switch (esz) {
    case 2: EINCONV_UNFOLD_CASE(uint16_t)
    case 4: EINCONV_UNFOLD_CASE(uint32_t)
    default: EINCONV_UNFOLD_CASE(unsigned long long)
  }
#undef EINCONV_UNFOLD_CASE
}

int launch_fold(const void* gu, void* gx, int dtype, const einconv_geom* geom,
                cudaStream_t stream) {
  DevGeom g;
  int st = make_dev_geom(geom, &g, /*need_c_out=*/false);
  if (st) return st;
  const long long total = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
  if (total == 0) return EINCONV_OK;
  const long long blocks = (total + 255) / 256;
  EINCONV_CHECK_ARG(blocks < (1ll << 31), "fold: tensor too large");
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    fold_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>((const T*)gu, (T*)gx, g, total);
    EINCONV_CHECK_LAUNCH("fold_kernel");
    return EINCONV_OK;
  });
}

// ---------------------------------------------------------------------------------------------
// unfoldNd_transpose (expressions/conv_transposeNd_unfold.py:81-91):
//   out[bc, k.., i..] = x[bc, o..] if i_d = s_d o_d + dil_d k_d - p_d for every d, else 0.
// One thread per (bc, i) walks all taps: consecutive threads hold consecutive i, so for every tap the
// warp writes one contiguous run of `out` and (for stride 1) reads one contiguous run of x.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) unfold_transpose_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                               const __grid_constant__ DevGeom g, long long total) {
  // per-CTA table: dil_d * k_d of every tap (kMaxDims bytes... ints per tap), so that the tap loop has
  // no odometer and can be unrolled with its loads issued ahead of the stores
  extern __shared__ int s_tap[];  // [k_total][n_dims]
  for (int e = threadIdx.x; e < g.k_total * g.n_dims; e += blockDim.x) {
    const int kf = e / g.n_dims, d = e - kf * g.n_dims;
    uint32_t rem = (uint32_t)kf, kd = 0;
    for (int dd = g.n_dims - 1; dd >= d; --dd) {
      uint32_t q;
      g.k_div[dd].divmod(rem, q, kd);
      rem = q;
    }
    s_tap[e] = g.dil[d] * (int)kd;
  }
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long bc = idx / g.in_total;
  uint32_t rem = (uint32_t)(idx - bc * g.in_total);
  const uint32_t iflat = rem;
  int ipos[kMaxDims];
  for (int d = g.n_dims - 1; d >= 0; --d) {
    uint32_t q, r;
    g.in_div[d].divmod(rem, q, r);
    ipos[d] = (int)r + g.pad[d];
    rem = q;
  }
  const T* src = x + bc * g.out_total;
  T* dst = out + bc * (long long)g.k_total * g.in_total + iflat;
  auto source = [&](int kf, bool& ok) {
    ok = true;
    int oflat = 0;
    const int* tap = s_tap + kf * g.n_dims;
    for (int d = 0; d < g.n_dims; ++d) {
      const int t = ipos[d] - tap[d];
      uint32_t o, r;
      g.stride_div[d].divmod((uint32_t)max(t, 0), o, r);  // multiply-high: no integer division in the loop
      ok = ok && t >= 0 && r == 0 && (int)o < g.out[d];
      oflat = oflat * g.out[d] + (int)o;
    }
    return oflat;
  };
  constexpr int U = 4;
  int kf = 0;
  for (; kf + U <= g.k_total; kf += U) {
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool ok;
      const int oflat = source(kf + u, ok);
      v[u] = T(0);
      if (ok) v[u] = __ldg(src + oflat);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) store_stream1<T>(dst + (long long)(kf + u) * g.in_total, v[u]);
  }
  for (; kf < g.k_total; ++kf) {
    bool ok;
    const int oflat = source(kf, ok);
    T v = T(0);
    if (ok) v = __ldg(src + oflat);
    store_stream1<T>(dst + (long long)kf * g.in_total, v);
  }
}

// Table-driven variant: per CTA, tab_d[k_d][i_d] = (o_d * prod(O_e, e > d)) or a large negative
// sentinel (one small table per dimension, sum_d K_d * I_d entries), the (b, c) plane of x staged in
// shared memory when it fits.  Per output element the tap loop is then ND table look-ups, one add, one
// sign test, one plane read and one coalesced streaming store, instead of ND divisions.
constexpr int kTransposeSentinel = -(1 << 28);

template <typename T, int ND>
__global__ void __launch_bounds__(256) unfold_transpose_tab_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                                   const __grid_constant__ DevGeom g, int lead,
                                                                   int tab_total, int plane_in_smem) {
  // geometry dims lead .. lead + n_dims - 1 of this kernel's ND-vector are real, leading ones are unit
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* tab = reinterpret_cast<int*>(smem_raw);           // per-dim tables, concatenated
  int* tapk = tab + tab_total;                            // [k_total][ND]: k_d * I_d (row offset into tab_d)
  T* plane = reinterpret_cast<T*>(smem_raw + ((((size_t)tab_total + (size_t)g.k_total * ND) * 4 + 15) & ~(size_t)15));
  const int n = g.n_dims;
  int I[ND], K[ND], base[ND];
  {
    int acc = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const bool real = d >= lead;
      I[d] = real ? g.in[d - lead] : 1;
      K[d] = real ? g.k[d - lead] : 1;
      base[d] = acc;
      acc += I[d] * K[d];
    }
  }
  // tables
  for (int e = threadIdx.x; e < tab_total; e += blockDim.x) {
    int d = 0;
#pragma unroll
    for (int dd = 1; dd < ND; ++dd)
      if (e >= base[dd]) d = dd;
    const int r = e - base[d];
    const int k = r / I[d], i = r - k * I[d];
    int val = 0;
    if (d >= lead) {
      const int gd = d - lead;
      const int t = i + g.pad[gd] - g.dil[gd] * k;
      const int o = t >= 0 ? t / g.stride[gd] : -1;
      if (t < 0 || o * g.stride[gd] != t || o >= g.out[gd]) {
        val = kTransposeSentinel;
      } else {
        int mul = 1;
        for (int e2 = gd + 1; e2 < n; ++e2) mul *= g.out[e2];
        val = o * mul;
      }
    }
    tab[e] = val;
  }
  for (int e = threadIdx.x; e < g.k_total * ND; e += blockDim.x) {
    const int kf = e / ND, d = e - kf * ND;
    int kd = 0;
    if (d >= lead) {
      uint32_t rem = (uint32_t)kf, r = 0;
      for (int dd = n - 1; dd >= d - lead; --dd) {
        uint32_t q;
        g.k_div[dd].divmod(rem, q, r);
        rem = q;
      }
      kd = (int)r;
    }
    tapk[e] = kd * I[d];
  }
  const long long bc = blockIdx.y;
  const T* src = x + bc * g.out_total;
  if (plane_in_smem)
    for (int e = threadIdx.x; e < g.out_total; e += blockDim.x) plane[e] = __ldg(src + e);
  __syncthreads();

  // a CTA walks several chunks of 256 positions so that its tables and staged plane are amortised
  for (int iflat = blockIdx.x * blockDim.x + threadIdx.x; iflat < g.in_total; iflat += gridDim.x * blockDim.x) {
    const int* tp[ND];
    {
      uint32_t rem = (uint32_t)iflat;
#pragma unroll
      for (int d = ND - 1; d >= 0; --d) {
        uint32_t q = 0, r = 0;
        if (d >= lead) {
          g.in_div[d - lead].divmod(rem, q, r);
          rem = q;
        }
        tp[d] = tab + base[d] + (int)r;
      }
    }
    T* dst = out + bc * (long long)g.k_total * g.in_total + iflat;
    constexpr int U = 4;
    int kf = 0;
    for (; kf + U <= g.k_total; kf += U) {
      T v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int* tk = tapk + (kf + u) * ND;
        int off = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) off += tp[d][tk[d]];
        v[u] = T(0);
        if (off >= 0) v[u] = plane_in_smem ? plane[off] : __ldg(src + off);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) store_stream1<T>(dst + (long long)(kf + u) * g.in_total, v[u]);
    }
    for (; kf < g.k_total; ++kf) {
      const int* tk = tapk + kf * ND;
      int off = 0;
#pragma unroll
      for (int d = 0; d < ND; ++d) off += tp[d][tk[d]];
      T v = T(0);
      if (off >= 0) v = plane_in_smem ? plane[off] : __ldg(src + off);
      store_stream1<T>(dst + (long long)kf * g.in_total, v);
    }
  }
}

template <typename T, int ND>
static int launch_unfold_transpose_tab(const void* x, void* out, const DevGeom& g, cudaStream_t stream) {
  const int lead = ND - g.n_dims;
  long long tab_total = 0;
  for (int d = 0; d < g.n_dims; ++d) tab_total += (long long)g.in[d] * g.k[d];
  tab_total += lead;  // unit dims: one entry each
  const long long planes = (long long)g.batch * g.groups * g.c_in_g;
  if (planes >= 65536 || g.out_total >= (1 << 27)) return 0;
  size_t smem = (size_t)(tab_total + (long long)g.k_total * ND) * 4;
  if (smem > 40 * 1024) return 0;
  smem = (smem + 15) & ~(size_t)15;
  int plane_in_smem = 0;
  if ((size_t)g.out_total * sizeof(T) + smem <= 44 * 1024) {
    plane_in_smem = 1;
    smem += (size_t)g.out_total * sizeof(T);
  }
  const long long chunks = (g.in_total + 255) / 256;
  const long long per_plane = std::max<long long>(1, std::min<long long>(chunks, (16ll * num_sms() + planes - 1) / planes));
  dim3 grid((unsigned)per_plane, (unsigned)planes);
  unfold_transpose_tab_kernel<T, ND><<<grid, 256, smem, stream>>>((const T*)x, (T*)out, g, lead, (int)tab_total,
                                                                  plane_in_smem);
  {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "unfold_transpose_tab_kernel");
    count_launch();
  }
  return 1;
}

template <typename T>
static int launch_unfold_transpose_tab_nd(const void* x, void* out, const DevGeom& g, cudaStream_t stream) {
  if (getenv("EINCONV_UNFOLD_TRANSPOSE_SIMPLE")) return 0;
  switch (g.n_dims) {
    case 1: return launch_unfold_transpose_tab<T, 1>(x, out, g, stream);
    case 2: return launch_unfold_transpose_tab<T, 2>(x, out, g, stream);
    case 3: return launch_unfold_transpose_tab<T, 3>(x, out, g, stream);
    default: return 0;  // N > 3: the per-element kernel
  }
}

int launch_unfold_transpose(const void* x, void* out, int dtype, const einconv_geom* geom, cudaStream_t stream) {
  DevGeom g;
  int st = make_dev_geom(geom, &g, /*need_c_out=*/false);
  if (st) return st;
  const int64_t esz = elem_size(dtype);
  EINCONV_CHECK_ARG(esz > 0, "unfold_transpose: unknown dtype %d", dtype);
  const long long total = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
  if (total == 0 || g.k_total == 0) return EINCONV_OK;
  {
    int took;
    switch (esz) {
      case 2: took = launch_unfold_transpose_tab_nd<uint16_t>(x, out, g, stream); break;
      case 4: took = launch_unfold_transpose_tab_nd<uint32_t>(x, out, g, stream); break;
      default: took = launch_unfold_transpose_tab_nd<unsigned long long>(x, out, g, stream); break;
    }
    if (took != 0) return took < 0 ? took : EINCONV_OK;
  }
  const long long blocks = (total + 255) / 256;
  EINCONV_CHECK_ARG(blocks < (1ll << 31), "unfold_transpose: tensor too large");
  const size_t smem = (size_t)g.k_total * g.n_dims * sizeof(int);
  EINCONV_CHECK_ARG(smem <= 40 * 1024, "unfold_transpose: more than %d taps", (int)(40 * 1024 / 4 / g.n_dims));
  switch (esz) {
    case 2:
      unfold_transpose_kernel<uint16_t><<<(unsigned)blocks, 256, smem, stream>>>((const uint16_t*)x, (uint16_t*)out, g, total);
      break;
    case 4:
      unfold_transpose_kernel<uint32_t><<<(unsigned)blocks, 256, smem, stream>>>((const uint32_t*)x, (uint32_t*)out, g, total);
      break;
    default:
      unfold_transpose_kernel<unsigned long long><<<(unsigned)blocks, 256, smem, stream>>>(
          (const unsigned long long*)x, (unsigned long long*)out, g, total);
      break;
  }
  EINCONV_CHECK_LAUNCH("unfold_transpose_kernel");
  return EINCONV_OK;
}

}  // namespace einconv
