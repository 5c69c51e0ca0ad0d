// TMA-fed tcgen05 kernels for the two reduction-dominated networks: convNd weight VJP
// (expressions/convNd_weight_vjp.py:51-58) and the KFC factor (expressions/convNd_kfc.py:75-90).
//
//   weight VJP   gW[co, (ci,tap)]        = sum_{b,o} V[b,co,o]        * X[b,ci, s*o + d*tap - p]
//   KFC          K [(ci,tap),(ci',tap')] = sum_{b,o} X[b,ci, i(o,tap)] * X[b,ci', i(o,tap')]
//
// Both reduce over output positions, and in the reference's channels-first layout positions are the
// contiguous axis — i.e. both operands are already K-major.  A K block is a small patch of output
// positions (BD x BH x BW = 64 bf16 / 32 fp32 positions = one 128-byte row per channel) of one sample.
// For tap t the matching input patch is the same box shifted by d*t - p: ONE TMA tiled load per
// (tap, channel block) brings `CB` channels x 128 B into shared memory with the 128-byte swizzle the
// tensor core expects; out-of-bounds coordinates are zero-filled by the TMA unit, which *is* the
// zero padding of the convolution (and of partial patches at the image border).  No index pattern,
// no unfolded tensor, no per-element address arithmetic on the SMs.
//
// GEMM tile: M = 128 rows = (128 / CB) consecutive taps x CB channels of the unfolded input,
//            N = BN columns = output channels (weight VJP, read from the cotangent) or another
//                (taps x channels) block (KFC),  K = positions, split across CTAs (split-K).
// Roles (256 threads): warp 0 lane 0 issues TMA; warp 1 lane 0 issues tcgen05.mma (kind::f16 or
// kind::tf32, M=128, fp32 accumulators in TMEM); warp 2 owns the TMEM allocation; warps 4-7 drain
// the accumulator and write fp32 partial tiles, which a small second kernel sums over the splits,
// scales, un-permutes (rows are produced tap-major) and casts.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include "dispatch.cuh"

namespace einconv {
namespace tma {

constexpr int kStages = 4;
constexpr int kTileM = 128;
constexpr int kThreads = 256;
constexpr int kRowBytes = 128;  // one swizzle-128B row = one channel's K block
constexpr int kMaxKw = 32;      // widest kernel (innermost dim) the TMA path handles
constexpr int kMaxMaps = 32;    // tensor maps kept at the head of the workspace

// ---------------------------------------------------------------------------------------------
// PTX wrappers (same conventions as tc_gemm.cu)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s: a protocol bug must not hang the GPU
  }
}
// for warps that wait for the whole main loop (epilogue): sleep between polls so that they do not
// take issue slots from the producer / MMA warps of their sub-partition
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(256);
    if (clock64() - t0 > 20000000000ll) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one lane of a converged warp (single-thread issue from warp-uniform code, see the MMA issuer of kfc_super_kernel)
__device__ __forceinline__ bool ptx_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                     uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 5-d TMA tiled load, completion on an mbarrier (coordinates innermost first)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// K-major swizzled shared-memory descriptor: rows of `row_bytes`, 8-row groups 8 * row_bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 @0, LBO >> 4 @16 (unused, 1), SBO >> 4 @32, version 1 @46,
//  layout SWIZZLE_128B = 2 @61)
// `row_bytes` = 128 / 64 / 32 selects SWIZZLE_128B (2) / 64B (4) / 32B (6); an 8-row group is 8 * row_bytes.
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N) {
  // c_format F32 = 1 @4, a/b_format @7/@10, both K-major, N >> 3 @17, M >> 4 @24
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
struct MapArray {
  CUtensorMap m[kMaxMaps];
};

struct Params {
  // reduction blocks: patches of positions, enumerated (b, d-tile, h-tile, w-tile), w fastest
  int n_wt, n_ht, n_dt;
  FastDiv wt_div, ht_div, dt_div;
  int BW, BH, BD;                 // patch extent in output positions
  int PB, n_pieces;               // bytes of one (d, h) line of the patch per channel; lines per patch
  int k_blocks, k_splits;
  // geometry of the three innermost spatial dims (W, H, D order), missing dims are unit
  int stride[3], dil[3], pad[3], k[3];
  FastDiv kw_div, kh_div;         // tap index -> (kd, kh, kw)
  int k_total;
  // rows: CB channels x (128 / CB) taps per 128-row tile
  int CB, taps_per_tile, tap_groups;   // tap_groups = ceil(k_total / taps_per_tile)
  int c_in_g, c_out_g, groups;
  int c_blocks;                        // ceil(c_in_g / CB)
  int m_tiles, n_tiles;                // m_tiles = c_blocks * tap_groups
  int Mp, Np;                          // padded partial extents: m_tiles * 128, n_tiles * BN
  // innermost-dimension taps: TMA needs 16-byte aligned start coordinates, so tap kw reads from the
  // pre-shifted copy `kw_map[kw]` of x at the aligned coordinate ow0 * stride + kw_c0[kw]
  signed char kw_map[kMaxKw];
  signed char kw_map_b[kMaxKw];        // KFC column operand: same data, W extent clipped to valid outputs
  int out_h, out_d;                    // output extents of the two outer dims (line validity)
  int kw_c0[kMaxKw];
  int v_map, v_c0;                     // tensor map index of the cotangent and its coordinate margin
  int is_kfc;
  int tri;                             // KFC: only tiles on or above the diagonal are computed (K is symmetric)
  int debug;                           // bring-up switch (EINCONV_TMA_DEBUG): 1 = no TMA/MMA, 2 = TMA only
  float* partial;                      // [groups][k_splits][Mp][Np]
  const float* scale_dev;              // optional extra factor read on the device by the finish kernel
  float* direct_out;                   // KFC, one split, fp32 output: the epilogue writes the (un-permuted) factor itself
  float direct_scale;
  int direct_a;                        // A = c_in_g * k_total (rows / columns of the factor)
  int s_tiles;                         // KFC super-tile kernel: ceil(m_tiles / 2) super-tiles (256 rows) per side
  // flat-plane mode (see flat_copy_kernel): per-tap tensor map (row / column operand) and start coordinate
  int st_n, st_bytes, st_a_bytes, st_rows;  // pipeline stages of the super-tile kernel (see kMaxStagesS)
  int box_split;                            // channel parts per TMA box (1, 2 or 4)
  int flat;
  signed char tap_map_a[kMaxMaps], tap_map_b[kMaxMaps];
  int tap_c0[kMaxMaps];
};

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr bool TF32 = true; static constexpr int FMT = 2; static constexpr int UMMA_K = 8; };
template <> struct Elem<__nv_bfloat16> { static constexpr bool TF32 = false; static constexpr int FMT = 1; static constexpr int UMMA_K = 16; };
template <> struct Elem<__half> { static constexpr bool TF32 = false; static constexpr int FMT = 0; static constexpr int UMMA_K = 16; };

template <int BN>
struct Smem {
  static constexpr int a_bytes = kTileM * kRowBytes;  // 16 KB
  static constexpr int b_bytes = BN * kRowBytes;
  static constexpr int stage_bytes = a_bytes + b_bytes;
  static constexpr int total = kStages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <typename T, int BN>
__global__ void __launch_bounds__(kThreads, 1)
tma_reduce_gemm_kernel(const __grid_constant__ MapArray map_array, const __grid_constant__ Params p) {
  // tensor maps travel as kernel parameters (4 KB; CUDA >= 12.1 allows 32 KB): no per-call upload and
  // nothing host-side is read when a captured graph replays
  const CUtensorMap* maps = map_array.m;
  // tensor maps live in global memory (head of the caller's workspace): one per shifted copy of the
  // input, plus the cotangent's
  using E = Elem<T>;
  using S = Smem<BN>;
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles must start on 1024-byte boundaries
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* accum_bar = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  int m_tile = blockIdx.x;
  int n_tile = blockIdx.y;
  if (p.tri) {
    // blockIdx.x enumerates the tiles (i, j >= i) of the symmetric factor row by row
    int rest = blockIdx.x, i = 0;
    while (rest >= p.m_tiles - i) {
      rest -= p.m_tiles - i;
      ++i;
    }
    m_tile = i;
    n_tile = i + rest;
  }
  const int grp = blockIdx.z / p.k_splits;
  const int split = blockIdx.z - grp * p.k_splits;
  const int kb_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(p.k_blocks, kb_begin + kb_per_split);
  const int n_kb = max(0, kb_end - kb_begin);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // rows of this CTA's tiles
  const int a_cblock = m_tile % p.c_blocks, a_tapgroup = m_tile / p.c_blocks;
  const int b_cblock = n_tile % p.c_blocks, b_tapgroup = n_tile / p.c_blocks;  // KFC only

  // patch walker: reduction block kb -> (b, d-tile, h-tile, w-tile), advanced incrementally
  struct Walker {
    int wt, ht, dt, b;
    __device__ void init(const Params& p, int kb) {
      uint32_t rest = (uint32_t)kb, a, c, d, e;
      p.wt_div.divmod(rest, rest, a);
      p.ht_div.divmod(rest, rest, c);
      p.dt_div.divmod(rest, e, d);
      wt = (int)a; ht = (int)c; dt = (int)d; b = (int)e;
    }
    __device__ void next(const Params& p) {
      if (++wt == p.n_wt) { wt = 0; if (++ht == p.n_ht) { ht = 0; if (++dt == p.n_dt) { dt = 0; ++b; } } }
    }
  };

  if (warp == 0) {
    // ================================= TMA producer ==========================================
    // A k block needs n_pieces * (a_taps + b_loads) boxes.  Every lane owns up to kLoadsPerLane of
    // them and precomputes everything that does not depend on the patch position, so the per-block
    // work of a lane is a handful of adds and the TMA instruction itself.
    constexpr int kLoadsPerLane = 4;
    const int b_rows_per_load = p.is_kfc ? p.CB : BN;
    const int b_loads = p.is_kfc ? BN / p.CB : 1;
    const int a_taps = min(p.taps_per_tile, p.k_total - a_tapgroup * p.taps_per_tile);
    const int b_taps = p.is_kfc ? min(b_loads, p.k_total - b_tapgroup * b_loads) : 1;
    const int per_piece = a_taps + b_taps;
    const int n_loads = p.n_pieces * per_piece;
    const uint32_t line_tx = (uint32_t)(a_taps * p.CB + b_taps * b_rows_per_load) * p.PB;
    const int bh_shift = 31 - __clz(p.BH);  // BH is a power of two

    const CUtensorMap* l_map[kLoadsPerLane];
    int l_dst[kLoadsPerLane], l_cw[kLoadsPerLane], l_ch[kLoadsPerLane], l_cd[kLoadsPerLane];
    int l_hh[kLoadsPerLane], l_dd[kLoadsPerLane], l_chan[kLoadsPerLane];
    bool l_on[kLoadsPerLane];
#pragma unroll
    for (int q = 0; q < kLoadsPerLane; ++q) {
      const int li = lane + 32 * q;
      l_on[q] = li < n_loads;
      const int pc = l_on[q] ? li / per_piece : 0;
      const int slot = l_on[q] ? li - pc * per_piece : 0;
      l_hh[q] = pc & (p.BH - 1);
      l_dd[q] = pc >> bh_shift;
      const bool is_a = slot < a_taps;
      int tap, dst, chan;
      const CUtensorMap* mp;
      if (is_a) {
        tap = a_tapgroup * p.taps_per_tile + slot;
        dst = pc * (kTileM * p.PB) + slot * p.CB * p.PB;
        chan = grp * p.c_in_g + a_cblock * p.CB;
      } else if (p.is_kfc) {
        tap = b_tapgroup * b_loads + (slot - a_taps);
        dst = S::a_bytes + pc * (BN * p.PB) + (slot - a_taps) * p.CB * p.PB;
        chan = grp * p.c_in_g + b_cblock * p.CB;
      } else {
        tap = -1;
        dst = S::a_bytes + pc * (BN * p.PB);
        chan = grp * p.c_out_g + n_tile * BN;
      }
      if (tap >= 0) {
        uint32_t kd, kh, kw, r2;
        p.kw_div.divmod((uint32_t)tap, r2, kw);
        p.kh_div.divmod(r2, kd, kh);
        mp = maps + (is_a ? p.kw_map[kw] : p.kw_map_b[kw]);
        l_cw[q] = p.kw_c0[kw];
        l_ch[q] = (int)kh * p.dil[1] - p.pad[1];
        l_cd[q] = (int)kd * p.dil[2] - p.pad[2];
      } else {
        mp = maps + p.v_map;
        l_cw[q] = p.v_c0;
        l_ch[q] = 0;
        l_cd[q] = 0;
      }
      l_map[q] = mp;
      l_dst[q] = dst;
      l_chan[q] = chan;
    }
    // the cotangent is indexed by output coordinates, the input by stride * output + tap offset
    const int sw = p.stride[0], sh = p.stride[1], sd = p.stride[2];

    Walker wk;
    wk.init(p, kb_begin);
    for (int i = 0; i < n_kb; ++i, wk.next(p)) {
      const int s = i % kStages;
      const uint32_t ph = (uint32_t)(i / kStages) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const int ow0 = wk.wt * p.BW, oh0 = wk.ht * p.BH, od0 = wk.dt * p.BD;
      if (lane == 0) {
        if (p.debug == 1) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
        } else {
          // lines of the patch beyond the output extent contribute nothing: not loaded, not multiplied
          int n_valid = 0;
          for (int pc = 0; pc < p.n_pieces; ++pc)
            n_valid += (oh0 + (pc & (p.BH - 1)) < p.out_h) && (od0 + (pc >> bh_shift) < p.out_d);
          mbar_expect_tx(&full_bar[s], line_tx * n_valid);
        }
      }
      __syncwarp();
      if (p.debug != 1) {
        unsigned char* stage = smem + s * S::stage_bytes;
#pragma unroll
        for (int q = 0; q < kLoadsPerLane; ++q) {
          if (!l_on[q]) continue;
          const int oh = oh0 + l_hh[q], od = od0 + l_dd[q];
          if (oh >= p.out_h || od >= p.out_d) continue;
          const bool is_v = l_map[q] == maps + p.v_map && !p.is_kfc;
          tma_load_5d(stage + l_dst[q], l_map[q], &full_bar[s],
                      (is_v ? ow0 : ow0 * sw) + l_cw[q], (is_v ? oh : oh * sh) + l_ch[q],
                      (is_v ? od : od * sd) + l_cd[q], l_chan[q], wk.b);
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer ============================================
    // warp-uniform loop, one elected lane issues: operands stay in uniform registers (see kfc_super_kernel)
    constexpr uint32_t idesc = make_idesc(E::FMT, kTileM, BN);
    const int steps_per_line = p.PB / 32;  // 32 bytes of K per instruction (16 bf16 / 8 tf32)
    const int bh_shift = 31 - __clz(p.BH);
    const uint64_t desc_const = make_desc_kmajor(0, p.PB);
    const uint32_t smem16 = smem_u32(smem) >> 4;
    constexpr uint32_t stage16 = (uint32_t)S::stage_bytes >> 4, oper16 = (uint32_t)S::a_bytes >> 4;
    const uint32_t a_piece16 = (uint32_t)(kTileM * p.PB) >> 4, b_piece16 = (uint32_t)(BN * p.PB) >> 4;
    Walker wk;
    wk.init(p, kb_begin);
    for (int i = 0; i < n_kb; ++i, wk.next(p)) {
      const int s = i % kStages;
      const uint32_t ph = (uint32_t)(i / kStages) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const int oh0 = wk.ht * p.BH, od0 = wk.dt * p.BD;
      if (ptx_elect_one()) {
        const uint32_t a16 = smem16 + (uint32_t)s * stage16, b16 = a16 + oper16;
        for (int pc = 0; pc < p.n_pieces; ++pc) {
          if (oh0 + (pc & (p.BH - 1)) >= p.out_h || od0 + (pc >> bh_shift) >= p.out_d) continue;
          for (int j = 0; j < steps_per_line; ++j)
            umma<E::TF32>(tmem_base, desc_const + (uint64_t)(a16 + (uint32_t)pc * a_piece16 + 2u * j),
                          desc_const + (uint64_t)(b16 + (uint32_t)pc * b_piece16 + 2u * j), idesc, (uint32_t)(i | pc | j));
        }
        umma_commit(&empty_bar[s]);
        if (i == n_kb - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================= epilogue ==============================================
    const int ew = warp - 4;              // TMEM lane quarter owned by this warp
    const int row = ew * 32 + lane;       // D row held by this thread
    if (n_kb > 0) {
      mbar_wait_relaxed(accum_bar, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    if (p.direct_out) {
      // One split: scale, un-permute (tile rows are tap-major) and store the tile and its mirror
      // straight into the factor.  Saves the partial round trip and the finish launch, which are
      // batch-independent costs and dominate at small per-rank batches.
      const float scale = p.scale_dev ? p.direct_scale * __ldg(p.scale_dev) : p.direct_scale;
      const int A = p.direct_a;  // c_in_g * k_total
      const int Kt = p.k_total;
      // tile row -> (tap, channel) -> factor index a = ci * Kt + tap   (CB is a power of two)
      const int cb_shift = 31 - __clz(p.CB);
      const int m_tg = m_tile / p.c_blocks, m_cb = m_tile - m_tg * p.c_blocks;
      const int n_tg = n_tile / p.c_blocks, n_cb = n_tile - n_tg * p.c_blocks;
      const int r_tap = m_tg * p.taps_per_tile + (row >> cb_shift);
      const int r_ci = m_cb * p.CB + (row & (p.CB - 1));
      const bool r_ok = r_tap < Kt && r_ci < p.c_in_g;
      const long long a = (long long)r_ci * Kt + r_tap;
      float* og = p.direct_out + (long long)grp * A * A;
#pragma unroll 1
      for (int n0 = 0; n0 < BN; n0 += 16) {
        uint32_t r[16];
        if (n_kb > 0) {
          tmem_ld16(lane_base + n0, r);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        // 16 consecutive columns share their tap (CB >= 32 is a multiple of 16)
        const int c_tap = n_tg * p.taps_per_tile + (n0 >> cb_shift);
        const int c_ci0 = n_cb * p.CB + (n0 & (p.CB - 1));
        if (r_ok && c_tap < Kt) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c_ci = c_ci0 + j;
            if (c_ci < p.c_in_g) {
              const long long a2 = (long long)c_ci * Kt + c_tap;
              const float v = __uint_as_float(r[j]) * scale;
              og[a * A + a2] = v;
              if (m_tile != n_tile) og[a2 * A + a] = v;
            }
          }
        }
      }
    } else {
    float* dst = p.partial + (((long long)grp * p.k_splits + split) * p.Mp + (m_tile * kTileM + row)) * p.Np +
                 n_tile * BN;
#pragma unroll 1
    for (int n0 = 0; n0 < BN; n0 += 16) {
      uint32_t r[16];
      if (n_kb > 0) {
        tmem_ld16(lane_base + n0, r);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      float4* d4 = reinterpret_cast<float4*>(dst + n0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        d4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// ---------------------------------------------------------------------------------------------
// KFC with 256 x 256 super-tiles.
//
// The 128 x 128 kernel above pulls (128 + 128) rows x 128 B from L2 for every 4 x 64 tensor cycles:
// 128 B/clk/SM, three times what the L2 can deliver to 148 SMs at once (~42 B/clk/SM, measured ~6300 B/clk
// chip-wide), so it ran at a third of the tensor rate.  Here a CTA owns TWO row tiles and TWO column tiles
// of the factor: a K block brings (256 + 256) rows and feeds 2 MMAs of N = 256 per 32-byte K step (the whole
// 512-column TMEM: accumulator (mi) = columns [256 mi, 256 mi + 256)), i.e. 64 KB per 1024 tensor cycles =
// 64 B/clk/SM; an N = 256 instruction also halves the shared-memory read per flop (12 KB per 128 cycles).
// Super-tiles on the diagonal use their row operand as the column operand too (the column operand's maps,
// whose positions beyond the output width read as zero, are used for both) and load nothing else.
// Only super-tiles (I, J >= I) exist; inside a diagonal one the mirror block (2I+1, 2I) is not computed.
// ---------------------------------------------------------------------------------------------
// Stage geometry (Params::st_*): in general 3 stages of [256 rows | 256 rows] x 128 B = 64 KB.  A factor with a
// single (diagonal) super-tile never loads a column operand, and with a single 128-row tile only half of the
// row operand: the same 192 KB then hold 6 stages of 32 KB or 12 of 16 KB.  These are the 1 x 1 layers with
// A <= 256, whose K blocks are small (8-32 KB): with 3 of them in flight the kernel was bound by the TMA round
// trip (Little: 24 KB / ~1.5 us per SM = 2.4 TB/s on the A = 64 layers).
constexpr int kMaxStagesS = 12;
constexpr int kSuperRing = 3 * 2 * 2 * kTileM * kRowBytes;   // 192 KB
constexpr int kSuperSmem = kSuperRing + 1024 + 512;

constexpr int kMaxOrdered = 1024;
struct TileOrder {
  unsigned short t[kMaxOrdered];  // super-tile ids, heaviest first (identity beyond kMaxOrdered tiles)
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
kfc_super_kernel(const __grid_constant__ MapArray map_array, const __grid_constant__ Params p,
                 const __grid_constant__ TileOrder order) {
  const CUtensorMap* maps = map_array.m;
  using E = Elem<T>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSuperRing);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStagesS;
  uint64_t* accum_bar = bars + 2 * kMaxStagesS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStagesS + 1);
  const int kStagesS = p.st_n;
  const int kSuperStageBytes = p.st_bytes;
  const int kSuperOperandBytes = p.st_a_bytes;
  const int kSuperRows = p.st_rows;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // Grid = (splits x groups, 1, super-tiles): CTAs are handed out x-fastest, so all splits of the heaviest
  // super-tile go first and the light ones (diagonal: 3 of 4 blocks, last row / column of an odd tile count:
  // half empty) fill the tail.  With one CTA per SM and a few waves of CTAs this longest-first order keeps
  // the SMs evenly loaded although the split count is the same for every tile.
  // blockIdx.z -> super-tile (I, J >= I), enumerated row by row
  int sI = 0, sJ = 0;
  {
    const int n_super = p.s_tiles * (p.s_tiles + 1) / 2;
    int rest = n_super <= kMaxOrdered ? (int)order.t[blockIdx.z] : (int)blockIdx.z, i = 0;
    while (rest >= p.s_tiles - i) {
      rest -= p.s_tiles - i;
      ++i;
    }
    sI = i;
    sJ = i + rest;
  }
  const bool diag = sI == sJ;
  const int grp = blockIdx.x / p.k_splits;
  const int split = blockIdx.x - grp * p.k_splits;
  const int kb_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(p.k_blocks, kb_begin + kb_per_split);
  const int n_kb = max(0, kb_end - kb_begin);
  // row tiles 2I, 2I+1 and column tiles 2J, 2J+1 (the second of each may not exist)
  const bool m1 = 2 * sI + 1 < p.m_tiles;
  const bool n1 = 2 * sJ + 1 < p.m_tiles;

  if (tid == 0) {
    for (int s = 0; s < kStagesS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  struct Walker {
    int wt, ht, dt, b;
    __device__ void init(const Params& p, int kb) {
      uint32_t rest = (uint32_t)kb, a, c, d, e;
      p.wt_div.divmod(rest, rest, a);
      p.ht_div.divmod(rest, rest, c);
      p.dt_div.divmod(rest, e, d);
      wt = (int)a; ht = (int)c; dt = (int)d; b = (int)e;
    }
    __device__ void next(const Params& p) {
      if (++wt == p.n_wt) { wt = 0; if (++ht == p.n_ht) { ht = 0; if (++dt == p.n_dt) { dt = 0; ++b; } } }
    }
  };

  if (warp == 0) {
    // ================================= TMA producer ==========================================
    // (Measured alternative: one elected lane issuing every box from warp-uniform code.  The instruction takes
    // uniform registers, so per-lane operands cost an ELECT + R2UR.BROADCAST loop per box; yet the single-lane
    // version was slower — sweep 18.0 -> 20.3 ms — because each issue stalls its thread for hundreds of cycles
    // and the per-lane loops of the four unrolled slots overlap.)
    // boxes of one K block: per patch line (piece), per operand tile (A0, A1, B0, B1 as far as they exist
    // and are needed), per tap of the tile.  Lane l owns boxes l, l + 32, ...; everything that does not
    // depend on the patch position is computed once.
    constexpr int kLoadsPerLane = 4;
    const int tiles_a = m1 ? 2 : 1;
    const int tiles_b = diag ? 0 : (n1 ? 2 : 1);
    const int n_tiles_cta = tiles_a + tiles_b;
    const int tpt = p.taps_per_tile;
    const int per_piece = n_tiles_cta * tpt;   // slots; slots of taps beyond k_total are skipped
    const int n_loads = p.n_pieces * per_piece;
    const int bh_shift = 31 - __clz(p.BH);

    const CUtensorMap* l_map[kLoadsPerLane];
    int l_dst[kLoadsPerLane], l_cw[kLoadsPerLane], l_ch[kLoadsPerLane], l_cd[kLoadsPerLane];
    int l_hh[kLoadsPerLane], l_dd[kLoadsPerLane], l_chan[kLoadsPerLane];
    bool l_on[kLoadsPerLane];
    int boxes_per_piece = 0;  // valid boxes of one patch line (same for every piece)
#pragma unroll
    for (int q = 0; q < kLoadsPerLane; ++q) {
      const int li = lane + 32 * q;
      bool on = li < n_loads;
      const int pc = on ? li / per_piece : 0;
      const int slot = on ? li - pc * per_piece : 0;
      const int t_idx = slot / tpt;            // operand tile of this CTA: [A0, A1, B0, B1] compacted
      const int tl = slot - t_idx * tpt;       // tap slot inside the tile
      const bool is_a = t_idx < tiles_a;
      const int sub = is_a ? t_idx : t_idx - tiles_a;      // 0 / 1 within the operand
      const int tile = is_a ? 2 * sI + sub : 2 * sJ + sub;  // row / column tile of the factor
      const int cblock = tile % p.c_blocks, tapgroup = tile / p.c_blocks;
      const int tap = tapgroup * tpt + tl;
      on = on && tap < p.k_total;
      l_on[q] = on;
      l_hh[q] = pc & (p.BH - 1);
      l_dd[q] = pc >> bh_shift;
      uint32_t kd = 0, kh = 0, kw = 0, r2;
      if (on) {
        p.kw_div.divmod((uint32_t)tap, r2, kw);
        p.kh_div.divmod(r2, kd, kh);
      }
      // the row operand may hold anything beyond the output width only if the column operand is zero there;
      // on the diagonal the row operand IS the column operand, so it uses the clipped maps as well
      l_map[q] = maps + ((is_a && !diag) ? p.kw_map[kw] : p.kw_map_b[kw]);
      l_cw[q] = p.kw_c0[kw];
      l_ch[q] = (int)kh * p.dil[1] - p.pad[1];
      l_cd[q] = (int)kd * p.dil[2] - p.pad[2];
      if (p.flat) {  // one flat coordinate per tap, its own (clipped) map
        const int tt = on ? tap : 0;
        l_map[q] = maps + ((is_a && !diag) ? p.tap_map_a[tt] : p.tap_map_b[tt]);
        l_cw[q] = p.tap_c0[tt];
        l_ch[q] = 0;
        l_cd[q] = 0;
      }
      l_dst[q] = (is_a ? 0 : kSuperOperandBytes) + pc * (kSuperRows * p.PB) + sub * (kTileM * p.PB) + tl * p.CB * p.PB;
      l_chan[q] = grp * p.c_in_g + cblock * p.CB;
    }
    for (int slot = 0; slot < per_piece; ++slot) {
      const int t_idx = slot / tpt, tl = slot - t_idx * tpt;
      const bool is_a = t_idx < tiles_a;
      const int sub = is_a ? t_idx : t_idx - tiles_a;
      const int tile = is_a ? 2 * sI + sub : 2 * sJ + sub;
      boxes_per_piece += ((tile / p.c_blocks) * tpt + tl) < p.k_total;
    }
    const uint32_t line_tx = (uint32_t)boxes_per_piece * p.CB * p.PB;
    const int sw = p.stride[0], sh = p.stride[1], sd = p.stride[2];

    Walker wk;
    wk.init(p, kb_begin);
    int s = -1;
    uint32_t ph = 1u;  // flips to 0 at the first wrap below
    for (int i = 0; i < n_kb; ++i, wk.next(p)) {
      if (++s == kStagesS || i == 0) { s = 0; ph ^= 1u; }  // (the stage count is a run-time value: no division per K block)
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const int ow0 = wk.wt * p.BW, oh0 = wk.ht * p.BH, od0 = wk.dt * p.BD;
      if (lane == 0) {
        int n_valid = 0;
        for (int pc = 0; pc < p.n_pieces; ++pc)
          n_valid += (oh0 + (pc & (p.BH - 1)) < p.out_h) && (od0 + (pc >> bh_shift) < p.out_d);
        mbar_expect_tx(&full_bar[s], line_tx * n_valid);
      }
      __syncwarp();
      unsigned char* stage = smem + s * kSuperStageBytes;
#pragma unroll
      for (int q = 0; q < kLoadsPerLane; ++q) {
        if (!l_on[q]) continue;
        const int oh = oh0 + l_hh[q], od = od0 + l_dd[q];
        if (oh >= p.out_h || od >= p.out_d) continue;
        tma_load_5d(stage + l_dst[q], l_map[q], &full_bar[s], ow0 * sw + l_cw[q], oh * sh + l_ch[q],
                    od * sd + l_cd[q], l_chan[q], wk.b);
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer ============================================
    // ONE thread runs the whole loop (waits, MMAs, commits).  tcgen05.mma is queued only one or two
    // instructions deep, so whatever the issuing thread does between the last MMA of a K block and the first of
    // the next (barrier wait, commit, loop control) is exposed as soon as it exceeds one MMA (128 cycles for
    // N = 256, 64 for N = 128).  Measured with scripts/mma_pipe.cu (no data movement, 8 x N = 256 per stage =
    // 1024 tensor cycles): whole-warp loop with an elected issuer 1246 cycles per stage, this form 1093; and
    // issuing from an `if (lane == 0)` INSIDE a warp-wide loop makes the compiler broadcast every descriptor
    // through R2UR (~350 cycles per MMA, the first version of this kernel).
    if (lane == 0) {
      constexpr uint32_t idesc256 = make_idesc(E::FMT, kTileM, 256);
      constexpr uint32_t idesc128 = make_idesc(E::FMT, kTileM, 128);
      const uint32_t idesc_row0 = n1 ? idesc256 : idesc128;
      const int steps_per_line = p.PB / 32;  // 32 bytes of K per instruction (16 fp16 / 8 tf32)
      const int bh_shift = 31 - __clz(p.BH);
      // descriptors differ only in the 14-bit start-address field (units of 16 bytes; shared-memory addresses
      // never carry out of it): constant part + offsets
      const uint64_t desc_const = make_desc_kmajor(0, p.PB);
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const uint32_t stage16 = (uint32_t)kSuperStageBytes >> 4;
      const uint32_t oper16 = diag ? 0u : ((uint32_t)kSuperOperandBytes >> 4);
      const uint32_t piece16 = (uint32_t)(kSuperRows * p.PB) >> 4, tile16 = (uint32_t)(kTileM * p.PB) >> 4;
      const bool one_piece = p.n_pieces == 1;
      Walker wk;
      wk.init(p, kb_begin);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_kb; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a16 = smem16 + (uint32_t)s * stage16;
        const uint32_t b16 = a16 + oper16;
        const int oh0 = wk.ht * p.BH, od0 = wk.dt * p.BD;
        for (int pc = 0; pc < p.n_pieces; ++pc) {
          if (!one_piece && (oh0 + (pc & (p.BH - 1)) >= p.out_h || od0 + (pc >> bh_shift) >= p.out_d)) continue;
          const uint32_t a_pc = a16 + (uint32_t)pc * piece16, b_pc = b16 + (uint32_t)pc * piece16;
          for (int j = 0; j < steps_per_line; ++j) {
            // piece 0 of a K block is always valid: the very first MMA of the tile overwrites the accumulator
            const uint32_t accumulate = (uint32_t)(i | pc | j);
            const uint64_t a0 = desc_const + (uint64_t)(a_pc + 2u * j), b0 = desc_const + (uint64_t)(b_pc + 2u * j);
            // row tile 0 against both column tiles (or the only one)
            umma<E::TF32>(tmem_base, a0, b0, idesc_row0, accumulate);
            if (m1) {
              const uint64_t a1 = a0 + (uint64_t)tile16;
              if (diag)  // the block (2I+1, 2I) is the mirror of (2I, 2I+1): only column tile 1
                umma<E::TF32>(tmem_base + 256 + 128, a1, b0 + (uint64_t)tile16, idesc128, accumulate);
              else
                umma<E::TF32>(tmem_base + 256, a1, b0, idesc_row0, accumulate);
            }
          }
        }
        umma_commit(&empty_bar[s]);
        if (i == n_kb - 1) umma_commit(accum_bar);
        if (++s == kStagesS) { s = 0; ph ^= 1u; }
        if (!one_piece) wk.next(p);
      }
    }
    __syncwarp();
  }

  // ================================= epilogue (all 8 warps) ====================================
  // Every thread of warps 4-7 holds one ROW of the accumulator (its TMEM lane).  Storing rows from there
  // makes each instruction touch 32 different rows (32 lone sectors: ~1 sector per clock and SM, 34 us per
  // super-tile measured).  Instead a 128 x 128 block goes through shared memory (the pipeline stages are
  // free by now; row pitch 129 floats is conflict-free for row-wise and column-wise accesses) and all 8
  // warps write whole 512-byte rows: the split-K partial tile, or — for single-tap kernels with one split,
  // where the factor index is the channel — the factor block and its mirror image directly.
  {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    if (warp >= 4 && n_kb > 0) {
      mbar_wait_relaxed(accum_bar, 0);
      tc_fence_after();
    }
    const int Kt = p.k_total;
    float* tile_s = reinterpret_cast<float*>(smem);
    constexpr int kPitch = 129;
    const bool direct = p.direct_out != nullptr;  // host sets it only for k_total == 1, one split, fp32 output
    const float scale = direct ? (p.scale_dev ? p.direct_scale * __ldg(p.scale_dev) : p.direct_scale) : 1.0f;
    for (int mi = 0; mi < (m1 ? 2 : 1); ++mi) {
      for (int nj = 0; nj < (n1 ? 2 : 1); ++nj) {
        const int m_tile = 2 * sI + mi, n_tile = 2 * sJ + nj;
        if (m_tile > n_tile) continue;  // mirror block of a diagonal super-tile: not computed
        if (warp >= 4) {
          const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(mi * 256 + nj * 128);
#pragma unroll 1
          for (int n0 = 0; n0 < 128; n0 += 16) {
            uint32_t r[16];
            if (n_kb > 0) {
              tmem_ld16(lane_base + n0, r);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) tile_s[row * kPitch + n0 + j] = __uint_as_float(r[j]) * scale;
          }
        }
        __syncthreads();
        int rows_valid = 128, cols_valid = 128;
        long long row_pitch;
        float* dst0;     // element (row 0, column 0) of the block in its destination
        float* mir0 = nullptr;
        if (direct) {
          // single tap: tile row r <-> channel tile * CB + r, only the first CB rows / columns of a tile exist
          const long long A = p.direct_a;
          rows_valid = min(p.CB, p.c_in_g - m_tile * p.CB);
          cols_valid = min(p.CB, p.c_in_g - n_tile * p.CB);
          float* og = p.direct_out + (long long)grp * A * A;
          row_pitch = A;
          dst0 = og + (long long)m_tile * p.CB * A + (long long)n_tile * p.CB;
          if (m_tile != n_tile) mir0 = og + (long long)n_tile * p.CB * A + (long long)m_tile * p.CB;
        } else {
          row_pitch = p.Np;
          dst0 = p.partial + (((long long)grp * p.k_splits + split) * p.Mp + (long long)m_tile * kTileM) * p.Np +
                 (long long)n_tile * 128;
        }
        // 4 rows x 4 column segments = 16 independent shared loads in flight per lane
        for (int r0 = warp * 4; r0 < rows_valid; r0 += 32) {
          float v[4][4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) v[a][k] = tile_s[min(r0 + a, 127) * kPitch + lane + 32 * k];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            if (r0 + a >= rows_valid) break;
            float* dst = dst0 + (long long)(r0 + a) * row_pitch;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (lane + 32 * k < cols_valid) dst[lane + 32 * k] = v[a][k];
          }
        }
        if (mir0) {
          for (int c0 = warp * 4; c0 < cols_valid; c0 += 32) {
            float v[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int k = 0; k < 4; ++k) v[a][k] = tile_s[(lane + 32 * k) * kPitch + min(c0 + a, 127)];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              if (c0 + a >= cols_valid) break;
              float* dst = mir0 + (long long)(c0 + a) * row_pitch;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (lane + 32 * k < rows_valid) dst[lane + 32 * k] = v[a][k];
            }
          }
        }
        __syncthreads();  // the block is out: shared memory may be refilled
        (void)Kt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// pre-pass of the flat-plane mode (2-d KFC whose output rows are shorter than a 128-byte box: 28, 14, 7
// wide images).  Per-row boxes waste the TMA unit there (a 14-wide fp16 row is a 32-byte request; the
// kernel measured 24 % tensor activity on 14 x 14 planes).  Instead every (sample, channel) plane is
// re-laid as ONE flat array per kernel column kw (stride 1) or per tap (stride > 1):
//     copy[q = r * pitch + j] = (j < O_w) ? x[r * s_h + kh_off - p_h][j * s_w + kw * d_w - p_w] : 0
// with pitch = O_w rounded up to 16 bytes.  Output position (oh, ow) is q = oh * pitch + ow; tap (kh, kw)
// of a stride-1 layer reads the kw-copy at q + kh * d_h * pitch (rows are shared between the kh), a strided
// layer reads its own copy at q.  A K block is then 128 contiguous bytes per channel whatever the image
// width (one box of CB x 128 B per tap and tile), columns j >= O_w are zero in the copy, and rows beyond
// the last output row are cut off by the extent of the column operand's tensor map (zero fill).
// One thread writes one 16-byte unit.
// ---------------------------------------------------------------------------------------------
struct FlatCopyParams {
  int W, H;               // input plane
  int pitch, rows;        // flat plane: rows x pitch elements (+ tail up to plane_stride)
  long long plane_stride; // elements between consecutive planes of one copy
  long long copy_stride;  // elements between copies
  int n_copies, per_tap;
  int k_w;                // kernel width (tap -> (kh, kw))
  int s_h, s_w, d_h, d_w, p_h, p_w, out_w;
  long long planes;
  uint32_t per_copy32;        // planes * units of one plane (the launch checks that it fits)
  FastDiv units_div, upr_div; // units per plane, units per row
};

// CONVERT: the source is the fp32 tensor itself and the copy is its power-of-two-scaled fp16 image (the scale
// follows from the absmax bits written by absmax_kernel; this replaces scale_to_half_kernel + a 16-bit
// flat_copy: one read of x instead of a read, a write and another read); thread 0 leaves 1 / s^2 for the finish.
template <typename T, bool CONVERT>
__global__ void __launch_bounds__(256) flat_copy_kernel(const void* __restrict__ x_any, T* __restrict__ out,
                                                        const __grid_constant__ FlatCopyParams fp,
                                                        const unsigned* __restrict__ absmax_bits,
                                                        float* __restrict__ inv_s2) {
  using S = std::conditional_t<CONVERT, float, T>;
  const S* x = reinterpret_cast<const S*>(x_any);
  constexpr int A = 16 / sizeof(T);
  float s = 1.f;
  if constexpr (CONVERT) {
    s = operand_scale(__ldg(absmax_bits));
    if (blockIdx.x == 0 && threadIdx.x == 0) *inv_s2 = 1.f / (s * s);
  }
  // blockIdx.y = copy; (plane, unit) from a 32-bit index with magic-number divisions (the 64-bit divisions of the
  // first version were most of this issue-bound kernel's instructions)
  const int units = (int)(fp.plane_stride / A);
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= fp.per_copy32) return;
  const int copy = blockIdx.y;
  const uint32_t plane = fp.units_div.quot(idx);
  const int u = (int)(idx - plane * (uint32_t)units);
  const int r = (int)fp.upr_div.quot((uint32_t)u), j0 = (u - r * (fp.pitch / A)) * A;  // one unit stays in one row
  int kh = 0, kw = copy;
  if (fp.per_tap) {
    kh = copy / fp.k_w;
    kw = copy - kh * fp.k_w;
  }
  const int i = fp.per_tap ? r * fp.s_h + kh * fp.d_h - fp.p_h : r - fp.p_h;
  const bool row_ok = r < fp.rows && i >= 0 && i < fp.H;
  const S* src = x + plane * ((long long)fp.W * fp.H) + (long long)(row_ok ? i : 0) * fp.W;
  alignas(16) T vals[A];
#pragma unroll
  for (int e = 0; e < A; ++e) {
    const int j = j0 + e;
    const int c = j * fp.s_w + kw * fp.d_w - fp.p_w;
    const bool ok = row_ok && j < fp.out_w && c >= 0 && c < fp.W;
    if constexpr (CONVERT) {
      vals[e] = ok ? (T)__half_as_ushort(__float2half_rn(__ldg(src + c) * s)) : T(0);
    } else {
      vals[e] = ok ? src[c] : T(0);
    }
  }
  *reinterpret_cast<uint4*>(out + (long long)copy * fp.copy_stride + (long long)plane * fp.plane_stride +
                            (long long)u * A) = *reinterpret_cast<const uint4*>(vals);
}

// ---------------------------------------------------------------------------------------------
// pre-pass: aligned shifted copies.  copy_r[row][j] = x[row][j - A + r] (zero outside [0, W)), rows =
// (b, c, d, h), j in [0, Wp), A = 16 / sizeof(T) elements, Wp a multiple of A.  A conv tap whose offset
// along the innermost dimension is t = A*q + r then reads copy_r at the ALIGNED coordinate
// w + A*(q + 1), which is what the TMA unit requires of its innermost start coordinate.
// One thread writes one 16-byte unit.
// ---------------------------------------------------------------------------------------------
// With a stride s > 1 along the innermost dimension the copy also subsamples (one copy per tap):
//   copy_t[row][j] = x[row][s * (j - A) + t],  so that  x[s*w + t] = copy_t[w + A].
template <typename T>
__global__ void __launch_bounds__(256) shift_copy_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                         long long rows, int W, int Wp, int r, int s) {
  constexpr int A = 16 / sizeof(T);
  const int units = Wp / A;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * units) return;
  // 32-bit division where the index fits (a 64-bit one is ~100 instructions of this 30-instruction kernel)
  const long long row = idx < (1ll << 32) ? (long long)((uint32_t)idx / (uint32_t)units) : idx / units;
  const int u = (int)(idx - row * units);
  const T* src = x + row * W;
  alignas(16) T vals[A];
#pragma unroll
  for (int e = 0; e < A; ++e) {
    const int j = s * (u * A + e - A) + r;
    vals[e] = (j >= 0 && j < W) ? src[j] : T(0);
  }
  *reinterpret_cast<uint4*>(out + row * Wp + (long long)u * A) = *reinterpret_cast<const uint4*>(vals);
}

// ---------------------------------------------------------------------------------------------
// second pass: sum the split-K partials, scale, un-permute rows/columns (tiles are tap-major:
// tile row r of (c-block, tap-group) is tap = group*TPT + r / CB, ci = block*CB + r % CB), cast.
//   weight VJP: out[grp*Cout_g + co][ci][tap]          (partial columns are co)
//   KFC:        out[grp][ci*Kt + tap][ci'*Kt + tap']
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) tma_finish_kernel(const float* __restrict__ partial, T* __restrict__ out,
                                                         const __grid_constant__ Params p, float scale_host) {
  const float scale = p.scale_dev ? scale_host * __ldg(p.scale_dev) : scale_host;
  const int Kt = p.k_total, A = p.c_in_g * Kt;
  const long long per_group = p.is_kfc ? (long long)A * A : (long long)p.c_out_g * A;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_group * p.groups) return;
  const int grp = (int)(idx / per_group);
  const long long rem = idx - (long long)grp * per_group;
  auto tile_row = [&](int a) {  // a = ci*Kt + tap  ->  row index in the padded, tap-major partial
    const int ci = a / Kt, tap = a - ci * Kt;
    const int cblock = ci / p.CB, cl = ci - cblock * p.CB;
    const int tg = tap / p.taps_per_tile, tl = tap - tg * p.taps_per_tile;
    return (tg * p.c_blocks + cblock) * kTileM + tl * p.CB + cl;
  };
  int prow, pcol;
  if (p.is_kfc) {
    const int a = (int)(rem / A), a2 = (int)(rem - (long long)a * A);
    prow = tile_row(a);
    if (p.tri) {
      // columns are tiled exactly like the rows; tiles below the diagonal are read from their mirror
      int r = prow, c = tile_row(a2);
      if (r / kTileM > c / kTileM) {
        const int t = r;
        r = c;
        c = t;
      }
      const long long tile = (long long)p.Mp * p.Np;
      const float* src = partial + (long long)grp * p.k_splits * tile + (long long)r * p.Np + c;
      float s = 0.f;
      for (int k = 0; k < p.k_splits; ++k) s += src[(long long)k * tile];
      Cvt<T>::store(out + idx, s * scale);
      return;
    }
    // KFC column tiles hold (BN / CB) taps each
    const int ci = a2 / Kt, tap = a2 - ci * Kt;
    const int cblock = ci / p.CB, cl = ci - cblock * p.CB;
    const int tpt_n = (p.Np / p.n_tiles) / p.CB;
    const int tg = tap / tpt_n, tl = tap - tg * tpt_n;
    pcol = (tg * p.c_blocks + cblock) * (p.Np / p.n_tiles) + tl * p.CB + cl;
  } else {
    const int co = (int)(rem / A), a = (int)(rem - (long long)co * A);
    prow = tile_row(a);
    pcol = co;
  }
  const long long tile = (long long)p.Mp * p.Np;
  const float* src = partial + (long long)grp * p.k_splits * tile + (long long)prow * p.Np + pcol;
  float s = 0.f;
  for (int k = 0; k < p.k_splits; ++k) s += src[(long long)k * tile];
  Cvt<T>::store(out + idx, s * scale);
}

// ---------------------------------------------------------------------------------------------
// Finish pass of the symmetric (upper-triangular) KFC tiling, in TILE space: one CTA per computed
// 128 x 128 block (mt <= nt).  Reads are whole 512-byte partial rows (one float4 per lane, independent
// loads over the splits), all index arithmetic is per CTA / per lane / per row (no per-element division);
// the block and its mirror image are scattered into the factor, whose order is channel-major
// (a = ci * Kt + tap) while tile rows are tap-major.  Replaces the element-space kernel, which spent its
// time on ~8 integer divisions per output element and ran at ~1 TB/s.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) kfc_finish_tri_kernel(const float* __restrict__ partial, T* __restrict__ out,
                                                             const __grid_constant__ Params p, float scale_host,
                                                             int rows_per_cta) {
  __shared__ float4 red[8][32];
  const float scale = p.scale_dev ? scale_host * __ldg(p.scale_dev) : scale_host;
  int mt = 0, nt = 0;
  {
    int rest = blockIdx.x, i = 0;
    while (rest >= p.m_tiles - i) {
      rest -= p.m_tiles - i;
      ++i;
    }
    mt = i;
    nt = i + rest;
  }
  const int grp = blockIdx.z;
  const int Kt = p.k_total;
  const long long A = (long long)p.c_in_g * Kt;
  const int cb_shift = 31 - __clz(p.CB);
  const int m_tg = mt / p.c_blocks, m_cb = mt - m_tg * p.c_blocks;
  const int n_tg = nt / p.c_blocks, n_cb = nt - n_tg * p.c_blocks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.y: chunk of rows_per_cta (1, 2, 4 or 8) rows; the 8 warps are 8 / rows_per_cta split lanes per row,
  // so that a factor with few tiles and many splits still spreads over the machine
  const int row_local = warp % rows_per_cta;
  const int split_lane = warp / rows_per_cta, split_lanes = 8 / rows_per_cta;
  const int r = blockIdx.y * rows_per_cta + row_local;
  // this lane's four columns: same tap slot, consecutive channels
  const int c0 = 4 * lane;
  const int c_tap = n_tg * p.taps_per_tile + (c0 >> cb_shift);
  const int c_ci0 = n_cb * p.CB + (c0 & (p.CB - 1));
  const long long a2_0 = (long long)c_ci0 * Kt + c_tap;
  const long long tile = (long long)p.Mp * p.Np;
  const int r_tap = m_tg * p.taps_per_tile + (r >> cb_shift);
  const int r_ci = m_cb * p.CB + (r & (p.CB - 1));
  const bool ok = r_tap < Kt && r_ci < p.c_in_g && c_tap < Kt;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    const float* src = partial + (long long)grp * p.k_splits * tile + (long long)(mt * kTileM + r) * p.Np +
                       nt * kTileM + c0;
    int k = split_lane;
    for (; k + 3 * split_lanes < p.k_splits; k += 4 * split_lanes) {
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src + (long long)k * tile));
      const float4 v1 = __ldcs(reinterpret_cast<const float4*>(src + (long long)(k + split_lanes) * tile));
      const float4 v2 = __ldcs(reinterpret_cast<const float4*>(src + (long long)(k + 2 * split_lanes) * tile));
      const float4 v3 = __ldcs(reinterpret_cast<const float4*>(src + (long long)(k + 3 * split_lanes) * tile));
      acc.x += (v0.x + v1.x) + (v2.x + v3.x);
      acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z);
      acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; k < p.k_splits; k += split_lanes) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (long long)k * tile));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (split_lanes > 1) {
    red[warp][lane] = acc;
    __syncthreads();
    if (split_lane != 0) return;
    for (int q = 1; q < split_lanes; ++q) {
      const float4 v = red[q * rows_per_cta + row_local][lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (!ok) return;
  T* og = out + (long long)grp * A * A;
  const bool mirror = mt != nt;
  const long long a = (long long)r_ci * Kt + r_tap;
  const float vals[4] = {acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (c_ci0 + e < p.c_in_g) {
      const long long a2 = a2_0 + (long long)e * Kt;
      Cvt<T>::store(og + a * A + a2, vals[e]);
      if (mirror) Cvt<T>::store(og + a2 * A + a, vals[e]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Finish pass in OUTPUT space for kernels with several taps.  The factor is channel-major (a = ci * Kt + tap)
// while tile rows are tap-major, so the tile-space kernel above writes every element as a lone 4-byte sector
// (A = 4608: 139 us for 130 MB, 0.9 TB/s).  Here a CTA owns a pair of channel blocks (bi <= bj, BC channels
// each) and ALL Kt x Kt taps: it gathers the BC x Kt x BC x Kt sub-block S[ci][t][ci2][t2] from the Kt^2 tile
// fragments (64-byte runs, summed over the splits; fragments below the tile diagonal are read transposed from
// their mirror tile), keeps it in shared memory, and writes whole rows of the factor: BC * Kt contiguous floats
// per row for the block itself and, transposed out of shared memory, for its mirror image.
// Shared-memory layout: addr(ci, t, ci2, t2) = (ci * Kt + t) * TS + ci2 * Kt + t2 with TS = BC * Kt + 1 (odd): the
// direct fragment stores (lanes = ci2, stride Kt), the transposed fragment stores (lanes = ci, stride Kt * TS) and
// the mirror reads (lanes = (ci, t), stride TS) are conflict-free for odd Kt.
// ---------------------------------------------------------------------------------------------
template <typename T, int BC>
__global__ void __launch_bounds__(256) kfc_finish_block_kernel(const float* __restrict__ partial, T* __restrict__ out,
                                                               const __grid_constant__ Params p, float scale_host,
                                                               int n_cblocks) {
  extern __shared__ float fin_smem[];
  const float scale = p.scale_dev ? scale_host * __ldg(p.scale_dev) : scale_host;
  int bi = 0, bj = 0;
  {
    int rest = blockIdx.x, i = 0;
    while (rest >= n_cblocks - i) {
      rest -= n_cblocks - i;
      ++i;
    }
    bi = i;
    bj = i + rest;
  }
  const int grp = blockIdx.z;
  const int Kt = p.k_total, tpt = p.taps_per_tile;
  const int TS = BC * Kt + 1;
  const long long A = (long long)p.c_in_g * Kt;
  const long long tile = (long long)p.Mp * p.Np;
  const float* base = partial + (long long)grp * p.k_splits * tile;
  // tile-space coordinates of channel block b, tap t: tile (t / tpt) * c_blocks + cblock, row (t % tpt) * CB + cl
  const int ci0 = bi * BC, cj0 = bj * BC;
  const int cbl_i = ci0 / p.CB, cl_i = ci0 - cbl_i * p.CB;
  const int cbl_j = cj0 / p.CB, cl_j = cj0 - cbl_j * p.CB;
  constexpr int kItemsPerPass = 256 / BC;
  const int sub = threadIdx.x / BC, ln = threadIdx.x % BC;
  // per (t, t2) fragment: offset of its first element in a split's partial matrix (x), shared-memory base
  // t * TS + t2 and the orientation bit (y).  Built once per CTA: the divisions and 64-bit products of this index
  // arithmetic, done per loaded element, made the first version issue-bound (ncu: 74 % issue-active, 130
  // instructions per 4-byte load).
  int2* tbl = reinterpret_cast<int2*>(fin_smem + (size_t)BC * Kt * TS);
  for (int pr = threadIdx.x; pr < Kt * Kt; pr += 256) {
    const int t = pr / Kt, t2 = pr - t * Kt;
    const int tg = t / tpt, tl = t - tg * tpt, tg2 = t2 / tpt, tl2 = t2 - tg2 * tpt;
    const int mt = tg * p.c_blocks + cbl_i, nt = tg2 * p.c_blocks + cbl_j;
    const int row_i = tl * p.CB + cl_i, row_j = tl2 * p.CB + cl_j;  // first row / column of the fragment in its tile
    const bool direct = mt <= nt;  // else: read transposed from the mirror tile (nt, mt)
    const int off = direct ? (mt * kTileM + row_i) * p.Np + nt * kTileM + row_j
                           : (nt * kTileM + row_j) * p.Np + mt * kTileM + row_i;
    tbl[pr] = make_int2(off, (t * TS + t2) | (direct ? (1 << 30) : 0));
  }
  __syncthreads();
  const int n_items = Kt * Kt * BC;
  constexpr int U = 8;  // fragments in flight per thread (one split at small batches: latency, not bytes, is the cost)
  for (int it0 = sub; it0 < n_items; it0 += U * kItemsPerPass) {
    const float* src[U];
    int dst[U];
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = it0 + u * kItemsPerPass;
      src[u] = nullptr;
      dst[u] = -1;
      acc[u] = 0.f;
      if (it < n_items) {
        const int pr = it / BC, rr = it - pr * BC;  // BC is a power of two
        const int2 e = tbl[pr];
        const bool direct = (e.y >> 30) & 1;
        // direct: row = ci = rr, lanes run over ci2; transposed: row = ci2 = rr, lanes over ci
        const int ci = direct ? rr : ln, ci2 = direct ? ln : rr;
        dst[u] = ci * Kt * TS + ci2 * Kt + (e.y & 0x3fffffff);
        if (ci0 + ci < p.c_in_g && cj0 + ci2 < p.c_in_g) src[u] = base + e.x + rr * p.Np + ln;
      }
    }
    for (int k = 0; k < p.k_splits; ++k) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (src[u]) acc[u] += __ldcs(src[u] + (long long)k * tile);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] >= 0) fin_smem[dst[u]] = acc[u] * scale;
  }
  __syncthreads();
  T* og = out + (long long)grp * A * A;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = BC * Kt;                      // rows of the block = (ci, t), columns = (ci2, t2)
  const int rows_i = min(rows, (p.c_in_g - ci0) * Kt), rows_j = min(rows, (p.c_in_g - cj0) * Kt);
  for (int r = warp; r < rows_i; r += 8) {
    T* dst = og + ((long long)ci0 * Kt + r) * A + (long long)cj0 * Kt;
    const float* srow = fin_smem + r * TS;
    for (int c = lane; c < rows_j; c += 32) Cvt<T>::store(dst + c, srow[c]);
  }
  if (bi != bj) {
    for (int r = warp; r < rows_j; r += 8) {     // mirror row (ci2, t2), columns (ci, t)
      T* dst = og + ((long long)cj0 * Kt + r) * A + (long long)ci0 * Kt;
      for (int c = lane; c < rows_i; c += 32) Cvt<T>::store(dst + c, fin_smem[c * TS + r]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  });
  return fn;
}

struct Plan {
  bool ok = false;
  int BW, BH, BD, PB, kpos;
  int CB, taps_per_tile, tap_groups, c_blocks;
  int bn, m_tiles, n_tiles, k_blocks, k_splits;
  bool tri = false;  // KFC: compute only the tiles on or above the diagonal
  bool super = false;  // KFC: 256 x 256 super-tiles (kfc_super_kernel)
  int s_tiles = 0;
  int box_split = 1;          // super-tile kernel: TMA boxes of CB / box_split channels
  bool flat = false;          // KFC: flat-plane operand copies (flat_copy_kernel)
  bool flat_per_tap = false;  // one copy per tap (strided layers) instead of one per kernel column
  int flat_pitch = 0, flat_rows = 0;
  long long flat_plane = 0;   // elements per (sample, channel) plane of a copy
  int n_wt, n_ht, n_dt;
  long long Mp, Np;
  // shifted copies of the input (and an aligned copy of the cotangent when its rows are not
  // 16-byte multiples): residue r -> copy index, -1 = not needed, kDirect = read x itself
  int A;                       // elements per 16 bytes
  int n_copies;                // copies of x to write
  int copy_res[16];            // residue of copy i
  int res_map[16];             // residue -> tensor map index (or -1)
  bool x_direct;               // residue 0 served by x itself (rows already 16-byte multiples)
  bool per_tap;                // innermost stride > 1: one subsampled copy per tap, copy_res = tap offset
  bool v_copy;                 // cotangent needs an aligned copy
  int Wp_x, Wp_v;              // padded row pitch (elements) of the copies
  long long x_rows, v_rows;    // rows (b, c, d, h) of x / v
  long long copy_x_bytes, copy_v_bytes;
};

// geometry in (W, H, D) order with unit dims for missing ones
struct Geo3 {
  int in[3], out[3], k[3], stride[3], dil[3], pad[3];
};
static Geo3 geo3(const DevGeom& g) {
  Geo3 q;
  for (int i = 0; i < 3; ++i) {
    const int d = g.n_dims - 1 - i;
    if (d >= 0) {
      q.in[i] = g.in[d]; q.out[i] = g.out[d]; q.k[i] = g.k[d];
      q.stride[i] = g.stride[d]; q.dil[i] = g.dil[d]; q.pad[i] = g.pad[d];
    } else {
      q.in[i] = 1; q.out[i] = 1; q.k[i] = 1; q.stride[i] = 1; q.dil[i] = 1; q.pad[i] = 0;
    }
  }
  return q;
}

static Plan make_plan(int op, const DevGeom& g, int dtype) {
  Plan pl;
  if (g.n_dims > 3) return pl;
  const int es = (int)elem_size(dtype);
  if (es != 2 && es != 4) return pl;
  const Geo3 q = geo3(g);
  if (q.k[0] > kMaxKw) return pl;
  const long long x_elems = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
  if (x_elems >= (1ll << 40)) return pl;
  pl.kpos = kRowBytes / es;
  pl.flat = op == EINCONV_OP_KFC && g.n_dims == 2 && g.k_total > 1 && g.k_total <= 25 &&
            q.out[0] * es < kRowBytes && !getenv("EINCONV_KFC_FULL") && !getenv("EINCONV_KFC_TILE128") &&
            !getenv("EINCONV_KFC_NO_FLAT");
  if (pl.flat) {
    const int A_ = 16 / es;
    pl.flat_per_tap = q.stride[0] > 1 || q.stride[1] > 1;
    pl.flat_pitch = ((q.out[0] + A_ - 1) / A_) * A_;
    pl.flat_rows = pl.flat_per_tap ? q.out[1] : q.out[1] + (q.k[1] - 1) * q.dil[1];
    pl.flat_plane = (long long)pl.flat_rows * pl.flat_pitch;
    if (pl.flat_plane >= (1ll << 30)) pl.flat = false;
  }
  // patch shape: one (d, h) line of the patch is PB = 128 / 64 / 32 bytes per channel (BW = PB / es
  // positions, the TMA box's inner extent and the swizzle span); 128 / PB lines (BH x BD) make one
  // 128-byte K block.  Minimise out-of-range positions; prefer long lines on ties.
  double best = 1e30;
  for (int pb = 128; pb >= 32; pb /= 2) {
    const int bw = pb / es, lines = kRowBytes / pb;
    for (int bh = 1; bh <= lines; bh *= 2) {
      const int bd = lines / bh;
      const double cover = (double)ceil_div(q.out[0], bw) * bw * ceil_div(q.out[1], bh) * bh *
                           ceil_div(q.out[2], bd) * bd;
      const double waste = cover / ((double)q.out[0] * q.out[1] * q.out[2]);
      const double score = waste - 1e-4 * pb - 1e-6 * bh;
      if (score < best) {
        best = score;
        pl.BW = bw; pl.BH = bh; pl.BD = bd; pl.PB = pb;
      }
    }
  }
  if (best > 1e29) return pl;
  pl.n_wt = ceil_div(q.out[0], pl.BW);
  pl.n_ht = ceil_div(q.out[1], pl.BH);
  pl.n_dt = ceil_div(q.out[2], pl.BD);
  if (pl.flat) {  // K blocks are 128-byte runs of the flat output grid of one sample
    pl.PB = kRowBytes; pl.BW = kRowBytes / es; pl.BH = 1; pl.BD = 1;
    pl.n_wt = ceil_div(q.out[1] * pl.flat_pitch, pl.BW);
    pl.n_ht = 1; pl.n_dt = 1;
  }
  const long long kb = (long long)g.batch * pl.n_wt * pl.n_ht * pl.n_dt;
  if (kb >= (1ll << 31) || kb == 0) return pl;
  pl.k_blocks = (int)kb;

  // channel block: largest power of two <= 128 that does not exceed c_in_g rounded up to 8
  int cb = 128;
  while (cb > 8 && cb > g.c_in_g) cb /= 2;
  pl.CB = cb;
  // tiny channel blocks mean many 1 KB TMA boxes per k block: the CUDA-core kernel is faster there
  if (cb < 32) return pl;
  pl.taps_per_tile = kTileM / cb;
  pl.tap_groups = ceil_div(g.k_total, pl.taps_per_tile);
  pl.c_blocks = ceil_div(g.c_in_g, cb);
  pl.m_tiles = pl.c_blocks * pl.tap_groups;
  if (op == EINCONV_OP_KFC) {
    pl.bn = 128;  // same tiling as the rows
    pl.n_tiles = pl.m_tiles;
  } else {
    pl.bn = g.c_out_g > 128 ? 256 : (g.c_out_g > 64 ? 128 : 64);
    pl.n_tiles = ceil_div(g.c_out_g, pl.bn);
  }
  {
    const int b_loads = (op == EINCONV_OP_KFC) ? pl.bn / pl.CB : 1;
    if ((kRowBytes / pl.PB) * (pl.taps_per_tile + b_loads) > 128) return pl;  // boxes per k block (4 per lane)
  }
  pl.Mp = (long long)pl.m_tiles * kTileM;
  pl.Np = (long long)pl.n_tiles * pl.bn;
  // split-K so that the grid is exactly one wave of resident CTAs (a partial second wave would
  // double the run time): resident = 148 SMs x CTAs that fit in shared memory
  pl.tri = (op == EINCONV_OP_KFC) && !getenv("EINCONV_KFC_FULL");
  pl.super = pl.tri && !getenv("EINCONV_KFC_TILE128");
  pl.s_tiles = (pl.m_tiles + 1) / 2;
  const long long tiles =
      (pl.super ? (long long)pl.s_tiles * (pl.s_tiles + 1) / 2
                : pl.tri ? (long long)pl.m_tiles * (pl.m_tiles + 1) / 2 : (long long)pl.m_tiles * pl.n_tiles) * g.groups;
  const int stage_bytes = (kTileM + pl.bn) * kRowBytes;
  const int ctas_per_sm = pl.super ? 1 : std::max(1, (227 * 1024) / (kStages * stage_bytes + 2048));
  const long long resident = (long long)num_sms() * ctas_per_sm;
  // Split K so that the CTAs fill whole waves of the resident slots: with `s` splits the main kernel
  // takes ceil(tiles * s / resident) rounds of 1/s of the reduction each, and the finish kernel reads
  // s partial copies.  Pick the s with the least estimated time (300 TFLOP/s per round of full
  // occupancy, 4 TB/s for the partials).
  {
    const long long cap = std::max<long long>(1, std::min<long long>(pl.k_blocks / 8, 512));
    const double k_total_pos = (double)pl.k_blocks * (kRowBytes / es);
    // all tiles, one pass: 128 x 128 tiles run at ~300 TFLOP/s (L2-bound), super-tiles at ~2-3x that
    const double t_main = pl.super ? 2.0 * 256 * 256 * k_total_pos * (double)tiles / (es == 2 ? 900e12 : 450e12)
                                   : 2.0 * kTileM * pl.bn * k_total_pos * (double)tiles / 300e12;
    const double t_partial = (double)g.groups * pl.Mp * pl.Np * 4.0 / 4e12;          // per split copy
    double best_t = 1e30;
    int best_s = 1;
    for (long long s = 1; s <= cap; ++s) {
      const long long ctas = tiles * s;
      const long long rounds = (ctas + resident - 1) / resident;
      // a round processes `resident` CTA-slots of 1/s of the K range each
      // one split writes the factor from the epilogue (KFC); more splits pay the partial copies, the
      // finish kernel's output pass and its launch
      const bool fused = op == EINCONV_OP_KFC && s == 1 && (!pl.super || g.k_total == 1);
      const double t_fin = fused ? 0.0 : t_partial * (double)(s + 1) + 4e-6;
      const double t = t_main * ((double)rounds * resident / (double)ctas) + t_fin + 2e-6 * rounds;
      if (t < best_t) {
        best_t = t;
        best_s = (int)s;
      }
    }
    pl.k_splits = best_s;
    // (measured: raising the split count to 3-4 waves so that the longest-first order can even out unequal
    // super-tiles costs more in partial traffic than it gains: 21.3 -> 22.6 ms per ResNet-50 sweep)
  }
  if (pl.m_tiles >= 65536 || pl.n_tiles >= 65536 || (long long)g.groups * pl.k_splits >= 65536) return pl;

  // which residues (mod A elements = 16 bytes) do the innermost tap offsets t = kw*dil - pad have?
  const int A = 16 / es;
  pl.A = A;
  for (int r = 0; r < 16; ++r) pl.res_map[r] = -1;
  pl.per_tap = q.stride[0] > 1;
  pl.x_direct = !pl.per_tap && (q.in[0] * es) % 16 == 0;
  pl.n_copies = 0;
  int n_maps = 0;
  if (pl.per_tap) {
    if (2 * q.k[0] + 1 >= kMaxMaps) return pl;
    for (int kw = 0; kw < q.k[0]; ++kw) pl.copy_res[pl.n_copies++] = kw * q.dil[0] - q.pad[0];
    n_maps = q.k[0];
    pl.Wp_x = ((q.out[0] + A - 1) / A) * A + A;
  } else {
    for (int kw = 0; kw < q.k[0]; ++kw) {
      const int t = kw * q.dil[0] - q.pad[0];
      const int r = ((t % A) + A) % A;
      if (pl.res_map[r] >= 0) continue;
      pl.res_map[r] = n_maps++;
      if (!(r == 0 && pl.x_direct)) pl.copy_res[pl.n_copies++] = r;
    }
    pl.Wp_x = ((q.in[0] + A + A - 1) / A) * A;
  }
  if (n_maps + q.k[0] + 1 >= kMaxMaps) return pl;
  pl.x_rows = (long long)g.batch * g.groups * g.c_in_g * q.in[2] * q.in[1];
  pl.copy_x_bytes = (((long long)pl.x_rows * pl.Wp_x * es + 255) / 256) * 256;
  if (pl.flat) {
    if (!pl.super) { pl.flat = false; }
  }
  if (pl.super && getenv("EINCONV_KFC_SPLIT")) {  // (measured: no gain; the producer issues uniformly now)
    // boxes per K block of a typical CTA (both operands, two tiles each unless the factor is smaller)
    const int tiles_cta = pl.s_tiles == 1 ? std::min(2, pl.m_tiles) : 4;
    const int boxes = (pl.flat ? 1 : kRowBytes / pl.PB) * tiles_cta * std::min(pl.taps_per_tile, g.k_total);
    while (pl.box_split < 4 && boxes * pl.box_split * 2 <= 32 && pl.CB / (pl.box_split * 2) >= 8) pl.box_split *= 2;
  }
  if (pl.flat) {
    pl.per_tap = false; pl.x_direct = false;
    pl.n_copies = pl.flat_per_tap ? g.k_total : q.k[0];
    const long long planes = (long long)g.batch * g.groups * g.c_in_g;
    pl.copy_x_bytes = ((planes * pl.flat_plane * es + 255) / 256) * 256;
  }
  pl.v_copy = (op == EINCONV_OP_CONV_WEIGHT_VJP) && (q.out[0] * es) % 16 != 0;
  pl.Wp_v = ((q.out[0] + A - 1) / A) * A + A;
  pl.v_rows = (long long)g.batch * g.groups * std::max(1, g.c_out_g) * q.out[2] * q.out[1];
  pl.copy_v_bytes = pl.v_copy ? (((long long)pl.v_rows * pl.Wp_v * es + 255) / 256) * 256 : 0;
  pl.ok = encode_fn() != nullptr;
  return pl;
}

bool tma_supported(int op, const DevGeom& g, int dtype) {
  if (op != EINCONV_OP_CONV_WEIGHT_VJP && op != EINCONV_OP_KFC) return false;
  if (g.batch == 0 || g.out_total == 0 || g.c_in_g == 0) return false;
  if (op == EINCONV_OP_CONV_WEIGHT_VJP && g.c_out_g == 0) return false;
  return make_plan(op, g, dtype).ok;
}

constexpr int64_t kMapBytes = kMaxMaps * 128;  // CUtensorMaps at the head of the workspace

// workspace layout: [tensor maps][shifted copies of x][aligned copy of v][fp32 partial tiles]
int64_t tma_workspace_bytes(int op, const DevGeom& g, int dtype) {
  const Plan pl = make_plan(op, g, dtype);
  return kMapBytes + pl.n_copies * pl.copy_x_bytes + pl.copy_v_bytes +
         (int64_t)g.groups * pl.k_splits * pl.Mp * pl.Np * 4;
}

static int encode_map(CUtensorMap* map, int dtype, const void* base, const int extent[5],
                      const int box[5], const int estride[5], int line_bytes, int pitch_w = 0,
                      long long batch_stride_elems = 0) {
  const int es = (int)elem_size(dtype);
  CUtensorMapDataType dt = dtype == EINCONV_F32    ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                           : dtype == EINCONV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t dims[5], strides[4];
  cuuint32_t bx[5], st[5];
  unsigned long long acc = (unsigned long long)es;
  for (int i = 0; i < 5; ++i) {
    dims[i] = (cuuint64_t)extent[i];
    if (i > 0) strides[i - 1] = acc;
    // the innermost extent may be clipped (zero fill beyond it) while rows keep their pitch
    acc *= (unsigned long long)((i == 0 && pitch_w > 0) ? pitch_w : extent[i]);
    bx[i] = (cuuint32_t)box[i];
    st[i] = (cuuint32_t)estride[i];
  }
  // samples may be further apart than their extent (blocks of gathered window sums, api.cu)
  if (batch_stride_elems > 0) strides[3] = (unsigned long long)batch_stride_elems * es;
  CUtensorMapSwizzle swz = line_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                           : line_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
#ifdef EINCONV_BRINGUP  // unchecked overrides that corrupt results: compiled out of product builds
  if (const char* e = getenv("EINCONV_TMA_SWZ")) swz = (CUtensorMapSwizzle)atoi(e);
  if (const char* e = getenv("EINCONV_TMA_L2P")) l2p = (CUtensorMapL2promotion)atoi(e);
#endif
  CUresult r = encode_fn()(map, dt, 5, const_cast<void*>(base), dims, strides, bx, st,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return EINCONV_ERR_CUDA;
  }
  return EINCONV_OK;
}

template <typename T, int BN>
static int launch(const MapArray& maps, const Params& p, const Plan& pl, int groups, cudaStream_t stream) {
  using S = Smem<BN>;
  auto kern = tma_reduce_gemm_kernel<T, BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tma_reduce_gemm)");
  dim3 grid(pl.m_tiles, pl.n_tiles, groups * pl.k_splits);
  if (pl.tri) grid = dim3(pl.m_tiles * (pl.m_tiles + 1) / 2, 1, groups * pl.k_splits);
  kern<<<grid, kThreads, S::total, stream>>>(maps, p);
  EINCONV_CHECK_LAUNCH("tma_reduce_gemm_kernel");
  return EINCONV_OK;
}

int tma_run_typed_out(int op, const void* x, const void* v, void* out, int dtype, int out_dtype, const DevGeom& g,
                      double scale, void* workspace, int64_t workspace_bytes, cudaStream_t stream,
                      const float* scale_dev, long long x_batch_stride, const float* x_f32,
                      const unsigned* absmax_bits, float* inv_s2) {
  const Plan pl = make_plan(op, g, dtype);
  EINCONV_CHECK_ARG(pl.ok, "tma path not available for this geometry");
  // fp32 source converted by the flat copy itself: only where nothing else reads x (see flat_reads_copies_only)
  EINCONV_CHECK_ARG(!x_f32 || (pl.flat && dtype == EINCONV_F16 && absmax_bits && inv_s2),
                    "converting flat copy needs the flat-plane fp16 plan");
  if (x_batch_stride > 0) {
    EINCONV_CHECK_ARG(pl.x_direct && pl.n_copies == 0 && (x_batch_stride * (long long)elem_size(dtype)) % 16 == 0,
                      "strided samples need a directly readable operand (16-byte multiples)");
  }
  const int64_t need = tma_workspace_bytes(op, g, dtype);
  if (workspace_bytes < need || !workspace) {
    set_error("workspace of %lld bytes required, got %lld", (long long)need, (long long)workspace_bytes);
    return EINCONV_ERR_WORKSPACE;
  }
  const Geo3 q = geo3(g);
  const bool kfc = op == EINCONV_OP_KFC;

  EINCONV_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const int es = (int)elem_size(dtype);
  const int A = pl.A;
  char* ws = (char*)workspace;
  char* copies_x = ws + kMapBytes;
  char* copy_v = copies_x + pl.n_copies * pl.copy_x_bytes;
  float* partial = (float*)(copy_v + pl.copy_v_bytes);

  // ---- pre-pass: aligned shifted copies -------------------------------------------------------
  auto shift_copy = [&](const void* src, void* dst, long long rows, int W, int Wp, int r, int sw) -> int {
    const long long units = rows * (Wp / A);
    const unsigned blocks = (unsigned)((units + 255) / 256);
    if (es == 2)
      shift_copy_kernel<uint16_t><<<blocks, 256, 0, stream>>>((const uint16_t*)src, (uint16_t*)dst, rows, W, Wp, r, sw);
    else
      shift_copy_kernel<uint32_t><<<blocks, 256, 0, stream>>>((const uint32_t*)src, (uint32_t*)dst, rows, W, Wp, r, sw);
    EINCONV_CHECK_LAUNCH("shift_copy_kernel");
    return EINCONV_OK;
  };
  if (pl.flat) {
    FlatCopyParams fp{};
    fp.W = q.in[0]; fp.H = q.in[1];
    fp.pitch = pl.flat_pitch; fp.rows = pl.flat_rows;
    fp.plane_stride = pl.flat_plane;
    fp.copy_stride = pl.copy_x_bytes / es;
    fp.n_copies = pl.n_copies; fp.per_tap = pl.flat_per_tap ? 1 : 0;
    fp.k_w = q.k[0];
    fp.s_h = q.stride[1]; fp.s_w = q.stride[0]; fp.d_h = q.dil[1]; fp.d_w = q.dil[0];
    fp.p_h = q.pad[1]; fp.p_w = q.pad[0]; fp.out_w = q.out[0];
    fp.planes = (long long)g.batch * g.groups * g.c_in_g;
    const long long per_copy = fp.planes * (fp.plane_stride / A);
    EINCONV_CHECK_ARG(per_copy < (1ll << 31), "kfc: flat copy of more than 2^31 units");
    fp.per_copy32 = (uint32_t)per_copy;
    fp.units_div = FastDiv((uint32_t)(fp.plane_stride / A));
    fp.upr_div = FastDiv((uint32_t)(fp.pitch / A));
    const dim3 blocks((unsigned)((per_copy + 255) / 256), (unsigned)fp.n_copies);
    if (x_f32)
      flat_copy_kernel<uint16_t, true><<<blocks, 256, 0, stream>>>(x_f32, (uint16_t*)copies_x, fp, absmax_bits, inv_s2);
    else if (es == 2)
      flat_copy_kernel<uint16_t, false><<<blocks, 256, 0, stream>>>(x, (uint16_t*)copies_x, fp, nullptr, nullptr);
    else
      flat_copy_kernel<uint32_t, false><<<blocks, 256, 0, stream>>>(x, (uint32_t*)copies_x, fp, nullptr, nullptr);
    EINCONV_CHECK_LAUNCH("flat_copy_kernel");
  }
  for (int i = 0; i < (pl.flat ? 0 : pl.n_copies); ++i) {
    int st = shift_copy(x, copies_x + i * pl.copy_x_bytes, pl.x_rows, q.in[0], pl.Wp_x, pl.copy_res[i],
                        pl.per_tap ? q.stride[0] : 1);
    if (st) return st;
  }
  if (pl.v_copy) {
    int st = shift_copy(v, copy_v, pl.v_rows, q.out[0], pl.Wp_v, 0, 1);
    if (st) return st;
  }

  // ---- tensor maps ------------------------------------------------------------------------------
  alignas(64) MapArray map_array;
  memset(&map_array, 0, sizeof(map_array));
  CUtensorMap* host_maps = map_array.m;
  int n_maps = 0, copy_i = 0;
  int res_margin[16] = {0};  // coordinate margin of the map serving residue r: A for copies, 0 for x itself
  signed char flat_map_a[kMaxMaps] = {0}, flat_map_b[kMaxMaps] = {0};
  int flat_c0[kMaxMaps] = {0};
  if (pl.flat) {
    const int chans = g.groups * g.c_in_g;
    const int box[5] = {pl.BW, 1, 1, pl.CB / pl.box_split, 1};
    const int est[5] = {1, 1, 1, 1, 1};
    const int plane = (int)pl.flat_plane;
    if (pl.flat_per_tap) {
      // one copy per tap, every tap reads its copy at the output position itself; rows >= O_h do not exist
      for (int t = 0; t < g.k_total; ++t) {
        const void* base = copies_x + (long long)t * pl.copy_x_bytes;
        const int extent[5] = {plane, 1, 1, chans, g.batch};
        int st = encode_map(&host_maps[n_maps], dtype, base, extent, box, est, pl.PB, plane);
        if (st) return st;
        flat_map_a[t] = flat_map_b[t] = (signed char)n_maps++;
        flat_c0[t] = 0;
      }
    } else {
      // one copy per kernel column; tap (kh, kw) starts kh * d_h rows further down.  Row operand: the whole
      // plane.  Column operand: cut off after the last output row as seen through this kh (zero fill).
      for (int kw = 0; kw < q.k[0]; ++kw) {
        const void* base = copies_x + (long long)kw * pl.copy_x_bytes;
        const int extent[5] = {plane, 1, 1, chans, g.batch};
        int st = encode_map(&host_maps[n_maps++], dtype, base, extent, box, est, pl.PB, plane);
        if (st) return st;
      }
      for (int kh = 0; kh < q.k[1]; ++kh)
        for (int kw = 0; kw < q.k[0]; ++kw) {
          const int t = kh * q.k[0] + kw;
          const void* base = copies_x + (long long)kw * pl.copy_x_bytes;
          const int clipped = std::min(plane, (q.out[1] + kh * q.dil[1]) * pl.flat_pitch);
          const int extent[5] = {clipped, 1, 1, chans, g.batch};
          EINCONV_CHECK_ARG(n_maps < kMaxMaps, "kfc: too many tensor maps");
          int st = encode_map(&host_maps[n_maps], dtype, base, extent, box, est, pl.PB, plane);
          if (st) return st;
          flat_map_a[t] = (signed char)kw;
          flat_map_b[t] = (signed char)n_maps++;
          flat_c0[t] = kh * q.dil[1] * pl.flat_pitch;
        }
    }
  } else if (pl.per_tap) {
    for (int kw = 0; kw < q.k[0]; ++kw) {
      const void* base = copies_x + (long long)kw * pl.copy_x_bytes;
      const int extent[5] = {pl.Wp_x, q.in[1], q.in[2], g.groups * g.c_in_g, g.batch};
      const int box[5] = {pl.BW, 1, 1, pl.CB / pl.box_split, 1};
      const int est[5] = {1, 1, 1, 1, 1};
      int st = encode_map(&host_maps[n_maps++], dtype, base, extent, box, est, pl.PB);
      if (st) return st;
    }
  } else {
    bool seen[16] = {false};
    for (int kw = 0; kw < q.k[0]; ++kw) {
      const int t = kw * q.dil[0] - q.pad[0];
      const int r = ((t % A) + A) % A;
      if (seen[r]) continue;
      seen[r] = true;
      const bool direct = (r == 0 && pl.x_direct);
      const void* base = direct ? x : (const void*)(copies_x + (copy_i++) * pl.copy_x_bytes);
      const int Wext = direct ? q.in[0] : pl.Wp_x;
      res_margin[r] = direct ? 0 : A;
      const int extent[5] = {Wext, q.in[1], q.in[2], g.groups * g.c_in_g, g.batch};
      const int box[5] = {pl.BW, 1, 1, pl.CB / pl.box_split, 1};
      const int est[5] = {1, 1, 1, 1, 1};
      int st = encode_map(&host_maps[n_maps++], dtype, base, extent, box, est, pl.PB, 0, direct ? x_batch_stride : 0);
      if (st) return st;
    }
  }
  const int v_map = n_maps;
  if (!kfc) {
    const void* base = pl.v_copy ? (const void*)copy_v : v;
    const int Wext = pl.v_copy ? pl.Wp_v : q.out[0];
    const int extent[5] = {Wext, q.out[1], q.out[2], g.groups * g.c_out_g, g.batch};
    const int box[5] = {pl.BW, 1, 1, pl.bn, 1};
    const int est[5] = {1, 1, 1, 1, 1};
    int st = encode_map(&host_maps[n_maps++], dtype, base, extent, box, est, pl.PB);
    if (st) return st;
  }

  Params p{};
  p.n_wt = pl.n_wt; p.n_ht = pl.n_ht; p.n_dt = pl.n_dt;
  p.wt_div = FastDiv((uint32_t)pl.n_wt); p.ht_div = FastDiv((uint32_t)pl.n_ht); p.dt_div = FastDiv((uint32_t)pl.n_dt);
  p.BW = pl.BW; p.BH = pl.BH; p.BD = pl.BD;
  p.PB = pl.PB; p.n_pieces = kRowBytes / pl.PB;
  p.k_blocks = pl.k_blocks; p.k_splits = pl.k_splits;
  for (int i = 0; i < 3; ++i) {
    p.stride[i] = q.stride[i]; p.dil[i] = q.dil[i]; p.pad[i] = q.pad[i]; p.k[i] = q.k[i];
  }
  p.kw_div = FastDiv((uint32_t)q.k[0]); p.kh_div = FastDiv((uint32_t)q.k[1]);
  p.k_total = g.k_total;
  p.CB = pl.CB; p.taps_per_tile = pl.taps_per_tile; p.tap_groups = pl.tap_groups;
  p.c_in_g = g.c_in_g; p.c_out_g = g.c_out_g; p.groups = g.groups;
  p.c_blocks = pl.c_blocks; p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles;
  p.Mp = (int)pl.Mp; p.Np = (int)pl.Np;
  p.is_kfc = kfc ? 1 : 0;
  p.tri = pl.tri ? 1 : 0;
  p.scale_dev = scale_dev;
  p.s_tiles = pl.s_tiles;
  p.box_split = pl.box_split;
  // stage geometry of the super-tile kernel
  p.st_rows = 2 * kTileM; p.st_a_bytes = 2 * kTileM * kRowBytes; p.st_bytes = 2 * p.st_a_bytes; p.st_n = 3;
  if (pl.super && pl.s_tiles == 1 && !getenv("EINCONV_KFC_NO_DEEP")) {
    if (pl.m_tiles == 1) { p.st_rows = kTileM; p.st_a_bytes = kTileM * kRowBytes; }
    p.st_bytes = p.st_a_bytes;            // no column operand: the only super-tile is diagonal
    p.st_n = (3 * 4 * kTileM * kRowBytes) / p.st_bytes;  // 6 or 12
  }
#ifdef EINCONV_BRINGUP
  {
    const char* dbg = getenv("EINCONV_TMA_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
#endif
  for (int kw = 0; kw < q.k[0]; ++kw) {
    if (pl.per_tap) {
      // copy_kw[j] = x[s*(j - A) + t_kw]  =>  x[s*w + t_kw] = copy_kw[w + A]
      p.kw_map[kw] = (signed char)kw;
      p.kw_c0[kw] = A;
      continue;
    }
    const int t = kw * q.dil[0] - q.pad[0];
    const int r = ((t % A) + A) % A;
    const int qq = (t - r) / A;  // floor(t / A)
    p.kw_map[kw] = (signed char)pl.res_map[r];
    // copy_r[j] = x[j - A + r]  =>  x[w + t] = copy_r[w + A*qq + A];  direct map (r == 0): x[w + t] = x[w + A*qq]
    p.kw_c0[kw] = A * qq + res_margin[r];
  }
  if (pl.per_tap) p.stride[0] = 1;  // the copies are already subsampled along the innermost dimension
  p.out_h = q.out[1];
  p.out_d = q.out[2];
  for (int kw = 0; kw < q.k[0]; ++kw) p.kw_map_b[kw] = p.kw_map[kw];
  if (pl.flat) {
    p.flat = 1;
    for (int t = 0; t < g.k_total; ++t) {
      p.tap_map_a[t] = flat_map_a[t]; p.tap_map_b[t] = flat_map_b[t]; p.tap_c0[t] = flat_c0[t];
    }
    for (int i = 0; i < 3; ++i) { p.stride[i] = 1; p.dil[i] = 0; p.pad[i] = 0; }
    p.out_h = 1; p.out_d = 1;
  }
  if (kfc && !pl.flat) {
    // column operand of KFC: positions beyond the output width must read as zero (the row operand
    // may then hold anything there).  Same memory as the row operand's map for this tap, with the
    // innermost extent clipped to the last valid output; rows keep their pitch.
    for (int kw = 0; kw < q.k[0]; ++kw) {
      const int t = kw * q.dil[0] - q.pad[0];
      const int r = ((t % A) + A) % A;
      const bool direct = !pl.per_tap && r == 0 && pl.x_direct;
      const void* base;
      if (pl.per_tap) base = copies_x + (long long)kw * pl.copy_x_bytes;
      else if (direct) base = x;
      else {
        int ci = 0;  // index of the copy that serves residue r
        for (int i = 0; i < pl.n_copies; ++i) if (pl.copy_res[i] == r) ci = i;
        base = copies_x + (long long)ci * pl.copy_x_bytes;
      }
      const int pitch = direct ? q.in[0] : pl.Wp_x;
      const int clipped = std::max(1, std::min(pitch, q.out[0] + p.kw_c0[kw]));
      const int extent[5] = {clipped, q.in[1], q.in[2], g.groups * g.c_in_g, g.batch};
      const int box[5] = {pl.BW, 1, 1, pl.CB / pl.box_split, 1};
      const int est[5] = {1, 1, 1, 1, 1};
      EINCONV_CHECK_ARG(n_maps < kMaxMaps, "kfc: too many tensor maps");
      p.kw_map_b[kw] = (signed char)n_maps;
      int st = encode_map(&host_maps[n_maps++], dtype, base, extent, box, est, pl.PB, pitch, direct ? x_batch_stride : 0);
      if (st) return st;
    }
  }
  p.v_map = v_map;
  p.v_c0 = pl.v_copy ? A : 0;
  p.partial = partial;
  // fused finish: KFC with one split, a 1x1 kernel whose partial rows are the channels in order
  // (one channel block, or full 128-channel blocks), upper-triangular tiling and fp32 output
  // (with several taps the factor's channel-major order makes every epilogue store a lone 4-byte sector write,
  // ~1 sector per clock and SM and serialised with the MMAs: measured 160 us per wave for A = 4608 against
  // 26 us of tensor work; the tile-space finish kernel does the same scatter from all SMs at once)
  const bool direct = kfc && pl.tri && pl.k_splits == 1 && out_dtype == EINCONV_F32 &&
                      (!pl.super || g.k_total == 1) && !getenv("EINCONV_KFC_NO_DIRECT");
  if (direct) {
    p.direct_out = (float*)out;
    p.direct_scale = (float)scale;
    p.direct_a = g.c_in_g * g.k_total;
  }
  auto run_main = [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    if (pl.super) {
      auto kern = kfc_super_kernel<T>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSuperSmem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(kfc_super)");
      const int n_super = pl.s_tiles * (pl.s_tiles + 1) / 2;
      TileOrder order;
      memset(&order, 0, sizeof(order));
      if (n_super <= kMaxOrdered) {
        // weight = 128 x 128 blocks a super-tile computes; stable sort keeps the row-major order within a class
        std::vector<std::pair<int, int>> w;  // (-weight, id)
        int id = 0;
        for (int i = 0; i < pl.s_tiles; ++i)
          for (int j = i; j < pl.s_tiles; ++j, ++id) {
            const int rows = (2 * i + 1 < pl.m_tiles) ? 2 : 1, cols = (2 * j + 1 < pl.m_tiles) ? 2 : 1;
            const int blocks = (i == j) ? (rows == 2 ? 3 : 1) : rows * cols;
            w.emplace_back(-blocks, id);
          }
        std::stable_sort(w.begin(), w.end());
        for (int k = 0; k < n_super; ++k) order.t[k] = (unsigned short)w[k].second;
      }
      EINCONV_CHECK_ARG(n_super < 65536, "kfc: too many super-tiles");
      dim3 grid(g.groups * pl.k_splits, 1, n_super);
      kern<<<grid, kThreads, kSuperSmem, stream>>>(map_array, p, order);
      EINCONV_CHECK_LAUNCH("kfc_super_kernel");
      return EINCONV_OK;
    }
    if (pl.bn == 256) return launch<T, 256>(map_array, p, pl, g.groups, stream);
    if (pl.bn == 128) return launch<T, 128>(map_array, p, pl, g.groups, stream);
    return launch<T, 64>(map_array, p, pl, g.groups, stream);
  };
  int st;
  switch (dtype) {
    case EINCONV_F32: st = run_main((float*)nullptr); break;
    case EINCONV_BF16: st = run_main((__nv_bfloat16*)nullptr); break;
    case EINCONV_F16: st = run_main((__half*)nullptr); break;
    default: set_error("tma path: unsupported dtype %d", dtype); return EINCONV_ERR_UNSUPPORTED;
  }
  if (st) return st;
  if (direct) return EINCONV_OK;
  const long long Atot = (long long)g.c_in_g * g.k_total;
  const long long total = (kfc ? Atot * Atot : (long long)g.c_out_g * Atot) * g.groups;
  auto run_finish = [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    if (kfc && pl.tri && g.k_total > 1 && !getenv("EINCONV_KFC_FINISH_TILE") && !getenv("EINCONV_KFC_FINISH_ELEMENTWISE")) {
      // several taps: output-space blocks (whole factor rows) of BC channels
      const int Kt = g.k_total;
      auto smem_of = [&](int bc) {  // block + fragment table
        return (size_t)bc * Kt * ((size_t)bc * Kt + 1) * sizeof(float) + (size_t)Kt * Kt * 8 + 8;
      };
      auto ctas_of = [&](int bc) {
        const long long nb = (g.c_in_g + bc - 1) / bc;
        return nb * (nb + 1) / 2 * g.groups;
      };
      // measured at the 8-rank shard size (one split): BC = 16 A = 4608 139 -> 101 us; blocks of few channels
      // leave too few CTAs (C = 64: 36 CTAs, 57 us against 10 us for the tile-space kernel), so the block kernel
      // is used only where at least two CTAs per SM exist
      int bc = 8;
      if (const char* e = getenv("EINCONV_KFC_FINISH_BC")) bc = atoi(e) == 16 ? 16 : 8;
      if (pl.CB % bc == 0 && smem_of(bc) <= 200 * 1024 && ctas_of(bc) >= 2ll * num_sms() &&
          (long long)pl.Mp * pl.Np < (1ll << 31)) {  // 32-bit fragment offsets
        const int nb = (g.c_in_g + bc - 1) / bc;
        dim3 grid((unsigned)(nb * (nb + 1) / 2), 1, (unsigned)g.groups);
        const size_t smem = smem_of(bc);
        cudaError_t e;
        if (bc == 16) {
          e = cudaFuncSetAttribute(kfc_finish_block_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(kfc_finish_block)");
          kfc_finish_block_kernel<T, 16><<<grid, 256, smem, stream>>>(p.partial, (T*)out, p, (float)scale, nb);
        } else {
          e = cudaFuncSetAttribute(kfc_finish_block_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(kfc_finish_block)");
          kfc_finish_block_kernel<T, 8><<<grid, 256, smem, stream>>>(p.partial, (T*)out, p, (float)scale, nb);
        }
        EINCONV_CHECK_LAUNCH("kfc_finish_block_kernel");
        return EINCONV_OK;
      }
    }
    if (kfc && pl.tri && !getenv("EINCONV_KFC_FINISH_ELEMENTWISE")) {
      const long long n_tri = (long long)pl.m_tiles * (pl.m_tiles + 1) / 2 * g.groups;
      int rows = 8;  // rows per CTA: fewer when there are few tiles, so that >= ~4 CTAs per SM exist
      while (rows > 1 && n_tri * (kTileM / rows) < 4ll * num_sms()) rows /= 2;
      dim3 grid((unsigned)(pl.m_tiles * (pl.m_tiles + 1) / 2), (unsigned)(kTileM / rows), (unsigned)g.groups);
      kfc_finish_tri_kernel<T><<<grid, 256, 0, stream>>>(p.partial, (T*)out, p, (float)scale, rows);
      EINCONV_CHECK_LAUNCH("kfc_finish_tri_kernel");
      return EINCONV_OK;
    }
    tma_finish_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p.partial, (T*)out, p, (float)scale);
    EINCONV_CHECK_LAUNCH("tma_finish_kernel");
    return EINCONV_OK;
  };
  switch (out_dtype) {
    case EINCONV_F32: return run_finish((float*)nullptr);
    case EINCONV_BF16: return run_finish((__nv_bfloat16*)nullptr);
    case EINCONV_F16: return run_finish((__half*)nullptr);
    default: set_error("tma path: unsupported output dtype %d", out_dtype); return EINCONV_ERR_UNSUPPORTED;
  }
}

bool flat_reads_copies_only(int op, const DevGeom& g, int dtype) {
  const Plan pl = make_plan(op, g, dtype);
  return pl.ok && pl.flat;
}

int tma_run(int op, const void* x, const void* v, void* out, int dtype, const DevGeom& g, double scale,
            void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return tma_run_typed_out(op, x, v, out, dtype, dtype, g, scale, workspace, workspace_bytes, stream);
}

}  // namespace tma
}  // namespace einconv
