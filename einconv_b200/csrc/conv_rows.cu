// Stride-1 2-d convolutions straight from the channels-FIRST tensor: no packed copy, no one-hot pattern.
//
//   y[n, co, h, w] = sum_{ci, kh, kw} x[n, ci, h + kh - ph, w + kw - pw] * W[co, ci, kh, kw]
//   (expressions/convNd_forward.py:48-55 with the index law i = o + k - p, conv_index_pattern.py:131-134)
//
// conv_tma.cu needs a zero-padded channels-LAST copy because its A operand has the channels contiguous; the
// copy is a third of the C1 step and its N = 64 MMAs are shared-memory bound (36 of them take 2100 cycles).
// Here the positions stay on the contiguous axis (MN-major A operand, tf32 through the 32-byte-atom swizzle,
// see conv_dense.cu) and the two kinds of tap shift are taken apart:
//   * kh (a shift by whole image rows) is a different TMA box of the same tensor: the slab of a tile holds
//     image rows h0 - ph .. h0 + R + KH - 2 as blocks [K channels][BW positions of one row], and the A operand
//     of tap row kh is the four consecutive blocks that start kh rows down.  Rows outside the image are TMA
//     zero fill, which is the padding.
//   * kw (a shift by single elements, which neither a TMA start coordinate nor a shared-memory descriptor can
//     express: both move in 16-byte steps) is not applied to the operand at all.  The weights of the KW taps of
//     a kernel row are stacked along N: one MMA of N = KW * BN computes, for UNSHIFTED input columns w',
//         D_kw[h, w'][co] = sum_{kh, ci} x[ci, h + kh - ph, w'] W[co, ci, kh, kw]      for every kw at once,
//     and the epilogue adds the KW accumulators of neighbouring columns, y[h, w] = sum_kw D_kw[h, w + kw - pw]:
//     a TMEM lane is an input column, so the neighbours are neighbouring LANES and the sum is two warp
//     shuffles per output.  Blocks of a row overlap by KW - 1 (rounded to 16 bytes) columns so that every
//     output finds its neighbours inside one warp; columns outside the image are zero fill again.
// The MMAs are N = 192 for a 3 x 3 kernel on 64 channels: tensor-bound (96 cycles) instead of shared-memory
// bound, 24 per 128 columns instead of 72.
//
// 16-bit tensors: the same with blocks of 32 positions in 64-byte rows (SWIZZLE_64B), so that a block is still the
// lanes of one warp.
//
// Roles (320 threads, persistent CTAs): warp 0 TMA, warp 1 MMA issue, warps 2-9 epilogue.  The packed weights
// Wp[kh][kw * BN + co][ci] (K-major, written by a small pre-pass) stay resident in shared memory.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dispatch.cuh"
#include "tc_ptx.cuh"

namespace einconv {
namespace crows {

using namespace ptx;

#ifndef EINCONV_ROWS_EPI_WARPS
#define EINCONV_ROWS_EPI_WARPS 8
#endif
constexpr int kEpilogueWarps = EINCONV_ROWS_EPI_WARPS;  // multiple of 4: kEpilogueWarps / 4 per TMEM lane quarter
constexpr int kEpiParts = kEpilogueWarps / 4;
constexpr int kThreads = 64 + 32 * kEpilogueWarps;
constexpr int kMaxStages = 6;
constexpr int kMaxKH = 8;
constexpr int kSmemBudget = 216 * 1024;

struct Params {
  int batch, cs, cd;              // samples, source / destination channels
  int H, W, OH, OW, KH, KW, ph, pw;
  int bw, step, nb_w;             // block width (positions per 128 bytes), block step, blocks per image row
  int R, nb_t, n_wtiles;          // tile = R rows x nb_t blocks (R * nb_t blocks = 128 rows of M); tiles per row
  int n_htiles, n_ntiles, bn;     // row tiles per sample, destination-channel tiles, channels per tile (N = KW * bn)
  long long total_tiles;
  int kb_rows, n_kblocks;         // channel rows per pipeline stage of the source, K blocks
  int wk_rows, n_wkblocks, ksplit;  // weight tiles hold 128 bytes of K per row (wk_rows channels = ksplit stages)
  int slab_blocks, block_bytes, stage_bytes, n_stages;
  int w_tile_bytes, w_bytes;      // one [N rows][128 B] weight tile; all KH * n_wkblocks of them
  int mn_layout;
  const void* bias;
  void* dst;
  long long* trace;  // bring-up only (EINCONV_ROWS_TRACE in an -DEINCONV_BRINGUP build): clock64 stamps of CTA 0
};

#define ROWS_TRACE(slot)                                                                             \
  do {                                                                                               \
    if (p.trace && blockIdx.x < (PAIR ? 2 : 1) && lane == 0 && tile - tile_begin + rank < 8)         \
      p.trace[(tile - tile_begin + rank) * 16 + (slot)] = clock64();                                 \
  } while (0)

struct Bars {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
  float bias[256];
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// CTA pair: the load lands in THIS CTA's shared memory, its bytes complete on the barrier `bar_cluster` (a
// shared::cluster address, here the leader's barrier at the same offset)
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                 int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(p)), "r"(rank));
  return raddr;
}

// MN-major operand descriptor: see conv_dense.cu (layout 2: SWIZZLE_128B, 8-row K groups 1024 bytes apart;
// layout 1: SWIZZLE_128B_BASE32B for 32-bit elements, 4-row K groups 512 bytes apart; layout 4: SWIZZLE_64B for
// 16-bit elements in 64-byte rows = 32 positions, one warp's lanes, 8-row K groups 512 bytes apart)
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t start16, uint32_t lbo16, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)(start16 & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)((layout == 2 ? 1024 : 512) >> 4) << 32;  // 8 rows x 128 B; 4 rows x 128 B (base 32 B); 8 rows x 64 B
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// ---- CTA pair (cta_group::2) helpers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// wait on a local barrier that a thread of the peer CTA arrives on
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                      uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr bool TF32 = true; static constexpr int FMT = 2; };
template <> struct Elem<__nv_bfloat16> { static constexpr bool TF32 = false; static constexpr int FMT = 1; };
template <> struct Elem<__half> { static constexpr bool TF32 = false; static constexpr int FMT = 0; };

// PAIR: two CTAs of a cluster share every MMA (cta_group::2, M = 256): each holds its own 128-row A tile and
// HALF of the weight rows, so the resident weights take half the shared memory (room for four stages instead of
// two or three) and an MMA reads 4 + 3 KB per CTA instead of 4 + 6 (the one-CTA kernel sits at the shared-memory
// bandwidth: its N = 192 MMAs take 121-136 cycles instead of 96).
// Position of a CTA's current tile, advanced without divisions (a tile index costs three integer divisions, ~500
// cycles per tile on the epilogue's critical path): class `ntile` is fixed for a CTA, (n, ht, wt) step through
// samples x row tiles x column tiles.  n >= batch marks the dummy tile past the end (CTA pairs, odd tile count).
struct TilePos {
  int ntile, n, ht, wt;
  __device__ __forceinline__ void init(int tile, int per_ntile, int per_sample, int n_wtiles, int n_ntiles) {
    ntile = min(tile / per_ntile, n_ntiles - 1);
    const int t2 = tile - ntile * per_ntile;
    n = t2 / per_sample;
    const int t3 = t2 - n * per_sample;
    ht = t3 / n_wtiles;
    wt = t3 - ht * n_wtiles;
  }
  __device__ __forceinline__ void advance(int steps, int n_htiles, int n_wtiles) {
    for (int s = 0; s < steps; ++s)
      if (++wt == n_wtiles) {
        wt = 0;
        if (++ht == n_htiles) {
          ht = 0;
          ++n;
        }
      }
  }
};

// KWS > 0: kernel width and left padding along W known at compile time (KWS, PWS): the epilogue's exchange becomes
// shuffles by constant distances and the tap of the column itself needs none; KWS = 0: both read from Params.
template <typename T, bool PAIR, int KWS, int PWS>
__global__ void __launch_bounds__(kThreads, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* w_region = smem;
  unsigned char* ring = smem + p.w_bytes;
  Bars* bars = reinterpret_cast<Bars*>(ring + (size_t)p.n_stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ES = (int)sizeof(T);
  constexpr int KSTEP = 32 / ES;  // K per MMA: 16 (16-bit) / 8 (tf32)
  const int n_cols = p.KW * p.bn; // N of the MMAs = TMEM columns of one accumulator
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
#ifdef EINCONV_BRINGUP
  unsigned long long t_entry = 0;
  if (p.trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
#endif
  const int w_rows = PAIR ? n_cols / 2 : n_cols;  // weight rows this CTA holds of every tile

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], PAIR ? 2 * kEpilogueWarps : kEpilogueWarps);  // PAIR: the leader's counts both CTAs
    }
    mbar_init(&bars->w_full, 1);
    fence_barrier_init();
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)n_cols) tmem_cols <<= 1;
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc2(&bars->tmem_base, tmem_cols);
    else tmem_alloc(&bars->tmem_base, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // tiles: index = ((ntile * batch + n) * n_htiles + htile) * n_wtiles + wtile; contiguous chunk per CTA
  // (PAIR: chunk of tile PAIRS per cluster; CTA `rank` takes tile 2 * pair + rank, which may lie past the end:
  // its loads are then all zero fill and its stores masked)
  // (tile indices are 32-bit: the host rejects geometries with 2^30 tiles or more; the 64-bit divisions of the
  // first version cost every warp ~700 cycles per tile)
  const long long n_units = PAIR ? (p.total_tiles + 1) / 2 : p.total_tiles;
  const long long n_workers = PAIR ? gridDim.x / 2 : gridDim.x, worker = PAIR ? blockIdx.x / 2 : blockIdx.x;
  int unit_begin = (int)(n_units * worker / n_workers), unit_end = (int)(n_units * (worker + 1) / n_workers);
  if (!PAIR && p.n_ntiles > 1) {
    // several destination-channel classes (N = KW * bn <= 256 each): a CTA keeps the weights of ONE class resident,
    // so the grid (a multiple of n_ntiles, see the host code) is cut into one group of CTAs per class
    const int per_class = (int)(p.total_tiles / p.n_ntiles), wpc = (int)n_workers / p.n_ntiles;
    const int cls = (int)worker / wpc, l = (int)worker - cls * wpc;
    unit_begin = cls * per_class + (int)((long long)per_class * l / wpc);
    unit_end = cls * per_class + (int)((long long)per_class * (l + 1) / wpc);
  }
  const int tile_step = PAIR ? 2 : 1;
  const int tile_begin = PAIR ? 2 * unit_begin + (int)rank : unit_begin;
  const int tile_end = PAIR ? 2 * unit_end + (int)rank : unit_end;  // exclusive bound of the stepped loop
  const int total_tiles = (int)p.total_tiles;
  const int per_ntile = p.batch * p.n_htiles * p.n_wtiles;
  const int per_sample = p.n_htiles * p.n_wtiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one box per lane)
    int st = 0;
    uint32_t ph = 0;
    int cur_ntile = -1;
    const uint32_t stage_tx = (uint32_t)(p.slab_blocks * p.block_bytes);
    TilePos pos;
    pos.init(tile_begin, per_ntile, per_sample, p.n_wtiles, p.n_ntiles);
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, pos.advance(tile_step, p.n_htiles, p.n_wtiles)) {
      const int ntile = pos.ntile;
      const int n = pos.n;  // >= batch for the dummy tile of an odd count: zero fill
      const int h0 = pos.ht * p.R, wb0 = pos.wt * p.nb_t;
      if (ntile != cur_ntile) {
        // (re)load the weights of this n-tile: KH * n_kblocks tiles of [n_cols rows][128 B of K]
        // (only at the start of a CTA's chunk in practice; a second class would need the MMA warp to be done
        // with the first: chunks are cut so that this cannot happen, see the host code)
        cur_ntile = ntile;
        // the packed weights are written by the preceding grid (launched with programmatic stream
        // serialisation: everything above this line overlaps its tail); no-op without the attribute
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int n_wt = p.KH * p.n_wkblocks;
        // PAIR: only the leader's barriers are waited on; they expect the bytes of both CTAs, and the peer's
        // loads complete on them (cta_group::2), so nothing has to be relayed
        if (lane == 0 && rank == 0) mbar_expect_tx(&bars->w_full, (uint32_t)((PAIR ? 2 : 1) * n_wt * w_rows * 128));
        __syncwarp();
        for (int i = lane; i < n_wt; i += 32) {
          const int kh = i / p.n_wkblocks, kb = i - kh * p.n_wkblocks;
          const int row0 = (min(ntile, p.n_ntiles - 1) * p.KH + kh) * n_cols + (int)rank * w_rows;
          if constexpr (PAIR)
            tma_load_2d_pair(w_region + (size_t)i * p.w_tile_bytes, &map_b, mapa_u32(&bars->w_full, 0), kb * p.wk_rows,
                             row0);
          else
            tma_load_2d(w_region + (size_t)i * p.w_tile_bytes, &map_b, &bars->w_full, kb * p.wk_rows, row0);
        }
      }
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(&bars->empty[st], ph ^ 1);
        if (kb < 2) ROWS_TRACE(kb);
        if (lane == 0 && rank == 0) mbar_expect_tx(&bars->full[st], (PAIR ? 2u : 1u) * stage_tx);
        __syncwarp();
        unsigned char* stage = ring + (size_t)st * p.stage_bytes;
        const uint32_t full_leader = PAIR ? mapa_u32(&bars->full[st], 0) : 0u;
        for (int bx = lane; bx < p.slab_blocks; bx += 32) {
          const int row = bx / p.nb_t, wb = wb0 + bx - row * p.nb_t;
          if constexpr (PAIR)
            tma_load_4d_pair(stage + (size_t)bx * p.block_bytes, &map_a, full_leader, wb * p.step, h0 - p.ph + row,
                             kb * p.kb_rows, n);
          else
            tma_load_4d(stage + (size_t)bx * p.block_bytes, &map_a, &bars->full[st], wb * p.step, h0 - p.ph + row,
                        kb * p.kb_rows, n);
        }
        __syncwarp();
        if (++st == p.n_stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && PAIR && rank == 1) {
    // peer CTA: its MMAs are issued by the leader; this warp only allocates / frees the tensor memory
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (PAIR: the leader CTA)
    const uint32_t idesc = make_idesc(Elem<T>::FMT, PAIR ? 256 : 128, n_cols) | (1u << 15);  // A MN-major, B K-major
    const uint32_t ring16 = smem_u32(ring) >> 4, stage16 = (uint32_t)p.stage_bytes >> 4;
    const uint32_t w16 = smem_u32(w_region) >> 4, wtile16 = (uint32_t)p.w_tile_bytes >> 4;
    const uint32_t blk16 = (uint32_t)p.block_bytes >> 4;
    const uint64_t kdesc = make_desc_kmajor(0, 128);
    const uint32_t kMnStep16 = (uint32_t)(KSTEP * p.bw * ES) >> 4;  // KSTEP rows of one block
    const int ksteps = p.kb_rows / KSTEP;
    const uint64_t layout = (uint64_t)p.mn_layout;
    int st = 0;
    uint32_t ph = 0, it = 0;
    bool w_ready = false;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      if (!w_ready) {
        mbar_wait(&bars->w_full, 0);
        w_ready = true;
      }
      ROWS_TRACE(4);
      if constexpr (PAIR) mbar_wait_cluster(&bars->tmem_empty[acc], acc_phase ^ 1);  // peer warps arrive remotely
      else mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      ROWS_TRACE(5);
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)n_cols;
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(&bars->full[st], ph);
        tc_fence_after();
        if (kb < 2) ROWS_TRACE(6 + 2 * kb);
        if (elect_one()) {
          const uint32_t a16 = ring16 + (uint32_t)st * stage16;
          for (int kh = 0; kh < p.KH; ++kh) {
            // tap row kh: the four blocks that start kh image rows into the slab
            const uint64_t a_base = make_desc_mnmajor(a16 + (uint32_t)(kh * p.nb_t) * blk16, blk16, layout);
            // weight tile of this stage's channels and, inside its 128-byte rows, the stage's part of K
            const int wkb = kb / p.ksplit, part = kb - wkb * p.ksplit;
            const uint64_t b_base = kdesc + (uint64_t)(w16 + (uint32_t)(kh * p.n_wkblocks + wkb) * wtile16 +
                                                       (uint32_t)(part * p.kb_rows * ES) / 16u);
            for (int k = 0; k < ksteps; ++k) {
              if constexpr (PAIR)
                umma2<Elem<T>::TF32>(d_tmem, a_base + (uint64_t)((uint32_t)k * kMnStep16), b_base + (uint64_t)(2 * k), idesc,
                                     (uint32_t)(kb | kh | k));
              else
                umma<Elem<T>::TF32>(d_tmem, a_base + (uint64_t)((uint32_t)k * kMnStep16), b_base + (uint64_t)(2 * k), idesc,
                                    (uint32_t)(kb | kh | k));
            }
          }
          if constexpr (PAIR) {
            umma_commit2(&bars->empty[st]);
            if (kb + 1 == p.n_kblocks) umma_commit2(&bars->tmem_full[acc]);
          } else {
            umma_commit(&bars->empty[st]);
            if (kb + 1 == p.n_kblocks) umma_commit(&bars->tmem_full[acc]);
          }
        }
        __syncwarp();
        if (kb < 2) ROWS_TRACE(7 + 2 * kb);
        if (++st == p.n_stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;      // block `quarter` of the tile: TMEM lanes 32 * quarter .. + 31
    const int half = (warp - 2) >> 2;  // 16-channel chunks half, half + kEpiParts, ..
    const int etid = threadIdx.x - 64;
    T* dst = reinterpret_cast<T*>(p.dst);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const int n_chunks = p.bn >> 4;
    const long long plane = (long long)p.OH * p.OW;
    const int r_local = quarter / p.nb_t, wb_local = quarter - r_local * p.nb_t;
    uint32_t it = 0;
    int cur_ntile = -1;
    TilePos pos;
    pos.init(tile_begin, per_ntile, per_sample, p.n_wtiles, p.n_ntiles);
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it, pos.advance(tile_step, p.n_htiles, p.n_wtiles)) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int ntile = pos.ntile;
      const int n = pos.n;
      const int h = pos.ht * p.R + r_local;
      const int n0 = ntile * p.bn;
      // this thread's input column (= output column) and the outputs its block owns: [lo, hi)
      const int wb = pos.wt * p.nb_t + wb_local;
      const int col = wb * p.step + lane;
      const int lo = wb == 0 ? 0 : wb * p.step + p.pw;
      const int hi = wb + 1 >= p.nb_w ? p.OW : (wb + 1) * p.step + p.pw;
      const bool col_ok = col >= lo && col < hi && col < p.OW && wb < p.nb_w;
      if (bias && ntile != cur_ntile) {
        cur_ntile = ntile;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
        for (int c = etid; c < p.bn; c += 32 * kEpilogueWarps)
          bars->bias[c] = n0 + c < p.cd ? (float)Cvt<T>::load(bias + n0 + c) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
      }
      const bool valid = col_ok && h < p.OH && tile < total_tiles;
      T* dst_row = dst + ((long long)n * p.cd + n0) * plane + (long long)h * p.OW + col;
      if (warp == 2) ROWS_TRACE(10);
      mbar_wait_relaxed(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 2) ROWS_TRACE(11);
      const uint32_t taddr = tmem_base + acc * (uint32_t)n_cols + ((uint32_t)(quarter * 32) << 16);
      for (int ch = half; ch < n_chunks; ch += kEpiParts) {
        const int c0 = ch << 4;
        // all KW accumulators of the chunk in flight before the one wait (the loads are ~150 cycles each when
        // issued back to back with a wait in between: the epilogue, not the MMAs, then bounds a CTA pair)
        constexpr bool SPEC = KWS > 0;
        const int kw_n = SPEC ? KWS : p.KW, pw_v = SPEC ? PWS : p.pw;
        uint32_t r[4][16];
#pragma unroll
        for (int kw = 0; kw < 4; ++kw)
          if (kw < kw_n) tmem_ld16_nowait(taddr + (uint32_t)(kw * p.bn + c0), r[kw]);
        tmem_ld_wait();
        if (ch + kEpiParts >= n_chunks) {
          // this warp's last read of the accumulator: hand it back before the shuffles and stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank == 1) mbar_arrive_remote(&bars->tmem_empty[acc], 0);
            else mbar_arrive(&bars->tmem_empty[acc]);
          }
        }
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = bias ? bars->bias[c0 + j] : 0.f;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          if (kw < kw_n) {
            // y[w] += D_kw[w + kw - pw]: the value of lane + (kw - pw); lanes outside the warp are columns this
            // block does not own outputs for, or padding (zero)
            if constexpr (SPEC) {
              const int delta = kw - PWS;  // a constant after unrolling
              if (delta == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) y[j] += __uint_as_float(r[kw][j]);
              } else if (delta > 0) {
                const bool ok = lane + delta < 32;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float v = __shfl_down_sync(0xffffffffu, __uint_as_float(r[kw][j]), (unsigned)delta);
                  y[j] += ok ? v : 0.f;
                }
              } else {
                const bool ok = lane + delta >= 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float v = __shfl_up_sync(0xffffffffu, __uint_as_float(r[kw][j]), (unsigned)(-delta));
                  y[j] += ok ? v : 0.f;
                }
              }
            } else {
              const int src = lane + kw - pw_v;
              const bool src_ok = src >= 0 && src < 32;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float v = __shfl_sync(0xffffffffu, __uint_as_float(r[kw][j]), src & 31);
                y[j] += src_ok ? v : 0.f;
              }
            }
          }
        }
        if (valid) {
          T* d0 = dst_row + (long long)c0 * plane;
          if (n0 + c0 + 16 <= p.cd) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              Cvt<T>::store(d0, y[j]);
              d0 += plane;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n0 + c0 + j < p.cd) Cvt<T>::store(d0, y[j]);
              d0 += plane;
            }
          }
        }
      }
      if (half >= n_chunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank == 1) mbar_arrive_remote(&bars->tmem_empty[acc], 0);
          else mbar_arrive(&bars->tmem_empty[acc]);
        }
      }
      if (warp == 2) ROWS_TRACE(12);
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef EINCONV_BRINGUP
  if (p.trace && threadIdx.x == 0 && blockIdx.x < 256) {
    unsigned long long t_exit;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
    p.trace[128 + 2 * blockIdx.x] = (long long)t_entry;
    p.trace[129 + 2 * blockIdx.x] = (long long)t_exit;
  }
#endif
  if constexpr (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// Wp[ntile][kh][kw * bn + c][ci] = W[ntile * bn + c][ci][kh][kw]   (zero where the channel does not exist)
// flip (input VJP): Wp[ntile][kh][kw * bn + c][co] = W[co][ntile * bn + c][KH - 1 - kh][KW - 1 - kw]
template <typename T>
__global__ void __launch_bounds__(256) pack_rows_weights_kernel(const T* __restrict__ w, T* __restrict__ wp, int cd,
                                                                int cs, int KH, int KW, int bn, int n_ntiles,
                                                                int cs_pitch, int flip) {
  // programmatic dependent launch: the row kernel may be scheduled now; it waits (griddepcontrol.wait) before
  // its first read of wp
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // 32-bit indices (resident weights are at most a few hundred KB: the host checks total < 2^31)
  const unsigned total = (unsigned)n_ntiles * KH * KW * bn * cs_pitch;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = (int)(idx % (unsigned)cs_pitch);
    unsigned r = idx / (unsigned)cs_pitch;
    const int c = (int)(r % (unsigned)bn);
    r /= (unsigned)bn;
    const int kw = (int)(r % (unsigned)KW);
    r /= (unsigned)KW;
    const int kh = (int)(r % (unsigned)KH);
    const int nt = (int)(r / (unsigned)KH);
    const int co = nt * bn + c;
    T v = T(0);
    if (co < cd && ci < cs)
      v = flip ? w[(((long long)ci * cd + co) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)]
               : w[(((long long)co * cs + ci) * KH + kh) * KW + kw];
    wp[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeFn)ptr;
  }
  return fn;
}

static int encode(CUtensorMap* map, int dtype, const void* base, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box, bool mn_major) {
  EncodeFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return EINCONV_ERR_CUDA;
  }
  CUtensorMapDataType dt = dtype == EINCONV_F32    ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                           : dtype == EINCONV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint32_t est[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, est,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  !mn_major                ? CU_TENSOR_MAP_SWIZZLE_128B
                  : dtype == EINCONV_F32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                         : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (conv_rows, rank %d) failed with CUresult %d", rank, (int)r);
    return EINCONV_ERR_CUDA;
  }
  return EINCONV_OK;
}

struct Plan {
  bool ok = false;
  bool pair = false;  // CTA pairs (cta_group::2): half of the weight rows per CTA
  int es, bw, step, nb_w, R, nb_t, n_wtiles, bn, n_ntiles, n_cols, kb_rows, n_kblocks, cs_pitch;
  int wk_rows, n_wkblocks, ksplit;
  int slab_blocks, block_bytes, stage_bytes, n_stages, w_tile_bytes, w_bytes, n_htiles;
  long long wp_bytes;
  size_t smem;
};

// Both supported ops as a stride-1 forward convolution of a source [n][cs][SH][SW] into a destination
// [n][cd][DH][DW].  The input VJP (expressions/convNd_input_vjp.py:52-60) is the forward convolution of the
// cotangent with the weights flipped along every kernel axis and the channel roles swapped, padded by k - 1 - p.
struct View {
  int cs, cd, SH, SW, DH, DW, ph, pw;
  bool flip;
};
static View make_view(int op, const DevGeom& g) {
  View v;
  if (op == EINCONV_OP_CONV_INPUT_VJP)
    v = View{g.c_out_g, g.c_in_g, g.out[0], g.out[1], g.in[0], g.in[1], g.k[0] - 1 - g.pad[0], g.k[1] - 1 - g.pad[1], true};
  else
    v = View{g.c_in_g, g.c_out_g, g.in[0], g.in[1], g.out[0], g.out[1], g.pad[0], g.pad[1], false};
  return v;
}

static Plan make_plan(int op, const DevGeom& g, int dtype, int math) {
  Plan pl;
  if (op != EINCONV_OP_CONV_FWD && op != EINCONV_OP_CONV_INPUT_VJP) return pl;
  if (getenv("EINCONV_NO_CONV_ROWS")) return pl;
  if (op == EINCONV_OP_CONV_INPUT_VJP && getenv("EINCONV_NO_CONV_ROWS_VJP")) return pl;
  if (dtype == EINCONV_F64) return pl;
  if (dtype == EINCONV_F32 && math != EINCONV_MATH_TF32) return pl;
  if (dtype != EINCONV_F32 && getenv("EINCONV_ROWS_NO_16BIT")) return pl;
  if (g.n_dims != 2 || g.groups != 1 || g.batch == 0) return pl;
  for (int d = 0; d < 2; ++d)
    if (g.stride[d] != 1 || g.dil[d] != 1) return pl;
  const View vw = make_view(op, g);
  const int KH = g.k[0], KW = g.k[1], W = vw.SW;
  if (KW < 2 || KW > 4 || KH > kMaxKH || g.pad[1] > KW - 1 || g.pad[0] > KH - 1) return pl;  // 1 x 1: conv_dense.cu
  // stride 1: every source element is covered, so the two sizes determine each other
  if (vw.DH != vw.SH + 2 * vw.ph - KH + 1 || vw.DW != vw.SW + 2 * vw.pw - KW + 1) return pl;
  const int es = (int)elem_size(dtype);
  pl.es = es;
  if (((long long)W * es) % 16 != 0 || vw.cs == 0 || vw.cd == 0) return pl;
  pl.bw = 32;  // positions per block = the lanes of one warp: 128-byte rows (fp32) / 64-byte rows (16-bit, SWIZZLE_64B)
  const int halo = ((KW - 1) * es + 15) / 16 * 16 / es;  // overlap of neighbouring blocks, whole 16-byte units
  pl.step = pl.bw - halo;
  pl.nb_w = W <= pl.bw ? 1 : (W - pl.bw + pl.step - 1) / pl.step + 1;
  // tile = R image rows x nb_t blocks of a row, R * nb_t = 4 blocks (M = 128): fewest tiles, then the smallest
  // slab ((R + KH - 1) * nb_t blocks: tall tiles re-read fewer halo rows and leave room for a third stage)
  {
    long long best = -1;
    for (int nbt : {1, 2, 4}) {
      const int R = 4 / nbt;
      const long long tiles = (long long)((vw.DH + R - 1) / R) * ((pl.nb_w + nbt - 1) / nbt);
      const long long cost = tiles * 64 + (long long)(R + KH - 1) * nbt;
      if (best < 0 || cost < best) {
        best = cost;
        pl.nb_t = nbt;
        pl.R = R;
      }
    }
    pl.n_wtiles = (pl.nb_w + pl.nb_t - 1) / pl.nb_t;
  }
  // destination channels per tile: N = KW * bn <= 256
  const int bn_max = 256 / KW / 16 * 16;
  pl.n_ntiles = (vw.cd + bn_max - 1) / bn_max;
  pl.bn = ((vw.cd + pl.n_ntiles - 1) / pl.n_ntiles + 15) / 16 * 16;
  pl.n_cols = KW * pl.bn;
  pl.wk_rows = 128 / es;                              // 128 bytes of K per weight-tile row
  pl.n_wkblocks = (vw.cs + pl.wk_rows - 1) / pl.wk_rows;
  pl.cs_pitch = pl.n_wkblocks * pl.wk_rows;
  // CTA pairs: measured slower on B200 (C1 28.6 us per call against 22.6 us): the MMAs reach their full rate but
  // each SM is then bound by the epilogue's TMEM reads (see the kernel's header comment), so pairs are opt-in
  pl.pair = (pl.n_cols / 2) % 8 == 0 && num_sms() % 2 == 0 && pl.n_ntiles == 1 && getenv("EINCONV_ROWS_PAIR") != nullptr;
  pl.w_tile_bytes = ((pl.pair ? pl.n_cols / 2 : pl.n_cols) * 128 + 1023) / 1024 * 1024;
  pl.w_bytes = KH * pl.n_wkblocks * pl.w_tile_bytes;
  // source stages: a whole weight K block (32 fp32 channels) per stage; EINCONV_ROWS_KSPLIT=2 halves them (C1: four
  // 16 KB stages instead of two of 32 KB next to the resident weights).  Measured neutral (24.5 us per call either
  // way): the MMA warp is not waiting for loads
  pl.ksplit = 1;
  if (const char* e = getenv("EINCONV_ROWS_KSPLIT")) pl.ksplit = atoi(e) == 2 ? 2 : 1;
  pl.kb_rows = pl.wk_rows / pl.ksplit;
  pl.n_kblocks = (vw.cs + pl.kb_rows - 1) / pl.kb_rows;
  pl.block_bytes = pl.kb_rows * pl.bw * es;
  pl.slab_blocks = (pl.R + KH - 1) * pl.nb_t;
  pl.stage_bytes = pl.slab_blocks * pl.block_bytes;
  if (pl.w_bytes + 2 * pl.stage_bytes > kSmemBudget) return pl;
  pl.n_stages = std::min(kMaxStages, (kSmemBudget - pl.w_bytes) / pl.stage_bytes);
  pl.n_htiles = (vw.DH + pl.R - 1) / pl.R;
  pl.wp_bytes = ((long long)pl.n_ntiles * KH * pl.n_cols * pl.cs_pitch * es + 1023) / 1024 * 1024;
  pl.smem = (size_t)pl.w_bytes + (size_t)pl.n_stages * pl.stage_bytes + sizeof(Bars) + 64;
  // a CTA's chunk of tiles stays inside one n-tile class (resident weights are loaded once): one group of CTAs
  // per class (the pair variant does not implement that: see pl.pair)
  if (pl.n_ntiles > num_sms()) return pl;
  if ((long long)g.batch * pl.n_htiles * pl.n_wtiles >= (1ll << 30)) return pl;  // 32-bit tile indices in the kernel
  pl.ok = true;
  return pl;
}

bool supported(int op, const DevGeom& g, int dtype, int math) { return make_plan(op, g, dtype, math).ok; }

int64_t workspace_bytes(int op, const DevGeom& g, int dtype, int math) {
  const Plan pl = make_plan(op, g, dtype, math);
  return pl.ok ? pl.wp_bytes + 1024 : 0;
}

template <typename T>
static int run_t(int op, const Plan& pl, const void* x, const void* w, const void* bias, void* y, int dtype,
                 const DevGeom& g, void* workspace, cudaStream_t stream) {
  const int es = pl.es;
  const View vw = make_view(op, g);
  uintptr_t base = ((uintptr_t)workspace + 1023) & ~(uintptr_t)1023;
  T* wp = reinterpret_cast<T*>(base);
  const int KH = g.k[0], KW = g.k[1];
  {
    const long long total = (long long)pl.n_ntiles * KH * KW * pl.bn * pl.cs_pitch;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 4 * num_sms());
#ifdef EINCONV_BRINGUP
    if (getenv("EINCONV_ROWS_PACK_CARVE"))
      cudaFuncSetAttribute(pack_rows_weights_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    static int calls = 0;
    if (!(getenv("EINCONV_ROWS_SKIP_PACK") && calls++ > 0))
#endif
    pack_rows_weights_kernel<T><<<blocks, 256, 0, stream>>>((const T*)w, wp, vw.cd, vw.cs, KH, KW, pl.bn, pl.n_ntiles,
                                                            pl.cs_pitch, vw.flip ? 1 : 0);
    EINCONV_CHECK_LAUNCH("pack_rows_weights_kernel");
  }
  Params p{};
  p.batch = g.batch;
  p.cs = vw.cs;
  p.cd = vw.cd;
  p.H = vw.SH; p.W = vw.SW; p.OH = vw.DH; p.OW = vw.DW;
  p.KH = KH; p.KW = KW; p.ph = vw.ph; p.pw = vw.pw;
  p.bw = pl.bw; p.step = pl.step; p.nb_w = pl.nb_w; p.R = pl.R; p.nb_t = pl.nb_t; p.n_wtiles = pl.n_wtiles;
  p.n_htiles = pl.n_htiles; p.n_ntiles = pl.n_ntiles; p.bn = pl.bn;
  p.total_tiles = (long long)pl.n_ntiles * g.batch * pl.n_htiles * pl.n_wtiles;
  p.kb_rows = pl.kb_rows; p.n_kblocks = pl.n_kblocks;
  p.wk_rows = pl.wk_rows; p.n_wkblocks = pl.n_wkblocks; p.ksplit = pl.ksplit;
  p.slab_blocks = pl.slab_blocks; p.block_bytes = pl.block_bytes; p.stage_bytes = pl.stage_bytes;
  p.n_stages = pl.n_stages; p.w_tile_bytes = pl.w_tile_bytes; p.w_bytes = pl.w_bytes;
  p.mn_layout = es == 4 ? 1 : 4;
  p.bias = bias;
  p.dst = y;

  CUtensorMap map_a, map_b;
  {
    // source [n][ci][h][w]
    cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.cs, (cuuint64_t)p.batch};
    cuuint64_t strides[3] = {(cuuint64_t)p.W * es, (cuuint64_t)p.W * p.H * es, (cuuint64_t)p.W * p.H * p.cs * es};
    cuuint32_t box[4] = {(cuuint32_t)pl.bw, 1, (cuuint32_t)pl.kb_rows, 1};
    int st = encode(&map_a, dtype, x, 4, dims, strides, box, true);
    if (st) return st;
  }
  {
    // packed weights [ntile * KH * n_cols rows][cs_pitch]
    cuuint64_t dims[2] = {(cuuint64_t)pl.cs_pitch, (cuuint64_t)pl.n_ntiles * KH * pl.n_cols};
    cuuint64_t strides[1] = {(cuuint64_t)pl.cs_pitch * es};
    cuuint32_t box[2] = {(cuuint32_t)pl.wk_rows, (cuuint32_t)(pl.pair ? pl.n_cols / 2 : pl.n_cols)};
    int st = encode(&map_b, dtype, wp, 2, dims, strides, box, false);
    if (st) return st;
  }
  const bool spec31 = KW == 3 && vw.pw == 1 && !getenv("EINCONV_ROWS_NO_SPEC");
  auto kern = spec31 ? (pl.pair ? conv_rows_kernel<T, true, 3, 1> : conv_rows_kernel<T, false, 3, 1>)
                     : (pl.pair ? conv_rows_kernel<T, true, 0, 0> : conv_rows_kernel<T, false, 0, 0>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_rows)");
  unsigned grid = pl.pair ? 2u * (unsigned)std::min<long long>((p.total_tiles + 1) / 2, num_sms() / 2)
                          : (unsigned)std::min<long long>(p.total_tiles, num_sms());
  if (pl.n_ntiles > 1) grid = grid / pl.n_ntiles * pl.n_ntiles;  // equal groups of CTAs per class (grid >= n_ntiles)
#ifdef EINCONV_BRINGUP  // allocates, synchronises and frees: never inside a product build
  long long* trace = nullptr;
  if (getenv("EINCONV_ROWS_TRACE")) {
    cudaMalloc(&trace, (8 * 16 + 512) * sizeof(long long));
    cudaMemset(trace, 0, (8 * 16 + 512) * sizeof(long long));
    p.trace = trace;
  }
#endif
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (pl.pair) {
      attr[n_attr].id = cudaLaunchAttributeClusterDimension;
      attr[n_attr].val.clusterDim.x = 2;
      attr[n_attr].val.clusterDim.y = 1;
      attr[n_attr].val.clusterDim.z = 1;
      ++n_attr;
    }
    if (getenv("EINCONV_ROWS_PDL")) {
      // overlap this kernel's launch, barrier / TMEM set-up and tensor-map fetch with the weight pre-pass
      // (measured: no gain under graph replay, 22.6 us either way, so opt-in)
      attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
      ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    cudaError_t el = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, p);
    if (el != cudaSuccess) return cuda_fail(el, "cudaLaunchKernelEx(conv_rows)");
  }
  EINCONV_CHECK_LAUNCH("conv_rows_kernel");
#ifdef EINCONV_BRINGUP
  if (trace) {
    long long h[8 * 16 + 512];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(trace);
    {
      long long first = 0, last = 0;
      for (unsigned c = 0; c < grid && c < 256; ++c) {
        if (!first || h[128 + 2 * c] < first) first = h[128 + 2 * c];
        if (h[129 + 2 * c] > last) last = h[129 + 2 * c];
      }
      fprintf(stderr, "conv_rows CTAs (globaltimer ns): first entry -> last exit %lld ns; per CTA entry offset / duration:\n",
              last - first);
      for (unsigned c = 0; c < grid && c < 256; ++c)
        fprintf(stderr, "%u:%lld/%lld%s", c, h[128 + 2 * c] - first, h[129 + 2 * c] - h[128 + 2 * c], c % 8 == 7 ? "\n" : "  ");
      fprintf(stderr, "\n");
    }
    long long t0 = 0;
    for (int i = 0; i < 8 * 16; ++i)
      if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
    fprintf(stderr, "conv_rows trace (cycles since first stamp)\n");
    fprintf(stderr, "tile  P:st0   P:st1 | M:start M:tmem  M:full0 M:iss0  M:full1 M:iss1 | E:start E:full  E:done\n");
    for (int t = 0; t < 8; ++t) {
      fprintf(stderr, "%4d ", t);
      for (int k : {0, 1, 4, 5, 6, 7, 8, 9, 10, 11, 12}) fprintf(stderr, "%7lld ", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
#endif
  return EINCONV_OK;
}

int run(int op, const void* x, const void* w, const void* bias, void* y, int dtype, const DevGeom& g, int math,
        void* workspace, int64_t workspace_bytes_given, cudaStream_t stream) {
  const Plan pl = make_plan(op, g, dtype, math);
  if (!pl.ok) {
    set_error("conv_rows: unsupported geometry");
    return EINCONV_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes_given < pl.wp_bytes + 1024) {
    set_error("conv_rows: workspace too small");
    return EINCONV_ERR_WORKSPACE;
  }
  switch (dtype) {
    case EINCONV_F32: return run_t<float>(op, pl, x, w, bias, y, dtype, g, workspace, stream);
    case EINCONV_BF16: return run_t<__nv_bfloat16>(op, pl, x, w, bias, y, dtype, g, workspace, stream);
    case EINCONV_F16: return run_t<__half>(op, pl, x, w, bias, y, dtype, g, workspace, stream);
    default: set_error("conv_rows: unsupported dtype"); return EINCONV_ERR_UNSUPPORTED;
  }
}

}  // namespace crows
}  // namespace einconv
