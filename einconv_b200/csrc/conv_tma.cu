// TMA-fed tcgen05 implicit GEMM for the convNd forward pass (expressions/convNd_forward.py:48-55)
// and, with the roles of the channels swapped and the taps flipped, its input VJP
// (expressions/convNd_input_vjp.py:53-60), for stride 1 and any N, dilation, padding, groups.
//
//   dst[b, g*Cd + n, o..] = sum_{c, k..} src[b, g*Cs + c, o.. + dil*k.. - pad..] * W[g][k..][n][c]
//
// The index pattern turns into a ONE-dimensional row shift.  A pre-pass writes `src` zero-padded and
// channels-last, xp[b][p_0]..[p_{N-1}][c] with p_d = i_d + pad_d, i.e. a 2-d matrix [rows, channels].
// Enumerate output positions by the flattened *padded* coordinate q = sum_d o_d * pitch_d
// (pitch_d = row pitch of the padded grid): the input row of tap (k_0..k_{N-1}) is simply
// q + shift(k), shift(k) = sum_d dil_d * k_d * pitch_d.  So the A operand of a tile of 128 consecutive
// q for one tap is 128 consecutive rows of xp (K-major: channels are contiguous), fetched by 2-d TMA
// boxes with no alignment constraint on the row coordinate, and all the taps of a group (e.g. the
// whole 3x3 window) read overlapping row ranges of ONE staged slab: the slab is loaded once and each
// tap's MMA descriptor just starts `shift` rows further down.  Rows with o_d >= O_d (the wrap-around
// columns of the padded grid) are computed and dropped by the epilogue.
//
// FUSED mode (opt-in, EINCONV_CTMA_FUSED=1; stride 1, innermost source rows that are multiples of 16 bytes):
// no packed copy of the source exists.  Two groups of eight producer warps read the channels-first tensor
// (16-byte loads of 4 fp32 / 8 16-bit positions of one channel), transpose blocks in registers (fp32:
// register renaming, 16-bit: PRMT) and store 16-byte chunks of channels into the 128-byte-swizzled K-major
// slab exactly where TMA would have put them; padding rows and lines outside the tensor are zeros.
// Measured on C1 (profiles/r02_c1_fused_vs_packed.txt): kernel 39.5 us against 16.8 (pack) + 27.4 (TMA-fed
// kernel) under ncu, 37.3 against 37.5 us per step in a graph -- no gain, and 2 x slower on wide layers
// (256 -> 256 channels).  With the loads AND the stores removed the kernel still takes 28 us: the C1 kernel is
// bound by its MMA issue rate (36 N = 64 MMAs of a stage take ~2100 cycles) and the 850 + 300 cycle bubbles of
// a two-stage ring, not by how the slab is fed; the gather itself reaches ~3.5 TB/s chip-wide (64-byte
// fragments of 8 channel planes per warp instruction) where TMA boxes reach twice that.
//
// Roles (192 threads, persistent CTAs, one per SM): warp 0 lane 0 issues TMA (A slabs, per-tap weight
// tiles); warp 1 lane 0 issues tcgen05.mma (kind::tf32 / kind::f16, M = 128, N = BN <= 256, fp32
// accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1);
// warps 2-5 drain TMEM, add the bias, cast and store channels-first (coalesced along positions).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "dispatch.cuh"
#include "tc_ptx.cuh"

namespace einconv {
namespace ctma {

using namespace ptx;

constexpr int kThreads = 192;
constexpr int kProducerWarps = 16;  // FUSED mode: warps 6..21 fill the A slabs from the channels-first source
constexpr int kThreadsFused = kThreads + 32 * kProducerWarps;
constexpr int kTileM = 128;
constexpr int kMaxGroups = 128;
constexpr int kMaxTpg = 64;
constexpr int kSmemBudget = 220 * 1024;

struct Params {
  // tiles: index = ((g * n_ntiles) + ntile) * n_mtiles + mtile
  int n_mtiles, n_ntiles, groups;
  long long total_tiles;
  // K loop: n_slices channel slices x n_groups tap groups; one pipeline stage per (slice, group)
  int n_slices, cb_bytes, cb_elems;
  int n_groups, tpg;
  int slab_rows, box_rows, slab_bytes;
  int bn, btile_bytes;
  int stage_bytes, n_stages;
  int resident;     // every weight tile of one (group, n-tile) class stays in shared memory
  int w_bytes;      // bytes of the resident weight region (0 when streaming)
  int cs;           // source channels per group (elements): TMA column of group g starts at g * cs
  int np;           // padded N rows per (g, tap) in the packed weights
  int taps;
  int group_shift[kMaxGroups];  // row shift of the first tap of each group
  uint32_t tap_off16[kMaxTpg];  // (row offset of tap j of a group inside the slab) * cb_bytes / 16
  // epilogue
  int n_dims;
  long long rows_total;  // B * Ptot
  int batch;
  FastDiv ptot_div;
  FastDiv pitch_div[kMaxDims];
  int out[kMaxDims];
  int cd, cd_total;      // dst channels per group / in total
  long long out_total;
  // destination element of output coordinates (o_0, .., o_{N-1}) of channel c of sample b:
  //   dst[(b * cd_total + c) * chan_stride + dst_off + sum_d o_d * dst_mul[d]]
  // (dense output: dst_mul = row-major strides of `out`, dst_off = 0, chan_stride = out_total; one phase of a
  // strided input VJP: the sub-grid i = st * o + phi of the full input)
  long long dst_mul[kMaxDims], dst_off, chan_stride;
  // tap subset (strided input VJP, tpg = 1): group g is tap group_tap[g] of the packed weights
  int sub;
  int group_tap[kMaxGroups];
  const void* bias;
  void* dst;
  // fp32 tensors on scaled fp16 operand copies: [0] max|src|, [1] max|w| (bit patterns); the epilogue
  // multiplies the accumulator by 1 / (s_src * s_w).  nullptr otherwise.
  const unsigned* absmax;
  long long* trace;  // bring-up only (EINCONV_CTMA_TRACE): clock64() stamps of CTA 0, 16 per tile
  // FUSED mode: the channels-first source on the padded grid
  struct Fill {
    const void* src;
    long long sample_stride, chan_stride;  // elements
    long long sstride[kMaxDims];           // element stride of spatial dim d
    int P[kMaxDims], S[kMaxDims], off[kMaxDims];
    FastDiv P_div[kMaxDims], gw_div;
    int groups_w;                          // S[N-1] / (positions per 16 bytes)
    int ctot;
    int debug;  // bring-up: 1 = no global loads (zeros), 2 = no shared stores, 4 = no row decode
  } fill;
};

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr bool TF32 = true; static constexpr int FMT = 2; };
template <> struct Elem<__nv_bfloat16> { static constexpr bool TF32 = false; static constexpr int FMT = 1; };
template <> struct Elem<__half> { static constexpr bool TF32 = false; static constexpr int FMT = 0; };

constexpr int kMaxStages = 6;
constexpr int kMaxWBars = 32;  // resident weights arrive stage by stage: the first MMAs need not wait for all

struct Bars {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t w_full[kMaxWBars], w_empty;
  uint32_t tmem_base;
  float bias[256];  // bias of the current (group, n-tile) class, zero where absent
};

// All MMAs of one pipeline stage: tpg taps x KS k-steps, issued by the elected thread.  Descriptors
// differ only in the 14-bit start-address field (units of 16 bytes), which never carries out of its
// field for shared-memory addresses: constant part + offsets, a few uniform adds per MMA.
template <bool TF32, int KS>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t btile16,
                                            const uint32_t* tap_off16, int tpg, uint32_t idesc,
                                            uint32_t accumulate) {
  for (int j = 0; j < tpg; ++j) {
    // The tap's 128 rows start tap_row rows into the slab.  The swizzle is a function of the
    // absolute shared-memory address (measured: base-offset field 0 is right, setting it to
    // row % 8 is wrong), so a descriptor may start at any row of a slab TMA wrote with that swizzle.
    const uint64_t a_desc = a_base + (uint64_t)tap_off16[j];
    const uint64_t b_desc = b_base + (uint64_t)((uint32_t)j * btile16);
#pragma unroll
    for (int k = 0; k < KS; ++k) umma<TF32>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accumulate | j | k);
  }
}

#define CTMA_TRACE(slot)                                                                     \
  do {                                                                                       \
    if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[(tile - tile_begin) * 16 + (slot)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// FUSED producers (see the header): same scheme as wgrad_slab.cu's, for a K-major slab whose rows hold
// 128 bytes of channels (32 fp32 / 64 16-bit).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (long long n = 0; !mbar_try_wait(bar, parity); ++n) {
    __nanosleep(32);
    if (n > 200000000ll) __trap();
  }
}
__device__ __forceinline__ uint32_t word_of(const uint4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// Fill rows [0, n_rows) of the slab at `slab_addr` with padded rows [R0, R0 + n_rows) of channels
// [chan0, chan0 + nch) (zero beyond nch) of the source.  ES = element size; a 16-byte load holds EPV = 16 / ES
// positions of one channel, a 16-byte slab chunk EPV channels of one position.  A task moves 128 bytes:
// 8 channels x EPV positions (fp32: two chunks per row, 16-bit: one), so that one round of a 256-thread
// producer group covers a 32 KB slab.
template <int ES>
__device__ __forceinline__ void fill_slab(uint32_t slab_addr, int R0, int n_rows, const Params& p, int cb_bytes,
                                          int chan0, int nch, int tid, int nthreads) {
  constexpr int EPV = 16 / ES;
  constexpr int CPT = 8 / EPV;  // chunks per task: 2 (fp32), 1 (16-bit)
  const Params::Fill& f = p.fill;
  const int N = p.n_dims;
  const int Pw = f.P[N - 1], Sw = f.S[N - 1], offw = f.off[N - 1], gw = f.groups_w;
  const int cols = max(1, (cb_bytes >> 4) / CPT);  // task columns per slab row: 4 / 8 for 128-byte rows
  const int cshift = cols >= 8 ? 3 : (cols == 4 ? 2 : (cols == 2 ? 1 : 0));
  // data groups line * gw + g touched by the rows, lines touched
  const int ra = R0, rb = R0 + n_rows;
  const int La = (int)f.P_div[N - 1].quot((uint32_t)ra), Lb = (int)f.P_div[N - 1].quot((uint32_t)(rb - 1));
  const int wa = ra - La * Pw, wb = rb - 1 - Lb * Pw;
  const int first = wa < offw ? La * gw : (wa < offw + Sw ? La * gw + (wa - offw) / EPV : (La + 1) * gw);
  const int last = wb < offw ? Lb * gw - 1 : (wb < offw + Sw ? Lb * gw + (wb - offw) / EPV : Lb * gw + gw - 1);
  const int n_groups = max(0, last - first + 1), n_lines = Lb - La + 1;
  const int Tdata = n_groups << cshift, Ttot = Tdata + (n_lines << cshift);
  const uint32_t swz_mask = 0x70u & (uint32_t)(cb_bytes - 16);  // address bits [4,7) ^= bits [7,10) (128 / 64 / 32 B)
  for (int task = tid; task < Ttot; task += nthreads) {
    const int c = task & (cols - 1);
    const int t2 = task >> cshift;
    if (task < Tdata) {
      const int slot = first + t2;
      const int L = (int)f.gw_div.quot((uint32_t)slot);
      const int grp = slot - L * gw;
      bool valid = true;
      long long soff = 0;
      uint32_t rem = (uint32_t)L;
      for (int d = N - 2; d >= 0; --d) {
        const uint32_t q = f.P_div[d].quot(rem);
        const int sd = (int)(rem - q * (uint32_t)f.P[d]) - f.off[d];
        rem = q;
        valid = valid && (unsigned)sd < (unsigned)f.S[d];
        soff += (long long)sd * f.sstride[d];
      }
      valid = valid && rem < (uint32_t)p.batch;
      int nvalid = valid ? min(8, nch - c * 8) : 0;  // channels of this task that exist
      if (f.debug & 1) nvalid = 0;
      uint4 in[8];
      const unsigned char* q = reinterpret_cast<const unsigned char*>(f.src) +
                               ((long long)rem * f.sample_stride + (long long)(chan0 + c * 8) * f.chan_stride + soff +
                                (long long)grp * EPV) * ES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        in[j] = make_uint4(0u, 0u, 0u, 0u);
        if (j < nvalid) in[j] = ldg128_stream(q + (long long)j * f.chan_stride * ES);
      }
      const int r0 = L * Pw + offw + grp * EPV - R0;
      // fp32: a quarter warp holds 4 task columns x 2 position groups whose rows are 4 apart; the odd group
      // stores its second chunk first, so the eight 16-byte stores of a quarter hit eight different bank groups
      const int flip = (ES == 4) ? (t2 & 1) : 0;
#pragma unroll
      for (int i = 0; i < EPV; ++i) {
        const int rr = r0 + i;
        if ((unsigned)rr < (unsigned)n_rows && !(f.debug & 2)) {
          const uint32_t row_addr = slab_addr + (uint32_t)rr * (uint32_t)cb_bytes;
          if constexpr (ES == 4) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int h = hh ^ flip;
              const uint32_t a = row_addr + ((uint32_t)(c * 2 + h) << 4);
              const uint32_t o0 = h ? word_of(in[4], i) : word_of(in[0], i), o1 = h ? word_of(in[5], i) : word_of(in[1], i);
              const uint32_t o2 = h ? word_of(in[6], i) : word_of(in[2], i), o3 = h ? word_of(in[7], i) : word_of(in[3], i);
              sts128(a ^ ((a >> 3) & swz_mask), o0, o1, o2, o3);
            }
          } else {
            const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
            const uint32_t a = row_addr + ((uint32_t)c << 4);
            sts128(a ^ ((a >> 3) & swz_mask), __byte_perm(word_of(in[0], i >> 1), word_of(in[1], i >> 1), sel),
                   __byte_perm(word_of(in[2], i >> 1), word_of(in[3], i >> 1), sel),
                   __byte_perm(word_of(in[4], i >> 1), word_of(in[5], i >> 1), sel),
                   __byte_perm(word_of(in[6], i >> 1), word_of(in[7], i >> 1), sel));
          }
        }
      }
    } else {
      // padding rows of line L: wp in [0, offw) and [offw + Sw, Pw)
      const int L = La + (t2 - n_groups);
      const int row_base = L * Pw - R0;
      for (int wp = 0; wp < Pw; ++wp) {
        if (wp == offw) wp += Sw;
        if (wp >= Pw) break;
        const int rr = row_base + wp;
        if ((unsigned)rr < (unsigned)n_rows) {
#pragma unroll
          for (int h = 0; h < CPT; ++h) {
            const uint32_t a = slab_addr + (uint32_t)rr * (uint32_t)cb_bytes + ((uint32_t)(c * CPT + h) << 4);
            sts128(a ^ ((a >> 3) & swz_mask), 0u, 0u, 0u, 0u);
          }
        }
      }
    }
  }
}

// T: operand type of the MMAs; TO: type of the destination and the bias (differs only for fp32 tensors on
// scaled fp16 operands)
template <typename T, typename TO = T, bool FUSED = false>
__global__ void __launch_bounds__(FUSED ? kThreadsFused : kThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // [resident weights | stages: slab (+ the group's weight tiles when streaming) | barriers]
  unsigned char* w_region = smem;
  unsigned char* ring = smem + p.w_bytes;
  Bars* bars = reinterpret_cast<Bars*>(ring + (size_t)p.n_stages * p.stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // FUSED: the producers' arrival completes a stage, plus (streamed weights) the TMA warp's expect_tx arrival
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&bars->full[i], FUSED && !p.resident ? 2 : 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], 4); }
    for (int i = 0; i < kMaxWBars; ++i) mbar_init(&bars->w_full[i], 1);
    mbar_init(&bars->w_empty, 1);
    fence_barrier_init();
    if constexpr (!FUSED) prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
  }
  // TMEM: two accumulators of BN fp32 columns (power of two >= 32)
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.bn) tmem_cols <<= 1;
  if (warp == 1) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // contiguous chunk of tiles per CTA (m-tile fastest): neighbouring tiles share halo rows in L2 and,
  // in resident mode, the weights of their (group, n-tile) class
  const long long tile_begin = p.total_tiles * blockIdx.x / gridDim.x;
  const long long tile_end = p.total_tiles * (blockIdx.x + 1) / gridDim.x;
  const int stages_per_tile = p.n_slices * p.n_groups;
  const int n_wbars = stages_per_tile <= kMaxWBars ? stages_per_tile : 1;

  if (FUSED && warp >= 6) {
    // ------------------------------------------------------------------ software producers (FUSED): A slabs
    // Two groups of eight warps fill alternate stages, so that two rounds of loads (one per stage buffer) are
    // in flight at once: a round is a full L2 / DRAM round trip, and the registers of the producers are the
    // only staging memory.
    constexpr int kGroupThreads = kProducerWarps * 32 / 2;
    const int ptid = threadIdx.x - 6 * 32;
    const int group = ptid / kGroupThreads, gtid = ptid - group * kGroupThreads;
    long long sidx = 0;  // running stage number of this CTA
    for (long long tile = tile_begin; tile < tile_end; ++tile) {
      const int mtile = (int)(tile % p.n_mtiles);
      const int g = (int)((tile / p.n_mtiles) / p.n_ntiles);
      const int q0 = mtile * kTileM;
      for (int s = 0; s < p.n_slices; ++s) {
        for (int grp = 0; grp < p.n_groups; ++grp, ++sidx) {
          if ((int)(sidx & 1) != group) continue;
          const int st = (int)(sidx % p.n_stages);
          const uint32_t ph = (uint32_t)((sidx / p.n_stages) & 1);
          mbar_wait_backoff(&bars->empty[st], ph ^ 1);
          fill_slab<(int)sizeof(T)>(smem_u32(ring + (size_t)st * p.stage_bytes), q0 + p.group_shift[grp], p.slab_rows, p,
                                    p.cb_bytes, g * p.cs + s * p.cb_elems, min(p.cb_elems, p.cs - s * p.cb_elems), gtid,
                                    kGroupThreads);
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          if (group == 0) asm volatile("bar.sync 2, %0;" ::"n"(kGroupThreads) : "memory");
          else asm volatile("bar.sync 3, %0;" ::"n"(kGroupThreads) : "memory");
          if (gtid == 0) mbar_arrive(&bars->full[st]);
        }
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (the whole warp runs the loops and waits; one elected lane issues)
    int st = 0;
    uint32_t ph = 0, pw = 0;
    long long cur_class = -1;
    const uint32_t stage_tx = (FUSED ? 0u : (uint32_t)(p.slab_rows * p.cb_bytes)) +
                              (p.resident ? 0u : (uint32_t)(p.tpg * p.bn * p.cb_bytes));
    for (long long tile = tile_begin; tile < tile_end; ++tile) {
      const int mtile = (int)(tile % p.n_mtiles);
      const long long cls = tile / p.n_mtiles;
      const int ntile = (int)(cls % p.n_ntiles);
      const int g = (int)(cls / p.n_ntiles);
      const int q0 = mtile * kTileM;
      const int wrow0 = g * p.taps * p.np + ntile * p.bn;
      bool load_w = false;
      if (p.resident && cls != cur_class) {
        // (re)load every weight tile of this class once the MMA warp is done with the old ones; the
        // tiles of stage sg are requested right before that stage's slab and signal w_full[sg]
        cur_class = cls;
        load_w = true;
        mbar_wait(&bars->w_empty, pw ^ 1);
        pw ^= 1;
      }
      CTMA_TRACE(0);
      if (load_w && n_wbars == 1) {
        // more stages per tile than weight barriers: ONE barrier for the whole class, so every tile must be
        // requested before the first slab (the MMA warp waits for all of them before it frees a ring slot;
        // requesting them stage by stage dead-locks as soon as the ring is full)
        const int per_slice = p.n_groups * p.tpg;  // tiles of one slice (= taps unless a tap subset is in use)
        if (lane == 0) mbar_expect_tx(&bars->w_full[0], (uint32_t)(p.n_slices * per_slice * p.bn * p.cb_bytes));
        __syncwarp();
        for (int idx = lane; idx < p.n_slices * per_slice; idx += 32) {
          const int s = idx / per_slice, gj = idx - s * per_slice;
          const int tap = p.sub ? p.group_tap[gj] : gj;
          tma_load_2d(w_region + (size_t)idx * p.btile_bytes, &map_b, &bars->w_full[0], s * p.cb_elems,
                      wrow0 + tap * p.np);
        }
        load_w = false;
      }
      int sg = 0;
      for (int s = 0; s < p.n_slices; ++s) {
        for (int grp = 0; grp < p.n_groups; ++grp, ++sg) {
          // One cp.async.bulk.tensor per LANE: a single issuing thread sustains only ~one 8 KB box per
          // 550 cycles (28 GB/s per SM, scripts/tma_bw.cu), while several issuing threads scale linearly.
          if (load_w) {
            uint64_t* wb = &bars->w_full[n_wbars > 1 ? sg : 0];
            if (lane == 0) {
              if (n_wbars > 1)
                mbar_expect_tx(wb, (uint32_t)(p.tpg * p.bn * p.cb_bytes));
              else if (sg == 0)
                mbar_expect_tx(wb, (uint32_t)(p.n_slices * p.taps * p.bn * p.cb_bytes));
            }
            __syncwarp();
            for (int j = lane; j < p.tpg; j += 32) {
              const int tap = p.sub ? p.group_tap[grp] : grp * p.tpg + j;
              tma_load_2d(w_region + (size_t)((s * p.n_groups + grp) * p.tpg + j) * p.btile_bytes, &map_b, wb,
                          s * p.cb_elems, wrow0 + tap * p.np);
            }
          }
          if (FUSED && p.resident) continue;  // nothing to load per stage: the producers own the ring
          mbar_wait(&bars->empty[st], ph ^ 1);
          if (sg < 2) CTMA_TRACE(1 + sg);
          {
            if (lane == 0) mbar_expect_tx(&bars->full[st], stage_tx);
            __syncwarp();
            unsigned char* stage = ring + (size_t)st * p.stage_bytes;
            if constexpr (!FUSED) {
              const int row0 = q0 + p.group_shift[grp];
              const int col_a = g * p.cs + s * p.cb_elems;
              const int n_boxes = p.slab_rows / p.box_rows;
              for (int bx = lane; bx < n_boxes; bx += 32)
                tma_load_2d(stage + (size_t)bx * p.box_rows * p.cb_bytes, &map_a, &bars->full[st], col_a,
                            row0 + bx * p.box_rows);
            }
            if (!p.resident) {
              unsigned char* bt = stage + p.slab_bytes;
              for (int j = lane; j < p.tpg; j += 32)
                tma_load_2d(bt + (size_t)j * p.btile_bytes, &map_b, &bars->full[st], s * p.cb_elems,
                            wrow0 + (p.sub ? p.group_tap[grp] : grp * p.tpg + j) * p.np);
            }
          }
          __syncwarp();
          if (++st == p.n_stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // (the whole warp runs the loops and waits; one elected lane issues tcgen05.mma / commit)
    const uint32_t idesc = make_idesc(Elem<T>::FMT, kTileM, p.bn);
    const uint64_t desc_const = make_desc_kmajor(0, p.cb_bytes);
    const uint32_t ring16 = smem_u32(ring) >> 4, w16 = smem_u32(w_region) >> 4;
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, btile16 = (uint32_t)p.btile_bytes >> 4;
    const uint32_t slab16 = (uint32_t)p.slab_bytes >> 4;
    int st = 0;
    uint32_t ph = 0, pw = 0;
    uint32_t it = 0;
    long long cur_class = -1;
    for (long long tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const long long cls = tile / p.n_mtiles;
      const bool new_w = p.resident && cls != cur_class;
      cur_class = cls;
      CTMA_TRACE(4);
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      CTMA_TRACE(5);
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.bn;
      for (int sg = 0; sg < stages_per_tile; ++sg) {
        if (new_w && (n_wbars > 1 || sg == 0)) mbar_wait(&bars->w_full[n_wbars > 1 ? sg : 0], pw);
        mbar_wait(&bars->full[st], ph);
        tc_fence_after();
        if (sg < 2) CTMA_TRACE(6 + 2 * sg);
        if (elect_one()) {
          const uint32_t stage_a16 = ring16 + (uint32_t)st * stage16;
          const uint64_t a_base = desc_const + (uint64_t)stage_a16;
          // resident: tile (slice, tap) of the class lives at (slice * taps + tap); sg = slice * n_groups + grp
          const uint64_t b_base = desc_const + (uint64_t)(p.resident ? w16 + (uint32_t)(sg * p.tpg) * btile16
                                                                      : stage_a16 + slab16);
          const uint32_t accumulate = sg != 0;
          switch (p.cb_bytes) {
            case 128: issue_stage<Elem<T>::TF32, 4>(d_tmem, a_base, b_base, btile16, p.tap_off16, p.tpg, idesc, accumulate); break;
            case 64: issue_stage<Elem<T>::TF32, 2>(d_tmem, a_base, b_base, btile16, p.tap_off16, p.tpg, idesc, accumulate); break;
            default: issue_stage<Elem<T>::TF32, 1>(d_tmem, a_base, b_base, btile16, p.tap_off16, p.tpg, idesc, accumulate); break;
          }
          umma_commit(&bars->empty[st]);
          if (sg + 1 == stages_per_tile) {
            umma_commit(&bars->tmem_full[acc]);
            // last tile of this class in my chunk: the resident weights may be replaced once these retire
            if (p.resident && (tile + 1 == tile_end || (tile + 1) / p.n_mtiles != cls)) umma_commit(&bars->w_empty);
          }
        }
        __syncwarp();
        if (sg < 2) CTMA_TRACE(7 + 2 * sg);
        if (++st == p.n_stages) { st = 0; ph ^= 1; }
      }
      if (new_w) pw ^= 1;
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lanes 32 * quarter .. + 31 are visible to this warp
    const int etid = threadIdx.x - 64;
    TO* dst = reinterpret_cast<TO*>(p.dst);
    const TO* bias = reinterpret_cast<const TO*>(p.bias);
    float unscale = 1.f;
    if (p.absmax) unscale = 1.f / (operand_scale(__ldg(p.absmax)) * operand_scale(__ldg(p.absmax + 1)));
    uint32_t it = 0;
    long long cur_class = -1;
    for (long long tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int mtile = (int)(tile % p.n_mtiles);
      const long long cls = tile / p.n_mtiles;
      const int ntile = (int)(cls % p.n_ntiles);
      const int g = (int)(cls / p.n_ntiles);
      const int n0 = ntile * p.bn;
      if (cls != cur_class) {
        // bias of this class into shared memory (the four epilogue warps only: named barrier 1)
        cur_class = cls;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = etid; c < p.bn; c += 128)
          bars->bias[c] = (bias && n0 + c < p.cd) ? (float)Cvt<TO>::load(bias + (long long)g * p.cd + n0 + c) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      // decode this thread's row
      const long long row = (long long)mtile * kTileM + quarter * 32 + lane;
      bool valid = row < p.rows_total;
      uint32_t b = 0, rem = 0;
      if (valid) p.ptot_div.divmod((uint32_t)row, b, rem);
      long long oflat = p.dst_off;
      for (int d = 0; d < p.n_dims; ++d) {
        uint32_t od, r2;
        p.pitch_div[d].divmod(rem, od, r2);
        rem = r2;
        valid = valid && (int)od < p.out[d];
        oflat += (long long)od * p.dst_mul[d];
      }
      TO* dst_row = dst + ((long long)b * p.cd_total + (long long)g * p.cd + n0) * p.chan_stride + oflat;

      if (warp == 2) CTMA_TRACE(10);
      mbar_wait_relaxed(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 2) CTMA_TRACE(11);
      const uint32_t taddr = tmem_base + acc * (uint32_t)p.bn + ((uint32_t)(quarter * 32) << 16);
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(taddr + c0, r);
        tmem_ld_wait();
        if (c0 + 16 >= p.bn) {
          // all of this warp's accumulator reads are done: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
          if (warp == 2) CTMA_TRACE(12);
        }
        if (valid) {
          TO* d0 = dst_row + (long long)c0 * p.chan_stride;
          if (n0 + c0 + 16 <= p.cd) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              Cvt<TO>::store(d0 + (long long)j * p.chan_stride,
                             fmaf(__uint_as_float(r[j]), unscale, bars->bias[c0 + j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < p.cd)
                Cvt<TO>::store(d0 + (long long)j * p.chan_stride,
                               fmaf(__uint_as_float(r[j]), unscale, bars->bias[c0 + j]));
          }
        }
      }
      if (warp == 2) CTMA_TRACE(13);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// pre-pass 1: src [B, Ctot, S..] -> xp [B, P_0.., P_{N-1}, Cp], zero-padded, channels-last
// ---------------------------------------------------------------------------------------------
struct PackParams {
  int n_dims;
  int ctot, cp;
  int S[kMaxDims], P[kMaxDims], pad[kMaxDims];
  // phase sub-images of a strided convolution (scalar kernel only): padded-grid position r of dimension d
  // reads source index sstep[d] * (r - pad[d]) + soff[d]; sstep = 1, soff = 0 otherwise
  int sstep[kMaxDims], soff[kMaxDims];
  long long s_total;  // prod(S)
  long long outer_rows;  // B * prod(P[0..N-2])
  int rows_per_cta;
  // optional second job fused into the same launch: pack the weights (see pack_weights below)
  const void* w_src;
  void* w_dst;
  int w_groups, w_taps, w_np, w_cgp, w_cog, w_cg, w_transpose;
  long long w_total;
  unsigned w_ctas, x_ctas;
  FastDiv outer_div[kMaxDims];  // / P[d] for d < n_dims - 1
};

// ---------------------------------------------------------------------------------------------
// pre-pass 2: weights -> wp [G][tap][Np][Cgp] (K-major rows of source channels), zero-padded.
//   forward   : wp[g][t][n][c] = w[g*Cog + n, c, t]                   (n: output channel, c: input channel)
//   input VJP : wp[g][t][n][c] = w[g*Cog + c, n, T-1-t]              (n: input channel,  c: output channel)
// Runs as extra CTAs of the source-packing launch (one launch less on a ~50 us critical path).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void pack_weights_block(const PackParams& pp, unsigned block) {
  const T* __restrict__ w = reinterpret_cast<const T*>(pp.w_src);
  T* __restrict__ wp = reinterpret_cast<T*>(pp.w_dst);
  const int taps = pp.w_taps, np = pp.w_np, cgp = pp.w_cgp, cog = pp.w_cog, cg = pp.w_cg;
  for (long long idx = (long long)block * 1024 + threadIdx.x; idx < min(pp.w_total, ((long long)block + 1) * 1024);
       idx += 256) {
    const int c = (int)(idx % cgp);
    long long r = idx / cgp;
    const int n = (int)(r % np);
    r /= np;
    const int t = (int)(r % taps);
    const int g = (int)(r / taps);
    T v = T(0);
    if (!pp.w_transpose) {
      if (n < cog && c < cg) v = w[(((long long)g * cog + n) * cg + c) * taps + t];
    } else {
      if (n < cg && c < cog) v = w[(((long long)g * cog + c) * cg + n) * taps + (pp.w_transpose == 2 ? t : taps - 1 - t)];
    }
    wp[idx] = v;
  }
}

// fp32 weights -> scaled fp16 packed weights (scale from absmax_w), same layout as pack_weights_block
__device__ __forceinline__ void pack_weights_block_convert(const PackParams& pp, unsigned block,
                                                           const unsigned* __restrict__ absmax_w) {
  const float* __restrict__ w = reinterpret_cast<const float*>(pp.w_src);
  __half* __restrict__ wp = reinterpret_cast<__half*>(pp.w_dst);
  const float scale = operand_scale(__ldg(absmax_w));
  const int taps = pp.w_taps, np = pp.w_np, cgp = pp.w_cgp, cog = pp.w_cog, cg = pp.w_cg;
  for (long long idx = (long long)block * 1024 + threadIdx.x; idx < min(pp.w_total, ((long long)block + 1) * 1024);
       idx += 256) {
    const int c = (int)(idx % cgp);
    long long r = idx / cgp;
    const int n = (int)(r % np);
    r /= np;
    const int t = (int)(r % taps);
    const int g = (int)(r / taps);
    float v = 0.f;
    if (!pp.w_transpose) {
      if (n < cog && c < cg) v = w[(((long long)g * cog + n) * cg + c) * taps + t];
    } else {
      if (n < cg && c < cog) v = w[(((long long)g * cog + c) * cg + n) * taps + (pp.w_transpose == 2 ? t : taps - 1 - t)];
    }
    wp[idx] = __float2half_rn(v * scale);
  }
}

// T = uint16_t / uint32_t (bit copy).  A CTA transposes tiles of 64 channels x 128 bytes of positions
// through shared memory: every warp-level global access is one full 128-byte line on both sides and
// 8 (x2 for 16-bit) independent loads per thread are in flight before the first shared store.
template <typename T>
__global__ void __launch_bounds__(256) pack_channels_last_kernel(const T* __restrict__ src, T* __restrict__ xp,
                                                                 const __grid_constant__ PackParams pp) {
  constexpr int EPW = 4 / sizeof(T);  // elements per 32-bit word
  constexpr int TW = 32 * EPW;        // positions per tile (128 bytes of one channel row)
  constexpr int TC = 64;              // channels per tile
  constexpr int NW = TC / EPW;        // 32-bit words per position in the destination (64 channels)
  __shared__ uint32_t tile[TW][NW + 1];  // odd pitch: conflict-free in both directions
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x >= pp.x_ctas) {  // fused weight job
    if (blockIdx.y == 0) pack_weights_block<T>(pp, blockIdx.x - pp.x_ctas);
    return;
  }
  const int W = pp.n_dims - 1;
  const int c0 = blockIdx.y * TC;
  // The first version of this kernel was instruction-issue bound (75% issue-active at 2.9 TB/s):
  // row pointers and predicates are now hoisted out of the tile loop, 16-bit channel pairs are
  // packed in registers, and shared memory is accessed in 32-bit words only.
  for (int rr = 0; rr < pp.rows_per_cta; ++rr) {
    const long long orow = (long long)blockIdx.x * pp.rows_per_cta + rr;
    if (orow >= pp.outer_rows) break;
    uint32_t rem = (uint32_t)orow;
    bool row_real = true;
    long long src_off = 0, mul = pp.S[W];
    for (int d = W - 1; d >= 0; --d) {
      uint32_t q, r;
      pp.outer_div[d].divmod(rem, q, r);
      rem = q;
      const int i = ((int)r - pp.pad[d]) * pp.sstep[d] + pp.soff[d];
      if (i < 0 || i >= pp.S[d]) row_real = false;
      src_off += (long long)i * mul;
      mul *= pp.S[d];
    }
    const long long b = rem;
    // this warp loads channels c0 + 8 * warp + i
    const T* rowp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + 8 * warp + i;
      rowp[i] = (row_real && c < pp.ctot) ? src + (b * pp.ctot + c) * pp.s_total + src_off : nullptr;
    }
    uint32_t* dst_row = reinterpret_cast<uint32_t*>(xp + orow * pp.P[W] * pp.cp + c0);
    const int cp_words = pp.cp / EPW;
    for (int w0 = 0; w0 < pp.P[W]; w0 += TW) {
      T vals[8][EPW];
#pragma unroll
      for (int e = 0; e < EPW; ++e) {
        const int iw = (w0 + lane + 32 * e - pp.pad[W]) * pp.sstep[W] + pp.soff[W];
        const bool ok = iw >= 0 && iw < pp.S[W];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          vals[i][e] = T(0);
          if (ok && rowp[i]) vals[i][e] = __ldg(rowp[i] + iw);
        }
      }
      if constexpr (EPW == 2) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int ip = 0; ip < 4; ++ip)
            tile[lane + 32 * e][4 * warp + ip] = (uint32_t)vals[2 * ip][e] | ((uint32_t)vals[2 * ip + 1][e] << 16);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[lane][8 * warp + i] = (uint32_t)vals[i][0];
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < TW / 8; ++j) {
        const int wl = warp + 8 * j, w = w0 + wl;
        if (w < pp.P[W]) {
          uint32_t* d = dst_row + (long long)w * cp_words;
#pragma unroll
          for (int h = 0; h < NW / 32; ++h) {
            const int idx = lane + 32 * h;
            // cp is a multiple of 16 bytes: the channels of one word exist together or not at all
            if (c0 + idx * EPW < pp.cp) d[idx] = tile[wl][idx];
          }
        }
      }
      __syncthreads();
    }
  }
}

// Vectorised variant for innermost extents that are a multiple of 4.  The scalar kernel above was
// instruction-issue bound (725 warp instructions per 8 KB tile, 73 % issue-active at 3.3 TB/s): here a
// lane loads FOUR source positions of a channel at once (8 bytes of 16-bit data, 16 bytes of fp32),
// builds whole 16-byte chunks of the destination rows in registers, and both the shared-memory and the
// global side move 16 bytes per lane (~130 warp instructions per tile).  The tile is indexed by SOURCE
// position (aligned loads); the padding columns and the all-padding rows are zero-filled directly.
// Shared layout: tile[position][16-byte chunk], chunks rotated by (position >> 2) so that the
// 4-rows-per-lane writes and the row-wise reads are both conflict-free.
// TS / T: source / destination element; they differ only for the fp32 -> scaled fp16 conversion
// (TS = float, T = uint16_t, `absmax_bits` given).
template <typename T, typename TS = T>
__global__ void __launch_bounds__(256) pack_vec4_kernel(const TS* __restrict__ src, T* __restrict__ xp,
                                                        const __grid_constant__ PackParams pp,
                                                        const unsigned* __restrict__ absmax_bits = nullptr) {
  constexpr int EPW = 4 / sizeof(T);
  constexpr bool CONVERT = !std::is_same_v<T, TS>;
  constexpr int TP = 128;            // source positions per tile
  constexpr int TC = 64;             // channels per tile
  constexpr int CPR = TC / (4 * EPW);  // 16-byte chunks per destination row: 8 (16-bit) or 16 (fp32)
  constexpr int RPW = 32 / CPR;      // destination rows per warp instruction
  constexpr int CCH = 4 * EPW;       // channels per chunk
  __shared__ uint4 tile[TP][CPR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x >= pp.x_ctas) {  // fused weight job
    if (blockIdx.y == 0) {
      if constexpr (CONVERT) pack_weights_block_convert(pp, blockIdx.x - pp.x_ctas, absmax_bits + 1);
      else pack_weights_block<T>(pp, blockIdx.x - pp.x_ctas);
    }
    return;
  }
  const int W = pp.n_dims - 1;
  const int c0 = blockIdx.y * TC;
  const long long orow = blockIdx.x;
  uint32_t rem = (uint32_t)orow;
  bool row_real = true;
  long long src_off = 0, mul = pp.S[W];
  for (int d = W - 1; d >= 0; --d) {
    uint32_t q, r;
    pp.outer_div[d].divmod(rem, q, r);
    rem = q;
    const int i = (int)r - pp.pad[d];
    if (i < 0 || i >= pp.S[d]) row_real = false;
    src_off += (long long)i * mul;
    mul *= pp.S[d];
  }
  const long long b = rem;
  const int PW = pp.P[W], SW = pp.S[W], pad = pp.pad[W];
  const int q = lane & (CPR - 1), rsub = lane / CPR;
  const bool chunk_ok = c0 + q * CCH < pp.cp;  // cp is a multiple of 16 bytes
  const long long row_chunks = (long long)pp.cp / CCH;  // 16-byte chunks per destination row
  uint4* dst = reinterpret_cast<uint4*>(xp + orow * PW * pp.cp + c0) + q;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (!row_real) {  // a line inside the padding of an outer dimension
    if (chunk_ok)
      for (int w = warp * RPW + rsub; w < PW; w += 8 * RPW) dst[w * row_chunks] = zero;
    return;
  }
  if (chunk_ok) {  // the two margins of the line
    for (int w = warp * RPW + rsub; w < pad; w += 8 * RPW) dst[w * row_chunks] = zero;
    for (int w = pad + SW + warp * RPW + rsub; w < PW; w += 8 * RPW) dst[w * row_chunks] = zero;
  }
  const TS* row0 = src + (b * pp.ctot + c0 + 8 * warp) * pp.s_total + src_off + 4 * lane;
  float scale = 1.f;
  if constexpr (CONVERT) scale = operand_scale(__ldg(absmax_bits));
  const int n_ch = min(8, pp.ctot - (c0 + 8 * warp));  // channels of this warp that exist (may be <= 0)
  for (int s0 = 0; s0 < SW; s0 += TP) {
    const bool ok = s0 + 4 * lane < SW;  // SW % 4 == 0: four positions exist together
    if constexpr (CONVERT) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && i < n_ch) v[i] = __ldg(reinterpret_cast<const float4*>(row0 + i * pp.s_total + s0));
      }
      auto h2 = [&](float a, float c) -> uint32_t {
        const __half2 h = __floats2half2_rn(a * scale, c * scale);
        return *reinterpret_cast<const uint32_t*>(&h);
      };
      uint4* t = &tile[4 * lane][(warp + lane) & (CPR - 1)];
      t[0 * CPR] = make_uint4(h2(v[0].x, v[1].x), h2(v[2].x, v[3].x), h2(v[4].x, v[5].x), h2(v[6].x, v[7].x));
      t[1 * CPR] = make_uint4(h2(v[0].y, v[1].y), h2(v[2].y, v[3].y), h2(v[4].y, v[5].y), h2(v[6].y, v[7].y));
      t[2 * CPR] = make_uint4(h2(v[0].z, v[1].z), h2(v[2].z, v[3].z), h2(v[4].z, v[5].z), h2(v[6].z, v[7].z));
      t[3 * CPR] = make_uint4(h2(v[0].w, v[1].w), h2(v[2].w, v[3].w), h2(v[4].w, v[5].w), h2(v[6].w, v[7].w));
    } else if constexpr (EPW == 2) {
      uint2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = make_uint2(0u, 0u);
        if (ok && i < n_ch) v[i] = __ldg(reinterpret_cast<const uint2*>(row0 + i * pp.s_total + s0));
      }
      uint4* t = &tile[4 * lane][(warp + lane) & (CPR - 1)];
      t[0 * CPR] = make_uint4(__byte_perm(v[0].x, v[1].x, 0x5410), __byte_perm(v[2].x, v[3].x, 0x5410),
                              __byte_perm(v[4].x, v[5].x, 0x5410), __byte_perm(v[6].x, v[7].x, 0x5410));
      t[1 * CPR] = make_uint4(__byte_perm(v[0].x, v[1].x, 0x7632), __byte_perm(v[2].x, v[3].x, 0x7632),
                              __byte_perm(v[4].x, v[5].x, 0x7632), __byte_perm(v[6].x, v[7].x, 0x7632));
      t[2 * CPR] = make_uint4(__byte_perm(v[0].y, v[1].y, 0x5410), __byte_perm(v[2].y, v[3].y, 0x5410),
                              __byte_perm(v[4].y, v[5].y, 0x5410), __byte_perm(v[6].y, v[7].y, 0x5410));
      t[3 * CPR] = make_uint4(__byte_perm(v[0].y, v[1].y, 0x7632), __byte_perm(v[2].y, v[3].y, 0x7632),
                              __byte_perm(v[4].y, v[5].y, 0x7632), __byte_perm(v[6].y, v[7].y, 0x7632));
    } else {
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = make_uint4(0u, 0u, 0u, 0u);
        if (ok && i < n_ch) v[i] = __ldg(reinterpret_cast<const uint4*>(row0 + i * pp.s_total + s0));
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4* t = &tile[4 * lane][(2 * warp + h + lane) & (CPR - 1)];
        t[0 * CPR] = make_uint4(v[4 * h].x, v[4 * h + 1].x, v[4 * h + 2].x, v[4 * h + 3].x);
        t[1 * CPR] = make_uint4(v[4 * h].y, v[4 * h + 1].y, v[4 * h + 2].y, v[4 * h + 3].y);
        t[2 * CPR] = make_uint4(v[4 * h].z, v[4 * h + 1].z, v[4 * h + 2].z, v[4 * h + 3].z);
        t[3 * CPR] = make_uint4(v[4 * h].w, v[4 * h + 1].w, v[4 * h + 2].w, v[4 * h + 3].w);
      }
    }
    __syncthreads();
    const int n_rows = min(TP, SW - s0);
    uint4* d = dst + (long long)(s0 + pad) * row_chunks;
#pragma unroll
    for (int it = 0; it < TP / (8 * RPW); ++it) {
      const int row = it * 8 * RPW + warp * RPW + rsub;
      if (row < n_rows && chunk_ok) d[row * row_chunks] = tile[row][(q + (row >> 2)) & (CPR - 1)];
    }
    __syncthreads();
  }
}

// fp32 source -> scaled fp16 destination, same padded channels-last layout (tile: 64 channels x 32 positions)
__global__ void __launch_bounds__(256) pack_convert_kernel(const float* __restrict__ src, __half* __restrict__ xp,
                                                           const __grid_constant__ PackParams pp,
                                                           const unsigned* __restrict__ absmax_bits) {
  constexpr int TW = 32, TC = 64, NW = TC / 2;
  __shared__ uint32_t tile[TW][NW + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x >= pp.x_ctas) {  // fused weight job
    if (blockIdx.y == 0) pack_weights_block_convert(pp, blockIdx.x - pp.x_ctas, absmax_bits + 1);
    return;
  }
  const float scale = operand_scale(__ldg(absmax_bits));
  const int W = pp.n_dims - 1;
  const int c0 = blockIdx.y * TC;
  const long long orow = blockIdx.x;
  uint32_t rem = (uint32_t)orow;
  bool row_real = true;
  long long src_off = 0, mul = pp.S[W];
  for (int d = W - 1; d >= 0; --d) {
    uint32_t q, r;
    pp.outer_div[d].divmod(rem, q, r);
    rem = q;
    const int i = (int)r - pp.pad[d];
    if (i < 0 || i >= pp.S[d]) row_real = false;
    src_off += (long long)i * mul;
    mul *= pp.S[d];
  }
  const long long b = rem;
  const float* rowp[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + 8 * warp + i;
    rowp[i] = (row_real && c < pp.ctot) ? src + (b * pp.ctot + c) * pp.s_total + src_off : nullptr;
  }
  uint32_t* dst_row = reinterpret_cast<uint32_t*>(xp + orow * pp.P[W] * pp.cp + c0);
  const int cp_words = pp.cp / 2;
  for (int w0 = 0; w0 < pp.P[W]; w0 += TW) {
    float vals[8];
    const int iw = w0 + lane - pp.pad[W];
    const bool ok = iw >= 0 && iw < pp.S[W];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      vals[i] = 0.f;
      if (ok && rowp[i]) vals[i] = __ldg(rowp[i] + iw);
    }
#pragma unroll
    for (int ip = 0; ip < 4; ++ip) {
      const __half2 h = __floats2half2_rn(vals[2 * ip] * scale, vals[2 * ip + 1] * scale);
      tile[lane][4 * warp + ip] = *reinterpret_cast<const uint32_t*>(&h);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TW / 8; ++j) {
      const int wl = warp + 8 * j, w = w0 + wl;
      if (w < pp.P[W] && c0 + lane * 2 < pp.cp) dst_row[(long long)w * cp_words + lane] = tile[wl][lane];
    }
    __syncthreads();
  }
}

int pack_channels_last_f32_to_f16(const float* src, void* xp, int n_dims, int batch, int ctot, int cp, const int* S,
                                  const long long* P, const int* pad, const unsigned* absmax_bits,
                                  cudaStream_t stream, const WeightJob* wj) {
  const int N = n_dims;
  PackParams pp{};
  pp.n_dims = N;
  pp.ctot = ctot;
  pp.cp = cp;
  pp.s_total = 1;
  long long outer = batch;
  for (int d = 0; d < N; ++d) {
    pp.S[d] = S[d];
    pp.P[d] = (int)P[d];
    pp.pad[d] = pad[d];
    pp.sstep[d] = 1;
    pp.soff[d] = 0;
    pp.s_total *= S[d];
    if (d < N - 1) {
      outer *= P[d];
      pp.outer_div[d] = FastDiv((uint32_t)P[d]);
    }
  }
  EINCONV_CHECK_ARG(outer < (1ll << 31), "pack_channels_last: too many rows");
  EINCONV_CHECK_ARG(cp % 8 == 0, "pack_channels_last: channel pitch must be a multiple of 16 bytes");
  if ((outer == 0 || cp == 0) && !(wj && wj->total > 0)) return EINCONV_OK;
  dim3 grid((unsigned)outer, (unsigned)std::max(1, (cp + 63) / 64));
  EINCONV_CHECK_ARG(grid.y < 65536, "pack_channels_last: too many channels");
  pp.x_ctas = grid.x;
  if (wj && wj->total > 0) {
    pp.w_src = wj->src; pp.w_dst = wj->dst;
    pp.w_groups = wj->groups; pp.w_taps = wj->taps; pp.w_np = wj->np; pp.w_cgp = wj->cgp;
    pp.w_cog = wj->cog; pp.w_cg = wj->cg; pp.w_transpose = wj->transpose;
    pp.w_total = wj->total;
    pp.w_ctas = (unsigned)((wj->total + 1023) / 1024);
    EINCONV_CHECK_ARG((long long)grid.x + pp.w_ctas < (1ll << 31), "pack_channels_last: grid too large");
    grid.x += pp.w_ctas;
  }
  if (S[N - 1] % 4 == 0 && ((uintptr_t)src & 15) == 0 && !getenv("EINCONV_PACK_SCALAR"))
    pack_vec4_kernel<uint16_t, float><<<grid, 256, 0, stream>>>(src, (uint16_t*)xp, pp, absmax_bits);
  else
    pack_convert_kernel<<<grid, 256, 0, stream>>>(src, (__half*)xp, pp, absmax_bits);
  EINCONV_CHECK_LAUNCH("pack_convert_kernel");
  return EINCONV_OK;
}

// src [B, ctot, S..] -> xp [B, P_0.., P_{N-1}, cp]: zero-padded by pad[d] in front, channels-last
int pack_channels_last(const void* src, void* xp, int dtype, int n_dims, int batch, int ctot, int cp,
                       const int* S, const long long* P, const int* pad, cudaStream_t stream,
                       const WeightJob* wj, const int* sstep, const int* soff) {
  const int N = n_dims;
  PackParams pp{};
  bool strided = false;
  for (int d = 0; d < kMaxDims; ++d) {
    pp.sstep[d] = (sstep && d < N) ? sstep[d] : 1;
    pp.soff[d] = (soff && d < N) ? soff[d] : 0;
    strided = strided || pp.sstep[d] != 1 || pp.soff[d] != 0;
  }
  pp.n_dims = N;
  pp.ctot = ctot;
  pp.cp = cp;
  pp.s_total = 1;
  long long outer = batch;
  for (int d = 0; d < N; ++d) {
    pp.S[d] = S[d];
    pp.P[d] = (int)P[d];
    pp.pad[d] = pad[d];
    pp.s_total *= S[d];
    if (d < N - 1) {
      outer *= P[d];
      pp.outer_div[d] = FastDiv((uint32_t)P[d]);
    }
  }
  EINCONV_CHECK_ARG(outer < (1ll << 31), "pack_channels_last: too many rows");
  if ((outer == 0 || cp == 0) && !(wj && wj->total > 0)) return EINCONV_OK;
  const int es = (int)elem_size(dtype);
  EINCONV_CHECK_ARG(es == 2 || es == 4, "pack_channels_last: unsupported dtype");
  pp.outer_rows = outer;
  pp.rows_per_cta = getenv("EINCONV_PACK_RG") ? atoi(getenv("EINCONV_PACK_RG")) : 1;
  if (pp.rows_per_cta < 1) pp.rows_per_cta = 1;
  dim3 grid((unsigned)((outer + pp.rows_per_cta - 1) / pp.rows_per_cta), (unsigned)((cp + 63) / 64));
  pp.x_ctas = grid.x;
  if (wj && wj->total > 0) {
    pp.w_src = wj->src; pp.w_dst = wj->dst;
    pp.w_groups = wj->groups; pp.w_taps = wj->taps; pp.w_np = wj->np; pp.w_cgp = wj->cgp;
    pp.w_cog = wj->cog; pp.w_cg = wj->cg; pp.w_transpose = wj->transpose;
    pp.w_total = wj->total;
    pp.w_ctas = (unsigned)((wj->total + 1023) / 1024);
    EINCONV_CHECK_ARG((long long)grid.x + pp.w_ctas < (1ll << 31), "pack_channels_last: grid too large");
    grid.x += pp.w_ctas;
  }
  EINCONV_CHECK_ARG(grid.y < 65536, "pack_channels_last: too many channels");
  const bool vec4 = !strided && S[N - 1] % 4 == 0 && ((uintptr_t)src & 15) == 0 && pp.rows_per_cta == 1 &&
                    !getenv("EINCONV_PACK_SCALAR");
  switch (es) {
    case 4:
      if (vec4) pack_vec4_kernel<uint32_t><<<grid, 256, 0, stream>>>((const uint32_t*)src, (uint32_t*)xp, pp);
      else pack_channels_last_kernel<uint32_t><<<grid, 256, 0, stream>>>((const uint32_t*)src, (uint32_t*)xp, pp);
      break;
    default:
      if (vec4) pack_vec4_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)src, (uint16_t*)xp, pp);
      else pack_channels_last_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)src, (uint16_t*)xp, pp);
      break;
  }
  EINCONV_CHECK_LAUNCH("pack_channels_last_kernel");
  return EINCONV_OK;
}

// ---------------------------------------------------------------------------------------------
// host
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

int encode_2d(CUtensorMap* map, int dtype, const void* base, long long rows, long long cols,
                     int box_cols, int box_rows, int swizzle_bytes) {
  const int es = (int)elem_size(dtype);
  CUtensorMapDataType dt = dtype == EINCONV_F32    ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                           : dtype == EINCONV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t est[2] = {1, 1};
  CUtensorMapSwizzle swz = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                           : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
  EncodeFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return EINCONV_ERR_CUDA;
  }
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (2-d) failed with CUresult %d", (int)r);
    return EINCONV_ERR_CUDA;
  }
  return EINCONV_OK;
}

// Rank-n tiled map of raw 2/4/8-byte elements, no swizzle, zero fill outside the tensor: dims / box /
// element strides innermost first, byte strides for dims 1..rank-1.
int encode_tiled_raw(CUtensorMap* map, const void* base, int elem_bytes, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box, const unsigned* elem_strides) {
  EncodeFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return EINCONV_ERR_CUDA;
  }
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = elem_strides[i];
    if (i > 0) st[i - 1] = strides_bytes[i - 1];
  }
  const CUtensorMapDataType dt = elem_bytes == 2   ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                 : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                                                   : CU_TENSOR_MAP_DATA_TYPE_UINT64;
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
    return EINCONV_ERR_CUDA;
  }
  return EINCONV_OK;
}

// The convolution in "role" form: src channels Cs -> dst channels Cd, stride 1.
struct Role {
  int n_dims, batch, groups, cs, cd;
  int S[kMaxDims], O[kMaxDims], K[kMaxDims], dil[kMaxDims], pad[kMaxDims];
  int transpose;  // weights packed for the input VJP
  // Strided forward (stride > 1 in some dimension): the source is packed as PHASE SUB-IMAGES
  //   sub_phi[i'] = x[st * i' + phi],   phi in [0, st)^N,
  // each zero-padded channels-last like a stride-1 source and stacked along the rows of ONE matrix.  Tap k
  // reads x[st * o + dil * k - pad] = sub_phi(k)[o + m(k)] with dil * k - pad = st * m + phi: a stride-1 row
  // shift inside its own sub-image.  The GEMM kernel is unchanged: every tap is its own group whose
  // group shift = (row block of its phase) + (m(k) - m_min) . pitch; only phases that some tap uses exist
  // (a 1x1 stride-2 convolution packs one quarter of x and nothing else: the down-sampling rule of
  // simplifications/opt.py:250-300).
  int strided;
  int st[kMaxDims];
};

static bool make_role(int op, const DevGeom& g, Role* r) {
  r->n_dims = g.n_dims;
  r->batch = g.batch;
  r->groups = g.groups;
  r->strided = 0;
  for (int d = 0; d < g.n_dims; ++d) {
    r->st[d] = g.stride[d];
    if (g.stride[d] != 1) r->strided = 1;
    r->K[d] = g.k[d];
    r->dil[d] = g.dil[d];
  }
  if (r->strided && getenv("EINCONV_CTMA_NO_STRIDE")) return false;
  if (r->strided && op == EINCONV_OP_CONV_INPUT_VJP) {
    // Strided input VJP: gx[st i' + phi] = sum_{k : (phi + p - dil k) % st == 0} v[i' + c(phi, k)] w[k],
    // c = (phi + p - dil k) / st.  One launch of the GEMM kernel per output phase phi over the SAME packed
    // cotangent, with the subset of taps that reach this phase and a strided destination (Params::dst_mul).
    r->strided = 2;
    r->cs = g.c_out_g;
    r->cd = g.c_in_g;
    r->transpose = 2;  // transposed weights, taps in their original order
    for (int d = 0; d < g.n_dims; ++d) {
      r->S[d] = g.out[d];   // spatial sizes of v
      r->O[d] = g.in[d];    // spatial sizes of gx
      r->pad[d] = g.pad[d]; // the convolution's own padding (phase arithmetic)
    }
    return true;
  }
  if (op == EINCONV_OP_CONV_FWD) {
    r->cs = g.c_in_g;
    r->cd = g.c_out_g;
    r->transpose = 0;
    for (int d = 0; d < g.n_dims; ++d) {
      r->S[d] = g.in[d];
      r->O[d] = g.out[d];
      r->pad[d] = g.pad[d];
    }
    return true;
  }
  if (op == EINCONV_OP_CONV_INPUT_VJP) {
    r->cs = g.c_out_g;
    r->cd = g.c_in_g;
    r->transpose = 1;
    for (int d = 0; d < g.n_dims; ++d) {
      r->S[d] = g.out[d];
      r->O[d] = g.in[d];
      r->pad[d] = g.dil[d] * (g.k[d] - 1) - g.pad[d];
      if (r->pad[d] < 0) return false;
    }
    return true;
  }
  return false;
}

struct CPlan {
  bool ok = false;
  bool fused = false;  // A slabs filled from the channels-first source by producer warps (no packed copy)
  int es;
  int cb_bytes, cb_elems, n_slices, cgp, cp;
  int bn, n_ntiles, np;
  int level, n_groups, tpg, taps;
  int slab_rows, box_rows, slab_bytes, btile_bytes, stage_bytes, n_stages, resident = 0, w_bytes = 0;
  size_t smem;
  long long P[kMaxDims], pitch[kMaxDims], ptot, rows_total;
  int n_mtiles;
  long long xp_bytes, wp_bytes;
  // strided forward: phases in use, per-dimension smallest row offset, sub-image padding
  int n_phases = 1;
  int m_min[kMaxDims] = {0};
  long long xp_rows = 0;  // rows of the packed source matrix (all phase sub-images)
};

// dil * k - pad = st * m + phi with 0 <= phi < st
static inline void phase_of(int off, int st, int* m, int* phi) {
  int q = off / st, r = off - q * st;
  if (r < 0) { r += st; --q; }
  *m = q; *phi = r;
}

// compact index of the phase tuple of tap t among the phases that occur (row-major over taps); -1 = count only
static int phase_index(const Role& r, int tap, int* n_phases_out) {
  const int N = r.n_dims;
  int seen[kMaxGroups], n_seen = 0, result = -1;
  int taps = 1;
  for (int d = 0; d < N; ++d) taps *= r.K[d];
  for (int t = 0; t < taps; ++t) {
    int rem = t, code = 0, mul = 1;
    for (int d = N - 1; d >= 0; --d) {
      const int kd = rem % r.K[d];
      rem /= r.K[d];
      int m, phi;
      phase_of(r.dil[d] * kd - r.pad[d], r.st[d], &m, &phi);
      code += phi * mul;
      mul *= r.st[d];
    }
    int idx = -1;
    for (int i = 0; i < n_seen; ++i)
      if (seen[i] == code) idx = i;
    if (idx < 0 && n_seen < kMaxGroups) {
      idx = n_seen;
      seen[n_seen++] = code;
    }
    if (t == tap) result = idx;
  }
  if (n_phases_out) *n_phases_out = n_seen;
  return result;
}

static CPlan make_cplan(const Role& r, int dtype, bool fused_ok = false) {
  CPlan pl;
  pl.es = (int)elem_size(dtype);
  if (dtype == EINCONV_F64) return pl;
  const int es = pl.es, N = r.n_dims;
  // opt-in (EINCONV_CTMA_FUSED=1): measured equal to or slower than the packed route, see the header
  pl.fused = fused_ok && !r.strided && ((long long)r.S[N - 1] * es) % 16 == 0 && getenv("EINCONV_CTMA_FUSED") &&
             !getenv("EINCONV_CTMA_NO_FUSED");
  if (r.groups > 1 && ((long long)r.cs * es) % 16 != 0) return pl;  // TMA column start must be 16-byte aligned
  // channel slice: widest swizzle span among those with the least padding
  {
    long long best_cost = -1;
    for (int cb : {128, 64, 32}) {
      const long long bytes = (long long)r.cs * es;
      const long long cost = (bytes + cb - 1) / cb * cb;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        pl.cb_bytes = cb;
      }
    }
    if (const char* e = getenv("EINCONV_CTMA_CB")) {  // tuning experiments
      const int cb = atoi(e);
      if ((cb == 128 || cb == 64 || cb == 32) && cb <= best_cost) pl.cb_bytes = cb;
    }
    pl.cb_elems = pl.cb_bytes / es;
    pl.n_slices = (r.cs + pl.cb_elems - 1) / pl.cb_elems;
    pl.cgp = pl.n_slices * pl.cb_elems;
    const int per16 = 16 / es;
    // row pitch of xp: all groups' channels, padded so that every slice box lies inside the row
    const long long need = (long long)(r.groups - 1) * r.cs + pl.cgp;
    pl.cp = (int)((need + per16 - 1) / per16 * per16);
  }
  pl.n_ntiles = (r.cd + 255) / 256;
  pl.bn = ((r.cd + pl.n_ntiles - 1) / pl.n_ntiles + 15) / 16 * 16;
  pl.np = pl.n_ntiles * pl.bn;
  pl.btile_bytes = (pl.bn * pl.cb_bytes + 1023) / 1024 * 1024;

  pl.ptot = 1;
  int m_span[kMaxDims] = {0};
  for (int d = N - 1; d >= 0; --d) {
    if (r.strided == 2) {
      // c(phi, k) over every phase and tap that reaches it; sub-grid sizes ceil((I - phi) / st) <= ceil(I / st)
      int lo = 1 << 30, hi = -(1 << 30);
      for (int phi = 0; phi < r.st[d]; ++phi)
        for (int k = 0; k < r.K[d]; ++k) {
          const int num = phi + r.pad[d] - r.dil[d] * k;
          if (((num % r.st[d]) + r.st[d]) % r.st[d] != 0) continue;
          const int c = (num - (((num % r.st[d]) + r.st[d]) % r.st[d])) / r.st[d];
          lo = std::min(lo, c);
          hi = std::max(hi, c);
        }
      if (lo > hi) { lo = hi = 0; }
      pl.m_min[d] = std::min(lo, 0);                       // rows of zero padding in front: -m_min
      m_span[d] = std::max(hi, 0) - pl.m_min[d];
      const long long n_sub = (r.O[d] + r.st[d] - 1) / r.st[d];
      pl.P[d] = std::max<long long>((long long)r.S[d] - pl.m_min[d], n_sub + m_span[d]);
    } else if (r.strided) {
      int lo = 1 << 30, hi = -(1 << 30);
      for (int k = 0; k < r.K[d]; ++k) {
        int m, phi;
        phase_of(r.dil[d] * k - r.pad[d], r.st[d], &m, &phi);
        lo = std::min(lo, m);
        hi = std::max(hi, m);
      }
      pl.m_min[d] = lo;
      m_span[d] = hi - lo;
      pl.P[d] = (long long)r.O[d] + m_span[d];  // sub-image index o + m - m_min for every output and tap
    } else {
      pl.P[d] = std::max<long long>(r.S[d] + r.pad[d], (long long)(r.O[d] - 1) + (long long)r.dil[d] * (r.K[d] - 1) + 1);
    }
    pl.pitch[d] = pl.ptot;
    pl.ptot *= pl.P[d];
    if (pl.ptot >= (1ll << 31)) return pl;
  }
  pl.rows_total = pl.ptot * r.batch;
  long long max_shift = 0;
  pl.taps = 1;
  for (int d = 0; d < N; ++d) {
    max_shift += (r.strided ? (long long)m_span[d] : (long long)r.dil[d] * (r.K[d] - 1)) * pl.pitch[d];
    pl.taps *= r.K[d];
  }
  pl.n_phases = 1;
  if (r.strided) {
    if (pl.taps > kMaxGroups) return pl;
    if (r.strided == 1) phase_index(r, -1, &pl.n_phases);
  }
  pl.xp_rows = pl.rows_total * pl.n_phases;
  if (pl.xp_rows + max_shift + 4096 >= (1ll << 31)) return pl;
  pl.n_mtiles = (int)((pl.rows_total + kTileM - 1) / kTileM);
  pl.box_rows = std::min(256, 8192 / pl.cb_bytes);

  // tap grouping: the trailing `level` dims of a group share one slab; one pipeline stage holds the
  // slab of a (slice, group) and, unless the weights are resident, the group's weight tiles
  // (strided: every tap is its own group — taps of different phases read different sub-images)
  int level_max = r.strided ? 0 : N;
  if (const char* e = getenv("EINCONV_CTMA_LEVEL")) level_max = std::min(level_max, std::max(0, atoi(e)));
  for (int level = level_max; level >= 0 && !pl.ok; --level) {
    long long tpg = 1, span = 0;
    for (int d = N - level; d < N; ++d) {
      tpg *= r.K[d];
      span += (long long)r.dil[d] * (r.K[d] - 1) * pl.pitch[d];
    }
    if (tpg > kMaxTpg || pl.taps / tpg > kMaxGroups) continue;
    if (getenv("EINCONV_CTMA_SMALL_BOXES")) {
      // boxes of just enough rows for one box per lane (instead of 8 KB boxes): slabs without the rounding
      const long long need = (kTileM + span + 7) / 8 * 8;
      pl.box_rows = (int)((need + 31) / 32 + 7) / 8 * 8;
    }
    // TMA moves whole boxes; the producer warps of FUSED mode any number of rows (8-row swizzle atoms)
    const long long rows = pl.fused ? (kTileM + span + 7) / 8 * 8 : (kTileM + span + pl.box_rows - 1) / pl.box_rows * pl.box_rows;
    const long long slab = (rows * pl.cb_bytes + 1023) / 1024 * 1024;
    pl.level = level;
    pl.tpg = (int)tpg;
    pl.n_groups = pl.taps / pl.tpg;
    pl.slab_rows = (int)rows;
    pl.slab_bytes = (int)slab;
    // weights resident in shared memory if every tile of one (group, n-tile) class fits next to two slabs
    const long long w_bytes = (long long)pl.taps * pl.n_slices * pl.btile_bytes;
    if (!getenv("EINCONV_CTMA_STREAM") && w_bytes + 2 * slab <= kSmemBudget) {
      pl.resident = 1;
      pl.w_bytes = (int)w_bytes;
      pl.stage_bytes = (int)slab;
      pl.n_stages = (int)std::min<long long>(kMaxStages, (kSmemBudget - w_bytes) / slab);
      pl.ok = true;
      break;
    }
    const long long stage = slab + tpg * pl.btile_bytes;
    if (2 * stage > kSmemBudget) continue;
    pl.resident = 0;
    pl.w_bytes = 0;
    pl.stage_bytes = (int)stage;
    pl.n_stages = (int)std::min<long long>(kMaxStages, kSmemBudget / stage);
    pl.ok = true;
  }
  if (!pl.ok) return pl;
  pl.smem = (size_t)pl.w_bytes + (size_t)pl.n_stages * pl.stage_bytes + sizeof(Bars) + 64;
  pl.xp_bytes = ((pl.xp_rows * pl.cp * es) + 1023) / 1024 * 1024;
  pl.wp_bytes = (((long long)r.groups * pl.taps * pl.np * pl.cgp * es) + 1023) / 1024 * 1024;
  return pl;
}

bool supported(int op, const DevGeom& g, int dtype, int math) {
  if (dtype == EINCONV_F64) return false;
  if (dtype == EINCONV_F32 && math != EINCONV_MATH_TF32) return false;
  if (getenv("EINCONV_NO_CONV_TMA")) return false;
  Role r;
  if (!make_role(op, g, &r)) return false;
  return make_cplan(r, dtype).ok;
}

int64_t workspace_bytes(int op, const DevGeom& g, int dtype) {
  Role r;
  if (!make_role(op, g, &r)) return 0;
  const CPlan pl = make_cplan(r, dtype);
  if (!pl.ok) return 0;
  int64_t need = pl.xp_bytes + pl.wp_bytes + 1024;
  if (dtype == EINCONV_F32) {  // scaled fp16 operand copies: [absmax header][xp][wp]
    const CPlan ph = make_cplan(r, EINCONV_F16);
    if (ph.ok) need = std::max<int64_t>(need, ph.xp_bytes + ph.wp_bytes + 2048);
  }
  return need;
}

// T: operand type of the MMAs (`dtype` is ITS enum); TO: tensor type.  TO = float with T = __half is the
// TF32-mode route on power-of-two-scaled fp16 operand copies (same 10-bit mantissa as TF32, round to
// nearest instead of truncation, twice the MMA rate on half the bytes; see convert.cu).
template <typename T, typename TO = T>
static int run_t(const Role& r, const CPlan& pl, const void* src, const void* w, const void* bias, void* dst,
                 int dtype, void* workspace, cudaStream_t stream) {
  constexpr bool CONVERT = !std::is_same<T, TO>::value;
  // workspace may be arbitrarily aligned: TMA needs 16 bytes for the global address
  uintptr_t base = ((uintptr_t)workspace + 1023) & ~(uintptr_t)1023;
  unsigned* absmax = nullptr;
  if (CONVERT) {
    absmax = reinterpret_cast<unsigned*>(base);  // [0] max|src|, [1] max|w|
    base += 1024;
  }
  T* xp = reinterpret_cast<T*>(base);
  T* wp = reinterpret_cast<T*>(base + pl.xp_bytes);
  const int N = r.n_dims;

  // ---- pack the source and (same launch) the weights
  {
    WeightJob wj{};
    wj.src = w; wj.dst = wp;
    wj.groups = r.groups; wj.taps = pl.taps; wj.np = pl.np; wj.cgp = pl.cgp;
    wj.cog = r.transpose ? r.cs : r.cd;
    wj.cg = r.transpose ? r.cd : r.cs;
    wj.transpose = r.transpose;
    wj.total = (long long)r.groups * pl.taps * pl.np * pl.cgp;
    int st0;
    if (pl.fused) {
      // weights only (the source is read by the kernel's producer warps): a packing launch with no source rows
      st0 = pack_channels_last(nullptr, xp, dtype, N, 0, 64, 64, r.S, pl.P, r.pad, stream, &wj);
    } else if (CONVERT) {
      long long n_src = (long long)r.batch * r.groups * r.cs, n_w = (long long)r.groups * r.cs * r.cd;
      for (int d = 0; d < N; ++d) {
        n_src *= r.S[d];
        n_w *= r.K[d];
      }
      cudaError_t e = cudaMemsetAsync(absmax, 0, 8, stream);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(absmax)");
      st0 = launch_absmax2((const float*)src, n_src, (const float*)w, n_w, absmax, stream);
      if (st0) return st0;
      st0 = pack_channels_last_f32_to_f16((const float*)src, xp, N, r.batch, r.groups * r.cs, pl.cp, r.S, pl.P,
                                          r.pad, absmax, stream, &wj);
    } else if (!r.strided) {
      st0 = pack_channels_last(src, xp, dtype, N, r.batch, r.groups * r.cs, pl.cp, r.S, pl.P, r.pad, stream, &wj);
    } else if (r.strided == 2) {
      int pad_v[kMaxDims];
      for (int d = 0; d < N; ++d) pad_v[d] = -pl.m_min[d];
      st0 = pack_channels_last(src, xp, dtype, N, r.batch, r.groups * r.cs, pl.cp, r.S, pl.P, pad_v, stream, &wj);
    } else {
      // one packing launch per phase in use: sub-image index i' = (padded position) + m_min reads x[st * i' + phi]
      st0 = EINCONV_OK;
      bool done[kMaxGroups] = {false};
      for (int t = 0; t < pl.taps && !st0; ++t) {
        const int ph = phase_index(r, t, nullptr);
        if (ph < 0 || done[ph]) continue;
        done[ph] = true;
        int pad_sub[kMaxDims], soff[kMaxDims];
        int rem = t;
        for (int d = N - 1; d >= 0; --d) {
          const int kd = rem % r.K[d];
          rem /= r.K[d];
          int m, phi;
          phase_of(r.dil[d] * kd - r.pad[d], r.st[d], &m, &phi);
          pad_sub[d] = -pl.m_min[d];
          soff[d] = phi;
        }
        T* dst_phase = xp + (long long)ph * pl.rows_total * pl.cp;
        st0 = pack_channels_last(src, dst_phase, dtype, N, r.batch, r.groups * r.cs, pl.cp, r.S, pl.P, pad_sub, stream,
                                 ph == 0 ? &wj : nullptr, r.st, soff);
      }
    }
    if (st0) return st0;
  }
  // ---- tensor maps
  CUtensorMap map_a, map_b;
  memset(&map_a, 0, sizeof(map_a));
  int st = pl.fused ? 0 : encode_2d(&map_a, dtype, xp, pl.xp_rows, pl.cp, pl.cb_elems, pl.box_rows, pl.cb_bytes);
  if (st) return st;
  st = encode_2d(&map_b, dtype, wp, (long long)r.groups * pl.taps * pl.np, pl.cgp, pl.cb_elems, pl.bn, pl.cb_bytes);
  if (st) return st;

  Params p{};
  p.n_mtiles = pl.n_mtiles;
  p.n_ntiles = pl.n_ntiles;
  p.groups = r.groups;
  p.total_tiles = (long long)pl.n_mtiles * pl.n_ntiles * r.groups;
  p.n_slices = pl.n_slices;
  p.cb_bytes = pl.cb_bytes;
  p.cb_elems = pl.cb_elems;
  p.n_groups = pl.n_groups;
  p.tpg = pl.tpg;
  p.slab_rows = pl.slab_rows;
  p.box_rows = pl.box_rows;
  p.slab_bytes = pl.slab_bytes;
  p.bn = pl.bn;
  p.btile_bytes = pl.btile_bytes;
  p.stage_bytes = pl.stage_bytes;
  p.n_stages = pl.n_stages;
  p.resident = pl.resident;
  p.w_bytes = pl.w_bytes;
  p.cs = r.cs;
  p.np = pl.np;
  p.taps = pl.taps;
  // shifts: tap t = (k_0, .., k_{N-1}) row-major; group = leading N - level coordinates
  for (int t = 0; t < pl.taps; ++t) {
    long long shift = 0;
    int rem = t;
    for (int d = N - 1; d >= 0; --d) {
      const int kd = rem % r.K[d];
      rem /= r.K[d];
      if (r.strided) {
        int m, phi;
        phase_of(r.dil[d] * kd - r.pad[d], r.st[d], &m, &phi);
        shift += (long long)(m - pl.m_min[d]) * pl.pitch[d];
      } else {
        shift += (long long)r.dil[d] * kd * pl.pitch[d];
      }
    }
    if (r.strided == 1) shift += (long long)phase_index(r, t, nullptr) * pl.rows_total;
    if (r.strided == 2) shift = 0;  // set per output phase below
    if (t % pl.tpg == 0) p.group_shift[t / pl.tpg] = (int)shift;
    // group 0: offsets relative to its first tap (shift 0)
    if (t < pl.tpg) p.tap_off16[t] = (uint32_t)((shift * pl.cb_bytes) >> 4);
  }
  p.n_dims = N;
  p.rows_total = pl.rows_total;
  p.batch = r.batch;
  p.ptot_div = FastDiv((uint32_t)pl.ptot);
  p.out_total = 1;
  for (int d = 0; d < N; ++d) {
    p.pitch_div[d] = FastDiv((uint32_t)pl.pitch[d]);
    p.out[d] = r.O[d];
    p.out_total *= r.O[d];
  }
  p.cd = r.cd;
  p.cd_total = r.groups * r.cd;
  p.bias = bias;
  p.dst = dst;
  p.absmax = absmax;
  p.chan_stride = p.out_total;
  p.dst_off = 0;
  {
    long long mul = 1;
    for (int d = N - 1; d >= 0; --d) {
      p.dst_mul[d] = mul;
      mul *= r.O[d];
    }
  }

  if (pl.fused) {
    Params::Fill& f = p.fill;
    f.src = src;
    f.ctot = r.groups * r.cs;
    long long mul = 1;
    for (int d = N - 1; d >= 0; --d) {
      f.P[d] = (int)pl.P[d];
      f.P_div[d] = FastDiv((uint32_t)pl.P[d]);
      f.S[d] = r.S[d];
      f.off[d] = r.pad[d];
      f.sstride[d] = mul;
      mul *= r.S[d];
    }
    f.chan_stride = mul;
    f.sample_stride = mul * f.ctot;
    f.groups_w = r.S[N - 1] / (16 / pl.es);
    f.gw_div = FastDiv((uint32_t)std::max(1, f.groups_w));
    f.debug = getenv("EINCONV_CTMA_FILL_DEBUG") ? atoi(getenv("EINCONV_CTMA_FILL_DEBUG")) : 0;
  }
  const unsigned grid = (unsigned)std::min<long long>(p.total_tiles, num_sms());
  if constexpr (!CONVERT) {
    if (pl.fused) {
      auto kernf = conv_tma_kernel<T, TO, true>;
      cudaError_t ef = cudaFuncSetAttribute(kernf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
      if (ef != cudaSuccess) return cuda_fail(ef, "cudaFuncSetAttribute(conv_tma fused)");
      kernf<<<grid, kThreadsFused, pl.smem, stream>>>(map_a, map_b, p);
      EINCONV_CHECK_LAUNCH("conv_tma_kernel");
      return EINCONV_OK;
    }
  }
  auto kern = conv_tma_kernel<T, TO>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tma)");
  // (programmatic dependent launch was measured here: no gain, the GEMM's CTAs need the whole SM)
  long long* trace = nullptr;
  const int trace_tiles = 16;
#ifdef EINCONV_BRINGUP  // allocates, synchronises and frees: never inside a product build (breaks graph capture)
  if (getenv("EINCONV_CTMA_TRACE")) {
    cudaMalloc(&trace, trace_tiles * 16 * sizeof(long long));
    cudaMemset(trace, 0, trace_tiles * 16 * sizeof(long long));
    p.trace = trace;
  }
#endif
  if (r.strided == 2) {
    // one launch per output phase; phases no tap reaches stay zero
    int n_phase_total = 1;
    for (int d = 0; d < N; ++d) n_phase_total *= r.st[d];
    bool need_zero = false;
    std::vector<Params> launches;
    for (int ph = 0; ph < n_phase_total; ++ph) {
      int phi[kMaxDims], rem_ph = ph;
      for (int d = N - 1; d >= 0; --d) {
        phi[d] = rem_ph % r.st[d];
        rem_ph /= r.st[d];
      }
      Params q = p;
      q.sub = 1;
      q.n_groups = 0;
      bool empty = false;
      q.dst_off = 0;
      for (int d = 0; d < N; ++d) {
        q.out[d] = (r.O[d] - phi[d] + r.st[d] - 1) / r.st[d];
        if (q.out[d] <= 0) empty = true;
        q.dst_mul[d] = p.dst_mul[d] * r.st[d];
        q.dst_off += (long long)phi[d] * p.dst_mul[d];
      }
      if (empty) continue;
      for (int t = 0; t < pl.taps; ++t) {
        long long shift = 0;
        int rem = t;
        bool ok = true;
        for (int d = N - 1; d >= 0; --d) {
          const int kd = rem % r.K[d];
          rem /= r.K[d];
          const int num = phi[d] + r.pad[d] - r.dil[d] * kd;
          const int mod = ((num % r.st[d]) + r.st[d]) % r.st[d];
          if (mod != 0) { ok = false; break; }
          shift += (long long)(num / r.st[d] - pl.m_min[d]) * pl.pitch[d];
        }
        if (!ok) continue;
        q.group_tap[q.n_groups] = t;
        q.group_shift[q.n_groups] = (int)shift;
        ++q.n_groups;
      }
      if (q.n_groups == 0) {
        need_zero = true;
        continue;
      }
      launches.push_back(q);
    }
    if (need_zero) {
      long long n_dst = (long long)r.batch * r.groups * r.cd * p.out_total;
      cudaError_t ez = cudaMemsetAsync(dst, 0, (size_t)n_dst * sizeof(TO), stream);
      if (ez != cudaSuccess) return cuda_fail(ez, "cudaMemsetAsync(gx)");
    }
    for (const Params& q : launches) {
      kern<<<grid, kThreads, pl.smem, stream>>>(map_a, map_b, q);
      EINCONV_CHECK_LAUNCH("conv_tma_kernel");
    }
    return EINCONV_OK;
  }
  kern<<<grid, kThreads, pl.smem, stream>>>(map_a, map_b, p);
  EINCONV_CHECK_LAUNCH("conv_tma_kernel");
  if (trace) {  // bring-up only: synchronises and prints the pipeline timeline of CTA 0
    long long h[trace_tiles * 16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(trace);
    long long t0 = 0;
    for (int i = 0; i < trace_tiles * 16; ++i)
      if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
    fprintf(stderr, "conv_tma trace (cycles since first stamp; P=producer M=mma E=epilogue warp 2)\n");
    fprintf(stderr, "tile  P:start  P:st0   P:st1 | M:start M:tmem  M:full0 M:iss0  M:full1 M:iss1 | E:start E:full  E:ldone E:done\n");
    for (int t = 0; t < trace_tiles; ++t) {
      if (!h[t * 16]) continue;
      fprintf(stderr, "%4d ", t);
      for (int k : {0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13}) fprintf(stderr, "%7lld ", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return EINCONV_OK;
}

int run(int op, const void* src, const void* w, const void* bias, void* dst, int dtype, const DevGeom& g,
        void* workspace, int64_t workspace_bytes_given, cudaStream_t stream) {
  Role r;
  if (!make_role(op, g, &r)) {
    set_error("conv_tma: unsupported geometry");
    return EINCONV_ERR_UNSUPPORTED;
  }
  CPlan pl = make_cplan(r, dtype);
  if (!pl.ok) {
    set_error("conv_tma: unsupported geometry");
    return EINCONV_ERR_UNSUPPORTED;
  }
  if (((uintptr_t)src % 16) == 0) {
    // FUSED producers (same workspace layout; the packed source is simply not written)
    const CPlan pf = make_cplan(r, dtype, true);
    if (pf.ok && pf.fused && pf.xp_bytes == pl.xp_bytes && pf.wp_bytes == pl.wp_bytes) pl = pf;
  }
  if (!workspace || workspace_bytes_given < pl.xp_bytes + pl.wp_bytes + 1024) {
    set_error("conv_tma: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes_given,
              (long long)(pl.xp_bytes + pl.wp_bytes + 1024));
    return EINCONV_ERR_WORKSPACE;
  }
  if ((long long)r.batch * r.groups * r.cd == 0 || pl.rows_total == 0) return EINCONV_OK;
  // fp32 tensors (TF32 mode): scaled fp16 operand copies run the main kernel a third faster (C1: 25.0 -> 16.7 us
  // under ncu) but cost an extra pass over the source for its maximum (13 us for C1's 25.7 MB), so the route
  // pays only when the GEMM outweighs the packing: many output channels x taps per source element.
  // EINCONV_CTMA_HALF=1 / EINCONV_CTMA_TF32=1 force either route.
  long long taps_total = 1;
  for (int d = 0; d < r.n_dims; ++d) taps_total *= r.K[d];
  const bool want_half = getenv("EINCONV_CTMA_HALF") != nullptr || (long long)r.cd * taps_total >= 2048;
  if (dtype == EINCONV_F32 && want_half && !r.strided && !getenv("EINCONV_CTMA_TF32") && ((uintptr_t)src % 16) == 0 &&
      ((uintptr_t)w % 16) == 0) {
    const CPlan ph = make_cplan(r, EINCONV_F16);
    if (ph.ok && workspace_bytes_given >= ph.xp_bytes + ph.wp_bytes + 2048)
      return run_t<__half, float>(r, ph, src, w, bias, dst, EINCONV_F16, workspace, stream);
  }
  switch (dtype) {
    case EINCONV_F32: return run_t<float>(r, pl, src, w, bias, dst, dtype, workspace, stream);
    case EINCONV_BF16: return run_t<__nv_bfloat16>(r, pl, src, w, bias, dst, dtype, workspace, stream);
    case EINCONV_F16: return run_t<__half>(r, pl, src, w, bias, dst, dtype, workspace, stream);
    default: set_error("conv_tma: unsupported dtype"); return EINCONV_ERR_UNSUPPORTED;
  }
}

}  // namespace ctma
}  // namespace einconv
