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
//
// Two producers fill the slabs:
//  * TMA boxes of channels-last padded copies xp / vp written by a packing pre-pass (any geometry), or
//  * FUSED: sixteen producer warps read the channels-FIRST tensors themselves (16-byte loads of 8
//    positions of one channel, 8 x 8 transposes of 16-bit elements in registers with PRMT, 16-byte
//    swizzled shared-memory stores of 8 channels of one position), padding rows written as zeros.  No
//    packed copies exist.  Needs innermost extents that are multiples of 8 (16-byte aligned source rows).
//    Measured SLOWER than the packed route on C4 (790 vs 503 us): the producers' registers are the only
//    staging memory (64 KB in flight per SM), so a stage costs a full DRAM round trip plus the tail of a
//    chip-wide burst, and their shared-memory stores queue behind the tensor core's operand reads.  Kept as
//    an opt-in (EINCONV_WGRAD_FUSED=1) and as the building block of the next mode.
//  * SELF-PACKING (opt-in, EINCONV_WGRAD_SELFPACK=1; every CTA must be resident): the same producer warps
//    write the packed copies to GLOBAL memory, one chunk of rows ahead of the TMA / MMA warps of the same
//    kernel, and publish each chunk through a release counter; the TMA warp acquires the counter of the
//    chunk its next k-block reads.  The packing passes then overlap the tensor-core work instead of
//    preceding it -- but with only two 104 KB stages the MMA pipeline is sensitive to TMA latency, and the
//    packing traffic lengthens it: C4 530 us against 503 us for pack-then-compute (tensor pipe 51 % busy).
// K-blocks are dealt round-robin over the K splits (all CTAs walk the tensors front to back together, so
// the three tap groups and neighbouring splits find each other's slabs in L2: DRAM reads of the main
// kernel 970 -> ~510 MB, C4 515 -> 503 us).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "dispatch.cuh"
#include "tc_ptx.cuh"

namespace einconv {
namespace wslab {

using namespace ptx;

constexpr int kThreads = 192;
constexpr int kProducerWarps = 18;  // FUSED mode: warps 6..23 fill the slabs from the channels-first tensors
constexpr int kThreadsFused = kThreads + 32 * kProducerWarps;
constexpr int kMaxNb = 4;           // N blocks of 64 columns per MMA in multi-tap mode
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
  signed char tap_tbl[kMaxTpg / 2][2][kMaxNb];
  // multi-tap mode (BN = 64 only): the B operand is N = 64 * nb with its 64-column blocks `blk_rows` ROWS
  // apart in a vp slab that starts (nb - 1) * blk_rows before the k-block, so one MMA pairs each of its
  // two xp units with vp[r - (nb - 1 - j) * blk_rows], j = 0..nb-1: 2 * nb taps per M = 128, N = 64 * nb
  // instruction, which is tensor-bound where N = 64 is shared-memory-read bound (3 x 3 windows: two
  // N = 192 MMAs per 16 rows instead of five N = 64 ones).
  int nb, blk_rows;
  int v_early;            // rows the vp slab starts before the k-block ((nb - 1) * blk_rows)
  int v_rows;             // rows of vp staged per 64-channel block
  int acc_cols;           // TMEM columns per MMA accumulator (bn, or 64 * nb in multi-tap mode)
  float* partial;         // [k_splits][groups * cog * cg * taps]
  long long gw_elems;
  // SELF-PACKING mode: packed copies in global memory, written chunk by chunk by the producer warps
  unsigned short* xp;
  unsigned short* vp;
  unsigned* chunk_flags;  // [n_chunks] CTAs that have finished packing their share of the chunk
  int n_chunks, chunk_rows, cxp, cvp, rows_total;
  int round_robin;        // k-blocks dealt round-robin over the K splits instead of contiguous ranges
  // FUSED / SELF-PACKING producers: the channels-first tensors and the padded grid
  int n_dims, batch;
  int P[kMaxDims];        // padded grid extents
  FastDiv P_div[kMaxDims];
  int bias_lines;         // rows are biased by bias_lines * P[N-1] so that the early vp rows stay non-negative
  struct Src {
    const void* ptr;
    long long sample_stride, chan_stride;   // elements
    long long sstride[kMaxDims];            // element stride of spatial dim d
    int S[kMaxDims];                        // source extents
    int off[kMaxDims];                      // padded coordinate of source index 0
    int ctot;                               // channels of the tensor
    int groups_w;                           // S[N-1] / 8
    FastDiv gw_div;
  } sx, sv;
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

// ---------------------------------------------------------------------------------------------
// FUSED producer: rows of the padded channels-last grid of one 64-channel block, read from the
// channels-first tensor and written as a 128-byte-swizzled MN-major slab (row r at r * 128, 16-byte chunk
// c of the row at chunk c ^ (r & 7): what TMA's SWIZZLE_128B writes for the same box).
// A data task = (group of 8 source positions along the innermost dim, 8 channels): eight 16-byte loads,
// an 8 x 8 transpose of 16-bit elements in registers (PRMT), eight 16-byte stores.  Consecutive lanes take
// consecutive channel octets, so a quarter warp stores one whole 128-byte row (conflict-free) and a warp
// load covers 64 contiguous bytes of each of 8 channels.  The data groups of a row range are numbered
// line * groups_w + g, so one stage is ONE round of loads for the 576 producer threads (the registers of
// the producers are the only staging memory, and a round costs a full L2 / DRAM round trip): the rows a
// slab shares with the previous k-block are copied from the previous stage's slab while the loads are
// in flight.  Padding rows are written as zeros by separate (load-free) tasks; lines outside the tensor
// (padding planes, rows past the batch, rows before the first) are all zeros.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void producers_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");
}
// producers are not latency-critical at the 100 ns scale: back off between polls instead of spinning
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (long long n = 0; !mbar_try_wait(bar, parity); ++n) {
    __nanosleep(64);
    if (n > 200000000ll) __trap();
  }
}

// rows [ra, rb) (biased padded rows) of a slab whose row 0 is biased padded row R0
struct SegGeom {
  int R0, ra, rb;
  int s0, n_groups;  // data groups line * groups_w + g touched by the range
  int l0, n_lines;   // lines touched (one padding task per line and channel octet)
};
__device__ __forceinline__ SegGeom seg_geom(const Params& p, const Params::Src& src, int R0, int ra, int rb) {
  SegGeom sg;
  sg.R0 = R0;
  sg.ra = ra;
  sg.rb = rb;
  const int N = p.n_dims;
  const int Pw = p.P[N - 1], Sw = src.S[N - 1], offw = src.off[N - 1], gw = src.groups_w;
  const int La = (int)p.P_div[N - 1].quot((uint32_t)ra), Lb = (int)p.P_div[N - 1].quot((uint32_t)(rb - 1));
  const int wa = ra - La * Pw, wb = rb - 1 - Lb * Pw;
  const int first = wa < offw ? La * gw : (wa < offw + Sw ? La * gw + ((wa - offw) >> 3) : (La + 1) * gw);
  const int last = wb < offw ? Lb * gw - 1 : (wb < offw + Sw ? Lb * gw + ((wb - offw) >> 3) : Lb * gw + gw - 1);
  sg.s0 = first;
  sg.n_groups = rb > ra ? max(0, last - first + 1) : 0;
  sg.l0 = La;
  sg.n_lines = rb > ra ? Lb - La + 1 : 0;
  return sg;
}

struct DataTask {
  uint4 in[8];
  uint32_t slab_addr;
  int r0, c8;      // slab row of the group's first position (relative to the slab), channel octet
  int lo, hi;      // slab rows that may be written: [lo, hi)
};
// issue the eight loads of data group `slot` (no wait)
__device__ __forceinline__ void load_task(DataTask& t, const Params& p, const Params::Src& src, const SegGeom& sg,
                                          uint32_t slab_addr, int chan0, int nch_valid, int slot, int c8) {
  const int N = p.n_dims;
  const int Lb = (int)src.gw_div.quot((uint32_t)slot);
  const int grp = slot - Lb * src.groups_w;
  const int L = Lb - p.bias_lines;
  bool valid = L >= 0 && c8 * 8 < nch_valid;
  long long soff = 0;
  uint32_t rem = L < 0 ? 0u : (uint32_t)L;
  for (int d = N - 2; d >= 0; --d) {
    const uint32_t q = p.P_div[d].quot(rem);
    const int sd = (int)(rem - q * (uint32_t)p.P[d]) - src.off[d];
    rem = q;
    valid = valid && (unsigned)sd < (unsigned)src.S[d];
    soff += (long long)sd * src.sstride[d];
  }
  valid = valid && rem < (uint32_t)p.batch;
  if (valid) {
    const unsigned short* q = reinterpret_cast<const unsigned short*>(src.ptr) + (long long)rem * src.sample_stride +
                              (long long)(chan0 + c8 * 8) * src.chan_stride + soff + grp * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) t.in[j] = ldg128_stream(q + (long long)j * src.chan_stride);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) t.in[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  t.slab_addr = slab_addr;
  t.r0 = Lb * p.P[N - 1] + src.off[N - 1] + grp * 8 - sg.R0;
  t.c8 = c8;
  t.lo = sg.ra - sg.R0;
  t.hi = sg.rb - sg.R0;
}
// transpose the 8 channels x 8 positions block and store the 8 rows (16 bytes of 8 channels each)
__device__ __forceinline__ void store_task(const DataTask& t) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = t.r0 + i;
    const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 v = t.in[j];
      w[j] = (i >> 1) == 0 ? v.x : ((i >> 1) == 1 ? v.y : ((i >> 1) == 2 ? v.z : v.w));
    }
    if (rr >= t.lo && rr < t.hi)
      sts128(t.slab_addr + (uint32_t)rr * 128u + (uint32_t)((t.c8 ^ (rr & 7)) << 4), __byte_perm(w[0], w[1], sel),
             __byte_perm(w[2], w[3], sel), __byte_perm(w[4], w[5], sel), __byte_perm(w[6], w[7], sel));
  }
}
// padding rows of biased line Lb inside the segment: wp in [0, offw) and [offw + Sw, Pw)
__device__ __forceinline__ void pad_task(const Params& p, const Params::Src& src, const SegGeom& sg, uint32_t slab_addr,
                                         int Lb, int c8) {
  const int N = p.n_dims;
  const int Pw = p.P[N - 1], Sw = src.S[N - 1], offw = src.off[N - 1];
  const int row_base = Lb * Pw - sg.R0;
  const int lo = sg.ra - sg.R0, hi = sg.rb - sg.R0;
  for (int wp = 0; wp < Pw; ++wp) {
    if (wp == offw) wp += Sw;
    if (wp >= Pw) break;
    const int rr = row_base + wp;
    if (rr >= lo && rr < hi) sts128(slab_addr + (uint32_t)rr * 128u + (uint32_t)((c8 ^ (rr & 7)) << 4), 0u, 0u, 0u, 0u);
  }
}

// global-memory versions (SELF-PACKING mode): row r of the packed copy at dst + r * pitch elements
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void store_task_global(const DataTask& t, unsigned short* dst, int pitch) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = t.r0 + i;
    const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 v = t.in[j];
      w[j] = (i >> 1) == 0 ? v.x : ((i >> 1) == 1 ? v.y : ((i >> 1) == 2 ? v.z : v.w));
    }
    if (rr >= t.lo && rr < t.hi)
      stg128(dst + (long long)rr * pitch + t.c8 * 8, __byte_perm(w[0], w[1], sel), __byte_perm(w[2], w[3], sel),
             __byte_perm(w[4], w[5], sel), __byte_perm(w[6], w[7], sel));
  }
}
__device__ __forceinline__ void pad_task_global(const Params& p, const Params::Src& src, const SegGeom& sg,
                                                unsigned short* dst, int pitch, int Lb, int c8) {
  const int N = p.n_dims;
  const int Pw = p.P[N - 1], Sw = src.S[N - 1], offw = src.off[N - 1];
  const int row_base = Lb * Pw - sg.R0;
  const int lo = sg.ra - sg.R0, hi = sg.rb - sg.R0;
  for (int wp = 0; wp < Pw; ++wp) {
    if (wp == offw) wp += Sw;
    if (wp >= Pw) break;
    const int rr = row_base + wp;
    if (rr >= lo && rr < hi) stg128(dst + (long long)rr * pitch + c8 * 8, 0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy <-> async-proxy ordering for every state space (global memory read by TMA)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct Bars {
  uint64_t full[4], empty[4];
  uint64_t done;
  uint32_t tmem_base;
  volatile int progress;  // SELF-PACKING mode: chunk the TMA warp is reading (n_chunks when it is done)
};

// MODE 0: TMA boxes of packed copies written by a pre-pass; 1: FUSED producers; 2: SELF-PACKING
template <typename T, int MODE>
__global__ void __launch_bounds__(MODE ? kThreadsFused : kThreads, 1)
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
  constexpr bool FUSED = MODE == 1, SELFPACK = MODE == 2;
  // K splits: contiguous ranges of k-blocks, or (SELF-PACKING) k-blocks dealt round-robin so that all CTAs
  // move through the rows together
  const bool rr = SELFPACK || (p.round_robin && !FUSED);  // the FUSED overlap copy needs consecutive k-blocks
  const int kb_per = (p.n_kblocks + p.k_splits - 1) / p.k_splits;
  const int kb_begin = rr ? split : split * kb_per;
  const int kb_step = rr ? p.k_splits : 1;
  const int n_kb = rr ? (split < p.n_kblocks ? (p.n_kblocks - split + p.k_splits - 1) / p.k_splits : 0)
                      : max(0, min(p.n_kblocks, kb_begin + kb_per) - kb_begin);
  const int kb_end = kb_begin + n_kb * kb_step;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    mbar_init(&bars->done, 1);
    bars->progress = 0;
    fence_barrier_init();
    if constexpr (!FUSED) {
      prefetch_tensormap(&map_x);
      prefetch_tensormap(&map_v);
    }
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(p.n_pairs * p.acc_cols)) tmem_cols <<= 1;
  if (warp == 1) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (FUSED && warp >= 6) {
    // ------------------------------------------------------------------ software producers (FUSED)
    const int ptid = threadIdx.x - 6 * 32;
    constexpr int kProducers = kProducerWarps * 32;
    const int N = p.n_dims;
    const int bias = p.bias_lines * p.P[N - 1];
    const int chan_x = g * p.cg + cblock * kCB, nch_x = min(kCB, p.cg - cblock * kCB);
    const int keep_rows = p.slab_rows - p.kb_rows;  // rows shared with the previous k-block's slab
    int st = 0;
    uint32_t ph = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait_backoff(&bars->empty[st], ph ^ 1);
      const uint32_t stage = smem_u32(ring + (size_t)st * p.stage_bytes);
      const uint32_t vslab = stage + (uint32_t)p.slab_bytes;
      const int r0 = kb * p.kb_rows;
      const int keep = kb == kb_begin ? 0 : keep_rows;
      const int XR0 = r0 + p.group_shift[tg] + bias, VR0 = r0 - p.v_early + bias;
      const SegGeom sgx = seg_geom(p, p.sx, XR0, XR0 + keep, XR0 + p.slab_rows);
      const SegGeom sgv = seg_geom(p, p.sv, VR0, VR0, VR0 + p.v_rows);
      const int Tx = sgx.n_groups * 8, Tv = sgv.n_groups * 8, Tdata = Tx + p.n_vblocks * Tv;
      const int Px = sgx.n_lines * 8, Pv = sgv.n_lines * 8, Ttot = Tdata + Px + p.n_vblocks * Pv;
      bool copied = false;
      for (int task = ptid; task < Ttot || !copied; task += kProducers) {
        DataTask t;
        const bool data = task < Tdata;
        if (data) {
          if (task < Tx) {
            load_task(t, p, p.sx, sgx, stage, chan_x, nch_x, sgx.s0 + (task >> 3), task & 7);
          } else {
            const int tv = task - Tx, nbk = tv / Tv, tt = tv - nbk * Tv;
            load_task(t, p, p.sv, sgv, vslab + (uint32_t)(nbk * p.v_rows) * kRowBytes,
                      g * p.cog + cotile * p.bn + nbk * kCB, min(kCB, p.cog - cotile * p.bn - nbk * kCB),
                      sgv.s0 + (tt >> 3), tt & 7);
          }
        }
        if (!copied) {
          // rows [0, keep) of this slab = rows [kb_rows, kb_rows + keep) of the previous stage's slab; kb_rows
          // is a multiple of 8, so the swizzle phase of every row is unchanged: plain 16-byte copies
          copied = true;
          if (keep > 0) {
            const int prev = st == 0 ? p.n_stages - 1 : st - 1;
            const uint32_t from = smem_u32(ring + (size_t)prev * p.stage_bytes) + (uint32_t)p.kb_rows * kRowBytes;
            for (int c = ptid; c < keep * 8; c += kProducers) {
              const uint4 q = lds128(from + (uint32_t)c * 16u);
              sts128(stage + (uint32_t)c * 16u, q.x, q.y, q.z, q.w);
            }
          }
        }
        if (data) {
          store_task(t);
        } else if (task < Ttot) {
          int tp = task - Tdata;
          if (tp < Px) {
            pad_task(p, p.sx, sgx, stage, sgx.l0 + (tp >> 3), tp & 7);
          } else {
            tp -= Px;
            const int nbk = tp / Pv, tt = tp - nbk * Pv;
            pad_task(p, p.sv, sgv, vslab + (uint32_t)(nbk * p.v_rows) * kRowBytes, sgv.l0 + (tt >> 3), tt & 7);
          }
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      producers_sync();     // also orders this stage's stores before the next stage's overlap copy
      if (ptid == 0) mbar_arrive(&bars->full[st]);
      if (++st == p.n_stages) { st = 0; ph ^= 1; }
    }
  } else if (SELFPACK && warp >= 6) {
    // ------------------------------------------------------------------ packing warps (SELF-PACKING)
    const int ptid = threadIdx.x - 6 * 32;
    constexpr int kProducers = kProducerWarps * 32;
    const int N = p.n_dims;
    const int bias = p.bias_lines * p.P[N - 1];
    const int xblocks = (p.cxp + kCB - 1) / kCB, vblocks = (p.cvp + kCB - 1) / kCB;
    for (int c = 0; c < p.n_chunks; ++c) {
      // stay at most two chunks ahead of this CTA's reads: what is packed now should still be in L2 when read
      {
        long long n = 0;
        while (c > bars->progress + 2) {
          __nanosleep(256);
          if (++n > 100000000ll) __trap();
        }
      }
      const int c_begin = c * p.chunk_rows, c_end = min(p.rows_total, c_begin + p.chunk_rows);
      const int per = ((c_end - c_begin + (int)gridDim.x - 1) / (int)gridDim.x + 7) & ~7;
      const int ra = min(c_end, c_begin + (int)blockIdx.x * per), rb = min(c_end, ra + per);
      const SegGeom sgx = seg_geom(p, p.sx, bias, ra + bias, rb + bias);
      const SegGeom sgv = seg_geom(p, p.sv, bias, ra + bias, rb + bias);
      const int Tx1 = sgx.n_groups * 8, Tv1 = sgv.n_groups * 8, Tx = Tx1 * xblocks, Tdata = Tx + Tv1 * vblocks;
      const int Px1 = sgx.n_lines * 8, Pv1 = sgv.n_lines * 8, Px = Px1 * xblocks, Ttot = Tdata + Px + Pv1 * vblocks;
      for (int task = ptid; task < Ttot; task += kProducers) {
        if (task < Tdata) {
          DataTask t;
          if (task < Tx) {
            const int blk = task / Tx1, tt = task - blk * Tx1;
            if (blk * kCB + (tt & 7) * 8 < p.cxp) {
              load_task(t, p, p.sx, sgx, 0u, blk * kCB, min(kCB, p.sx.ctot - blk * kCB), sgx.s0 + (tt >> 3), tt & 7);
              store_task_global(t, p.xp + blk * kCB, p.cxp);
            }
          } else {
            const int tv = task - Tx, blk = tv / Tv1, tt = tv - blk * Tv1;
            if (blk * kCB + (tt & 7) * 8 < p.cvp) {
              load_task(t, p, p.sv, sgv, 0u, blk * kCB, min(kCB, p.sv.ctot - blk * kCB), sgv.s0 + (tt >> 3), tt & 7);
              store_task_global(t, p.vp + blk * kCB, p.cvp);
            }
          }
        } else {
          int tp = task - Tdata;
          if (tp < Px) {
            const int blk = tp / Px1, tt = tp - blk * Px1;
            if (blk * kCB + (tt & 7) * 8 < p.cxp)
              pad_task_global(p, p.sx, sgx, p.xp + blk * kCB, p.cxp, sgx.l0 + (tt >> 3), tt & 7);
          } else {
            tp -= Px;
            const int blk = tp / Pv1, tt = tp - blk * Pv1;
            if (blk * kCB + (tt & 7) * 8 < p.cvp)
              pad_task_global(p, p.sv, sgv, p.vp + blk * kCB, p.cvp, sgv.l0 + (tt >> 3), tt & 7);
          }
        }
      }
      fence_proxy_async_all();  // generic-proxy global writes -> TMA (async-proxy) reads of other CTAs
      __threadfence();
      producers_sync();
      if (ptid == 0) red_release_gpu_add(p.chunk_flags + c, 1u);
    }
  } else if (!FUSED && warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int st = 0;
    uint32_t ph = 0;
    int ready = -1;  // SELF-PACKING: chunks [0, ready] are known to be packed
    const uint32_t stage_tx = (uint32_t)(p.slab_rows * kRowBytes + p.n_vblocks * p.v_rows * kRowBytes);
    const int col_x = g * p.cg + cblock * kCB;
    const int col_v = g * p.cog + cotile * p.bn;
    for (int kb = kb_begin; kb < kb_end; kb += kb_step) {
      if constexpr (SELFPACK) {
        // the slab of this k-block ends at row r0 + group_shift + slab_rows: wait for the chunk that holds it
        const int need = min(p.n_chunks - 1, (kb * p.kb_rows + p.group_shift[tg] + p.slab_rows) / p.chunk_rows);
        if (need > ready) {
          if (lane == 0) {
            bars->progress = min(p.n_chunks, (kb * p.kb_rows) / p.chunk_rows);
            for (int c = ready + 1; c <= need; ++c) {
              long long n = 0;
              while (ld_acquire_gpu(p.chunk_flags + c) < gridDim.x) {
                __nanosleep(128);
                if (++n > 100000000ll) __trap();
              }
            }
          }
          __syncwarp();
          fence_proxy_async_all();
          ready = need;
        }
      }
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
            // multi-tap mode: the slab starts v_early rows early (negative rows of the first k-block zero-fill)
            tma_load_2d(vt + (size_t)(nb * p.v_rows + r) * kRowBytes, &map_v, &bars->full[st], col_v + nb * kCB,
                        r0 - p.v_early + r);
          }
        }
      }
      __syncwarp();
      if (++st == p.n_stages) { st = 0; ph ^= 1; }
    }
    if (SELFPACK && lane == 0) bars->progress = p.n_chunks;  // this CTA reads no more: pack the rest freely
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_mn(FMT, 128, p.acc_cols);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, slab16 = (uint32_t)p.slab_bytes >> 4;
    // distance between the 64-column blocks of B: the next channel block of vp, or (multi-tap mode) the
    // same channels blk_rows further down the slab
    const uint32_t vlbo16 = (uint32_t)((p.nb > 1 ? p.blk_rows : p.v_rows) * kRowBytes) >> 4;
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
  } else if (warp >= 2 && warp < 6) {
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
        const int nblock = p.nb > 1 ? (c0 >> 6) : 0;
        const int tl = p.tap_tbl[pr][unit][nblock];  // tap within the group
        const bool tap_ok = tl >= 0 && ci < p.cg;
        const int t = tg * p.tpg + tl;
        const int cbase = p.nb > 1 ? (c0 & 63) : c0;
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
  bool fused = false;     // MODE 1: slabs filled from the channels-first tensors by producer warps (opt-in)
  bool selfpack = false;  // MODE 2: packed copies written by the kernel's own producer warps, chunk by chunk
  int n_chunks = 0, chunk_rows = 0;
  int n_dims;
  long long P[kMaxDims], pitch[kMaxDims], ptot, rows_total;
  int cxp, cvp;
  int taps, tpg, n_tgroups, n_pairs;
  int nb = 1, upr = 1, blk_rows = 0, v_early = 0, v_rows = 0;  // multi-tap mode: see Params
  int kb_rows, slab_rows, slab_bytes, vtile_bytes, stage_bytes, n_stages;
  int bn, n_vblocks, n_cotiles, n_cblocks, n_kblocks, k_splits;
  long long xp_bytes, vp_bytes, partial_bytes, gw_elems;
  size_t smem;
};

// cycles of one M = 128 MMA over a 32-byte K step: the tensor pipe needs N / 2, shared memory
// (4 KB of A + 32 N bytes of B at 128 B/clk) 32 + N / 4 (measured: N = 64 48, N = 128 64, N = 256 128)
static inline double mma_cycles(int n) { return std::max(n / 2.0, 32.0 + n / 4.0); }

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
  // FUSED producer: 16-bit tensors whose channel planes and innermost rows start on 16-byte boundaries
  // (the pointers themselves are checked by the dispatcher), whole channel octets
  const bool cf_ok = dtype != EINCONV_F32 && g.in[N - 1] % 8 == 0 && g.out[N - 1] % 8 == 0 && g.c_in_g % 8 == 0 &&
                     g.c_out_g % 8 == 0;
  pl.fused = cf_ok && getenv("EINCONV_WGRAD_FUSED");
  const int row_round = pl.fused ? 8 : kBoxRows;  // TMA moves whole 64-row boxes

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
  // (MMAs * N <= 512 columns) and shared memory (two stages) allow; prefer larger k-blocks.
  for (int level = N; level >= 0 && !pl.ok; --level) {
    long long tpg = 1, span = 0;
    for (int d = N - level; d < N; ++d) {
      tpg *= g.k[d];
      span += (long long)g.dil[d] * (g.k[d] - 1) * pl.pitch[d];
    }
    if (tpg > kMaxTpg || pl.taps / tpg > kMaxGroups) continue;
    // candidates: nb = 1 (one tap per unit, N = BN), or multi-tap units of up to nb taps adjacent along the
    // innermost kernel dimension (N = 64 nb); two units per M = 128 MMA.  Cost = modelled cycles per 16
    // rows, with the shared-memory bytes as a tie-breaker (the producers write into the same memory).
    const int kw = level >= 1 ? g.k[N - 1] : 1;
    const long long q = (long long)g.dil[N - 1];
    int best_nb = 0, best_upr = 1, best_mmas = 0;
    double best_cost = 1e30;
    const int nb_forced = getenv("EINCONV_WGRAD_NB") ? atoi(getenv("EINCONV_WGRAD_NB")) : 0;
    for (int nb = 1; nb <= kMaxNb; ++nb) {
      if (nb_forced && nb != nb_forced && !(nb == 1)) continue;
      if (nb > 1 && (pl.bn != 64 || pl.n_vblocks != 1 || kw < 2 || nb > kw || (nb - 1) * q > kBoxRows ||
                     getenv("EINCONV_WGRAD_NO_QUAD")))
        continue;
      const int upr = nb == 1 ? kw : (kw + nb - 1) / nb;
      const int units = (int)(tpg / kw) * upr;
      const int mmas = (units + 1) / 2;
      const int n = nb == 1 ? pl.bn : 64 * nb;
      if ((long long)mmas * n > 512) continue;
      double cost = mmas * (mma_cycles(n) + 0.25 * (32.0 + n / 4.0));
      if (nb_forced && nb == nb_forced) cost = 0;
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        best_nb = nb;
        best_upr = upr;
        best_mmas = mmas;
      }
    }
    if (!best_nb) continue;
    const int kb_forced = getenv("EINCONV_WGRAD_KB") ? atoi(getenv("EINCONV_WGRAD_KB")) : 0;
    for (int kb : {256, 128, 64}) {
      if (kb_forced && kb != kb_forced) continue;
      const long long rows = (kb + span + row_round - 1) / row_round * row_round;
      const long long v_early = best_nb > 1 ? (best_nb - 1) * q : 0;
      const long long v_rows = pl.fused ? (kb + v_early + 7) / 8 * 8 : (v_early ? kb + kBoxRows : kb);
      const long long stage = rows * kRowBytes + (long long)pl.n_vblocks * v_rows * kRowBytes;
      if (2 * stage > kSmemBudget) continue;
      if (getenv("EINCONV_WGRAD_MIN3") && 3 * stage > kSmemBudget && kb > 64) continue;  // at least three stages
      pl.tpg = (int)tpg;
      pl.n_pairs = best_mmas;
      pl.nb = best_nb;
      pl.upr = best_upr;
      pl.blk_rows = best_nb > 1 ? (int)q : 0;
      pl.v_early = (int)v_early;
      pl.v_rows = (int)v_rows;
      pl.n_tgroups = pl.taps / pl.tpg;
      pl.kb_rows = kb;
      pl.slab_rows = (int)rows;
      pl.slab_bytes = (int)(rows * kRowBytes);
      pl.vtile_bytes = (int)(pl.n_vblocks * v_rows * kRowBytes);
      pl.stage_bytes = (int)stage;
      pl.n_stages = (int)std::min<long long>(4, kSmemBudget / stage);
      pl.ok = true;
      break;
    }
  }
  if (!pl.ok) return pl;
  // multi-tap mode reads vp[r - v_early]: r must run v_early past the last row
  pl.n_kblocks = (int)((pl.rows_total + pl.v_early + pl.kb_rows - 1) / pl.kb_rows);
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
  // SELF-PACKING (opt-in): every CTA must be resident (each one waits for chunks that all of them pack),
  // chunks longer than the reach of a slab
  if (cf_ok && !pl.fused && getenv("EINCONV_WGRAD_SELFPACK") && types * pl.k_splits <= num_sms()) {
    int nc = 16;
    while (nc >= 1) {
      const long long cr = ((pl.rows_total + nc - 1) / nc + 7) / 8 * 8;
      if (cr >= max_shift + pl.slab_rows + pl.kb_rows) {
        pl.selfpack = true;
        pl.n_chunks = (int)((pl.rows_total + cr - 1) / cr);
        pl.chunk_rows = (int)cr;
        break;
      }
      nc /= 2;
    }
  }
  pl.xp_bytes = pl.fused ? 0 : (((pl.rows_total + max_shift + 2 * kBoxRows) * pl.cxp * 2) + 1023) / 1024 * 1024;
  pl.vp_bytes = pl.fused ? 0 : ((pl.rows_total * pl.cvp * 2) + 1023) / 1024 * 1024;
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
  if (pl.fused) {
    // no packed copies: the kernel's producer warps read x and v themselves
  } else if (pl.selfpack) {
    // the kernel packs xp / vp itself, chunk by chunk, ahead of its TMA reads; per-chunk completion counters
    cudaError_t e = cudaMemsetAsync(absmax + 16, 0, 256, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(chunk flags)");
  } else if (dtype == EINCONV_F32) {
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
  memset(&map_x, 0, sizeof(map_x));
  memset(&map_v, 0, sizeof(map_v));
  if (!pl.fused) {
    st = ctma::encode_2d(&map_x, op_dtype, xp, pl.rows_total, pl.cxp, kCB, kBoxRows, 128);
    if (st) return st;
    st = ctma::encode_2d(&map_v, op_dtype, vp, pl.rows_total, pl.cvp, kCB, kBoxRows, 128);
    if (st) return st;
  }

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
  p.nb = pl.nb;
  p.blk_rows = pl.blk_rows;
  p.v_early = pl.v_early;
  p.v_rows = pl.v_rows;
  p.acc_cols = pl.nb > 1 ? 64 * pl.nb : pl.bn;
  // units: the M = 64 halves of the MMAs.  nb = 1: one tap each.  Multi-tap mode: unit i of a kernel row
  // holds the taps kw = i * nb .. i * nb + nb - 1 of the innermost dimension; N block j pairs the unit's
  // rows with vp[r - (nb - 1 - j) * blk_rows], i.e. tap kw = i * nb + (nb - 1 - j); the unit's slab offset
  // is the shift of its first tap (N block nb - 1, the un-shifted vp).
  int unit_tap[kMaxTpg][kMaxNb], n_units = 0;
  if (pl.nb > 1) {
    const int kw = g.k[N - 1];
    for (int row = 0; row < pl.tpg / kw; ++row)
      for (int i = 0; i < pl.upr; ++i) {
        for (int j = 0; j < kMaxNb; ++j) {
          const int k = i * pl.nb + (pl.nb - 1 - j);
          unit_tap[n_units][j] = (j < pl.nb && k < kw) ? row * kw + k : -1;
        }
        ++n_units;
      }
  } else {
    for (int t = 0; t < pl.tpg; ++t) {
      for (int j = 0; j < kMaxNb; ++j) unit_tap[n_units][j] = j == 0 ? t : -1;
      ++n_units;
    }
  }
  EINCONV_CHECK_ARG((n_units + 1) / 2 == pl.n_pairs, "wgrad_slab: inconsistent plan");
  for (int pr = 0; pr < pl.n_pairs; ++pr) {
    const int u1 = 2 * pr, u2 = 2 * pr + 1 < n_units ? 2 * pr + 1 : -1;
    const int first = pl.nb - 1;  // the N block that holds the unit's un-shifted tap
    const long long s1 = tap_shift[unit_tap[u1][first]];
    const long long s2 = u2 >= 0 ? tap_shift[unit_tap[u2][first]] : s1;
    p.pair_off16[pr] = (uint32_t)((s1 * kRowBytes) >> 4);
    p.pair_lbo16[pr] = (uint32_t)(((s2 - s1) * kRowBytes) >> 4);
    EINCONV_CHECK_ARG(s2 >= s1 && (s2 - s1) * kRowBytes < (1 << 18),
                      "wgrad_slab: tap distance exceeds the descriptor range");
    for (int nb = 0; nb < kMaxNb; ++nb) {
      p.tap_tbl[pr][0][nb] = (signed char)unit_tap[u1][nb];
      p.tap_tbl[pr][1][nb] = (signed char)(u2 >= 0 ? unit_tap[u2][nb] : -1);
    }
  }
  p.partial = partial;
  p.gw_elems = pl.gw_elems;
  p.n_dims = N;
  p.batch = g.batch;
  for (int d = 0; d < N; ++d) {
    p.P[d] = (int)pl.P[d];
    p.P_div[d] = FastDiv((uint32_t)pl.P[d]);
  }
  p.bias_lines = (int)((pl.v_early + pl.P[N - 1] - 1) / pl.P[N - 1]);
  auto fill_src = [&](Params::Src& sp, const void* ptr, const int* S, const int* off, int ctot) {
    sp.ptr = ptr;
    sp.ctot = ctot;
    long long st_ = 1;
    for (int d = N - 1; d >= 0; --d) {
      sp.S[d] = S[d];
      sp.off[d] = off[d];
      sp.sstride[d] = st_;
      st_ *= S[d];
    }
    sp.chan_stride = st_;
    sp.sample_stride = st_ * ctot;
    sp.groups_w = S[N - 1] / 8;
    sp.gw_div = FastDiv((uint32_t)std::max(1, S[N - 1] / 8));
  };
  fill_src(p.sx, x, g.in, g.pad, g.groups * g.c_in_g);
  fill_src(p.sv, v, g.out, zero_pad, g.groups * g.c_out_g);

  p.xp = reinterpret_cast<unsigned short*>(xp);
  p.vp = reinterpret_cast<unsigned short*>(vp);
  p.chunk_flags = absmax + 16;
  p.n_chunks = pl.n_chunks;
  p.chunk_rows = pl.chunk_rows;
  p.cxp = pl.cxp;
  p.cvp = pl.cvp;
  p.rows_total = (int)pl.rows_total;
  p.round_robin = getenv("EINCONV_WGRAD_CONTIGUOUS_SPLITS") ? 0 : 1;
  const long long grid = (long long)g.groups * pl.n_cblocks * pl.n_tgroups * pl.n_cotiles * pl.k_splits;
  if (pl.selfpack) {
    auto kern = wgrad_slab_kernel<T, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad_slab self-packing)");
    kern<<<(unsigned)grid, kThreadsFused, pl.smem, stream>>>(map_x, map_v, p);
  } else if (pl.fused) {
    auto kern = wgrad_slab_kernel<T, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad_slab fused)");
    kern<<<(unsigned)grid, kThreadsFused, pl.smem, stream>>>(map_x, map_v, p);
  } else {
    auto kern = wgrad_slab_kernel<T, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad_slab)");
    kern<<<(unsigned)grid, kThreads, pl.smem, stream>>>(map_x, map_v, p);
  }
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
