// tcgen05 / TMEM tensor-core evaluation of the dense contraction networks: convNd forward
// (expressions/convNd_forward.py:48-55), weight VJP (expressions/convNd_weight_vjp.py:51-58) and
// the KFC factor (expressions/convNd_kfc.py:75-90), for 1-3 spatial dims, any stride / padding /
// dilation / groups.  bf16 / fp16 inputs run kind::f16, fp32 inputs kind::tf32 (operands rounded to
// tf32 with cvt.rna while staging); accumulation is fp32 in tensor memory.
//
// Implicit GEMM.  Nothing is unfolded in global memory and no index pattern exists: producer warps
// evaluate i = s*o + d*k - p per element and write the operand tiles straight into shared memory in
// the canonical UMMA "interleaved" (no-swizzle) layout, one 16-byte chunk per thread:
//
//     image[chunk][line] (16 B each)   chunk = 16 B worth of consecutive indices of the "chunked"
//                                      dimension, line = index of the other dimension
//
// The lanes of a warp always walk `line` = spatial positions, which are contiguous in global memory
// (channels-first layout), so every gathered load is coalesced:
//   forward     D[pos, co]      A: line = pos (M),  chunk over k = (ci,tap)  -> K-major  descriptor
//   weight VJP  D[(ci,tap), co] A: line = pos (K),  chunk over (ci,tap) (M)  -> MN-major descriptor
//   KFC         D[(ci,tap), (ci',tap')]  both operands as weight-VJP's A
// (one image, two readings: K-major   LBO = LINES*16 B between K chunks,  SBO = 128 B between 8-line groups;
//                           MN-major  SBO = LINES*16 B between MN chunks, LBO = 128 B between 8-line groups.)
//
// Warp roles (160 threads): warps 0-3 gather operands into a 4-stage shared-memory ring and later
// drain the accumulator (TMEM lanes 32w..32w+31 belong to warp w); one elected thread of warp 4
// issues tcgen05.mma (M=128, N<=128, K=16 bf16 / 8 tf32) and releases ring slots with tcgen05.commit.
// Full/empty mbarriers connect the two; every spin has a clock-based timeout that traps instead of
// hanging the GPU.
#include <algorithm>
#include <type_traits>

#include "dispatch.cuh"

namespace einconv {
namespace tc {

constexpr int kStages = 4;
constexpr int kTileM = 128;            // UMMA M (cta_group::1)
constexpr int kProducerThreads = 128;  // warps 0-3
constexpr int kThreads = 160;          // + MMA warp
constexpr int kChunksPerStage = 8;     // 16-byte chunks along the chunked dim per stage
constexpr int ND = 3;

enum Mode { kFwd = 0, kWgrad = 1, kKfc = 2 };

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box.  ~2 s at 2 GHz, then trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

// shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start >> 4   [16,30) LBO >> 4   [32,46) SBO >> 4   [46,48) version = 1   [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format @7/@10
// (F16 = 0, BF16 = 1, TF32 = 2), a/b_major @15/@16 (0 = K, 1 = MN), N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int a_mn, int b_mn, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) |
         ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// operand element types: what is stored in shared memory
// ---------------------------------------------------------------------------------------------
template <typename T> struct Op;  // EPC = elements per 16-byte chunk
template <> struct Op<float> {
  static constexpr int EPC = 4;
  static constexpr bool TF32 = true;
  static constexpr int FMT = 2;
  static __device__ __forceinline__ uint32_t bits(const float* p) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(__ldg(p)));
    return r;
  }
  static __device__ __forceinline__ uint4 pack(const uint32_t* e) { return make_uint4(e[0], e[1], e[2], e[3]); }
};
template <> struct Op<__nv_bfloat16> {
  static constexpr int EPC = 8;
  static constexpr bool TF32 = false;
  static constexpr int FMT = 1;
  static __device__ __forceinline__ uint32_t bits(const __nv_bfloat16* p) {
    return __ldg(reinterpret_cast<const unsigned short*>(p));
  }
  static __device__ __forceinline__ uint4 pack(const uint32_t* e) {
    return make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
  }
};
template <> struct Op<__half> {
  static constexpr int EPC = 8;
  static constexpr bool TF32 = false;
  static constexpr int FMT = 0;
  static __device__ __forceinline__ uint32_t bits(const __half* p) {
    return __ldg(reinterpret_cast<const unsigned short*>(p));
  }
  static __device__ __forceinline__ uint4 pack(const uint32_t* e) { return Op<__nv_bfloat16>::pack(e); }
};

// ---------------------------------------------------------------------------------------------
// problem description
// ---------------------------------------------------------------------------------------------
struct PosEntry {  // an output position (b, o..) in input coordinates
  int base;        // b * batch_stride + sum_d (s_d*o_d - p_d) * in_stride_d        (may be negative)
  short v[ND];     // s_d*o_d - p_d
  short valid;
};
struct RowEntry {  // a (channel, tap) pair
  int off;         // ch * in_total + sum_d dil_d*k_d * in_stride_d
  short dk[ND];    // dil_d * k_d
  short valid;
};

struct Params {
  DevGeom g;
  int in_stride[ND];
  int n_pos;        // B * out_total
  int n_rows;       // c_in_g * k_total          (rows of the unfolded input per group)
  int c_out_g;
  int x_batch_stride, x_group_stride;   // elements
  int v_batch_stride, v_group_stride;   // elements (cotangent / output tensor)
  int k_splits, k_blocks;               // reduction blocks (of BK) in total and splits of them
  int m_tiles, n_tiles;
  const void* x;    // input
  const void* w;    // forward: weight
  const void* v;    // weight VJP: cotangent
  const void* bias;
  void* out;        // forward: y (T);  weight VJP / KFC: fp32 partials [G][splits][M][N]
  float scale;
};

template <int NDIMS>
__device__ __forceinline__ PosEntry make_pos_entry(const Params& p, int idx) {
  PosEntry e;
  const DevGeom& g = p.g;
  e.valid = idx < p.n_pos;
  uint32_t rem = e.valid ? (uint32_t)idx : 0u;
  int base = 0;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      uint32_t q, o;
      g.out_div[d].divmod(rem, q, o);
      rem = q;
      const int v = g.stride[d] * (int)o - g.pad[d];
      e.v[d] = (short)v;
      base += v * p.in_stride[d];
    } else {
      e.v[d] = 0;
    }
  }
  e.base = base + (int)rem * p.x_batch_stride;
  return e;
}

__device__ __forceinline__ RowEntry make_row_entry(const Params& p, int idx) {
  RowEntry e;
  const DevGeom& g = p.g;
  e.valid = idx < p.n_rows;
  uint32_t rem = e.valid ? (uint32_t)idx : 0u;
  int off = 0;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      uint32_t q, k;
      g.k_div[d].divmod(rem, q, k);
      rem = q;
      const int dk = g.dil[d] * (int)k;
      e.dk[d] = (short)dk;
      off += dk * p.in_stride[d];
    } else {
      e.dk[d] = 0;
    }
  }
  e.off = off + (int)rem * g.in_total;
  return e;
}

template <typename T>
__device__ __forceinline__ uint32_t gather_x(const T* __restrict__ x, const DevGeom& g, const PosEntry& pe,
                                             const RowEntry& re) {
  bool ok = pe.valid && re.valid;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    if (d < g.n_dims) {
      const int i = (int)pe.v[d] + (int)re.dk[d];
      ok = ok && (i >= 0) && (i < g.in[d]);
    }
  }
  uint32_t r = 0;
  if (ok) r = Op<T>::bits(x + pe.base + re.off);
  return r;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int BN>
struct Smem {
  // forward:               A image = 8 K-chunks x 128 lines (positions), B image = 8 K-chunks x BN lines (co)
  // wgrad / kfc, 16-bit:    MN-major images: [rows / EPC chunks][BKpos lines (positions)]
  // wgrad / kfc, tf32:      K-major images (tf32 MN-major operands are not usable): [BKpos / 4 chunks][rows]
  //                         with a 64-byte pad per chunk so the transposing stores are conflict-free
  static constexpr int BKpos = Op<T>::TF32 ? 32 : 64;  // reduction positions per stage (wgrad/kfc)
  static constexpr bool kTransposed = (MODE != kFwd) && Op<T>::TF32;
  static constexpr int a_chunk_stride = (MODE == kFwd) ? kTileM * 16 : (kTransposed ? kTileM * 16 + 64 : BKpos * 16);
  static constexpr int b_chunk_stride = (MODE == kFwd) ? BN * 16 : (kTransposed ? BN * 16 + 64 : BKpos * 16);
  static constexpr int a_bytes = (MODE == kFwd) ? kChunksPerStage * a_chunk_stride
                                 : kTransposed  ? (BKpos / 4) * a_chunk_stride
                                                : (kTileM / Op<T>::EPC) * a_chunk_stride;
  static constexpr int b_bytes = (MODE == kFwd) ? kChunksPerStage * b_chunk_stride
                                 : kTransposed  ? (BKpos / 4) * b_chunk_stride
                                                : (BN / Op<T>::EPC) * b_chunk_stride;
  static constexpr int stage_bytes = ((a_bytes + b_bytes + 127) / 128) * 128;
  static constexpr int table_bytes = 512 * 16;  // row / tap tables
  static constexpr int total = kStages * stage_bytes + table_bytes + 256;
};

// 4x4 transpose inside each quad of lanes: in: e[r] = value of row r at this lane's position;
// out: e[q] = value of row (lane & 3) at the position of lane (lane & ~3) + q
__device__ __forceinline__ void quad_transpose(uint32_t (&e)[4], int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  // exchange across lane ^ 1: swap the off-diagonal 1x1 blocks of each 2x2 block
  uint32_t s0 = b0 ? e[0] : e[1], s1 = b0 ? e[2] : e[3];
  s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
  s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  if (b0) { e[0] = s0; e[2] = s1; } else { e[1] = s0; e[3] = s1; }
  // exchange across lane ^ 2: swap the off-diagonal 2x2 blocks
  uint32_t t0 = b1 ? e[0] : e[2], t1 = b1 ? e[1] : e[3];
  t0 = __shfl_xor_sync(0xffffffffu, t0, 2);
  t1 = __shfl_xor_sync(0xffffffffu, t1, 2);
  if (b1) { e[0] = t0; e[1] = t1; } else { e[2] = t0; e[3] = t1; }
}

template <typename T, int MODE, int BN>
__global__ void __launch_bounds__(kThreads, 1) tc_gather_gemm_kernel(const __grid_constant__ Params p) {
  using O = Op<T>;
  using S = Smem<T, MODE, BN>;
  constexpr int EPC = O::EPC;
  constexpr int UMMA_K = O::TF32 ? 8 : 16;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage_base = smem;
  RowEntry* row_tab = reinterpret_cast<RowEntry*>(smem + kStages * S::stage_bytes);  // 512 entries x 16 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::stage_bytes + S::table_bytes);
  uint64_t* full_bar = bars;                  // [kStages], 128 producer arrivals each
  uint64_t* empty_bar = bars + kStages;       // [kStages], 1 arrival (tcgen05.commit)
  uint64_t* accum_bar = bars + 2 * kStages;   // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const DevGeom& g = p.g;

  // ---- tile coordinates -------------------------------------------------------------------
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int grp = blockIdx.z / p.k_splits;
  const int split = blockIdx.z - grp * p.k_splits;
  const int kb_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(p.k_blocks, kb_begin + kb_per_split);
  const int n_kb = max(0, kb_end - kb_begin);

  // ---- one-time setup -----------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], kProducerThreads);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const T* x = reinterpret_cast<const T*>(p.x) + (long long)grp * p.x_group_stride;

  if (warp < 4) {
    // =============================== producers ===============================================
    if constexpr (MODE == kFwd) {
      // A: line = output position (this thread's), chunks over k = (ci, tap); B: line = co, chunks over k
      const PosEntry pe = make_pos_entry<ND>(p, m_tile * kTileM + tid);
      const T* w = reinterpret_cast<const T*>(p.w) + (long long)grp * p.c_out_g * p.n_rows;
      constexpr int BK = kChunksPerStage * EPC;
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (uint32_t)(i / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int k0 = (kb_begin + i) * BK;
        // row table for this stage: BK (ci,tap) entries, written by the first BK producer threads
        // into the stage's slice of the table (double use is safe: slot reuse is ordered by empty_bar)
        RowEntry* tab = row_tab + s * 64;
        if (tid < BK) tab[tid] = make_row_entry(p, k0 + tid);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        unsigned char* a_img = stage_base + s * S::stage_bytes;
        unsigned char* b_img = a_img + S::a_bytes;
#pragma unroll
        for (int c = 0; c < kChunksPerStage; ++c) {
          uint32_t e[EPC];
#pragma unroll
          for (int j = 0; j < EPC; ++j) e[j] = gather_x<T>(x, g, pe, tab[c * EPC + j]);
          *reinterpret_cast<uint4*>(a_img + c * S::a_chunk_stride + tid * 16) = O::pack(e);
        }
        // B image: BN lines x 8 chunks, 128 threads: thread -> (line = tid % BN, chunk parity)
        for (int item = tid; item < BN * kChunksPerStage; item += kProducerThreads) {
          const int line = item % BN, c = item / BN;
          const int n = n_tile * BN + line;
          uint32_t e[EPC];
#pragma unroll
          for (int j = 0; j < EPC; ++j) {
            const int k = k0 + c * EPC + j;
            e[j] = (n < p.c_out_g && k < p.n_rows) ? O::bits(w + (long long)n * p.n_rows + k) : 0u;
          }
          *reinterpret_cast<uint4*>(b_img + c * S::b_chunk_stride + line * 16) = O::pack(e);
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    } else {
      // weight VJP / KFC: line = reduction position (b, o), chunks over operand rows.
      // A rows: (ci, tap) of this M tile; B rows: co (weight VJP, read from v) or (ci', tap') (KFC).
      constexpr int BKpos = S::BKpos;
      constexpr int a_chunks = kTileM / EPC, b_chunks = BN / EPC;
      // row tables, once per CTA: entries [0,128) = A rows, [128, 128+BN) = B rows (KFC)
      for (int r = tid; r < kTileM; r += kProducerThreads) row_tab[r] = make_row_entry(p, m_tile * kTileM + r);
      if constexpr (MODE == kKfc) {
        for (int r = tid; r < BN; r += kProducerThreads) row_tab[kTileM + r] = make_row_entry(p, n_tile * BN + r);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const T* vptr = reinterpret_cast<const T*>(p.v) + (long long)grp * p.v_group_stride;
      // thread -> (position line = tid % BKpos, chunk phase = tid / BKpos)
      const int line = tid % BKpos;
      const int phase = tid / BKpos;
      constexpr int n_phase = kProducerThreads / BKpos;
      // Chunk c holds rows [c*EPC, (c+1)*EPC) at this thread's position.
      //  16-bit: MN-major image, one 16-byte unit per (row chunk, position).
      //  tf32:   transpose 4 rows x 4 positions inside the lane quad and store K-major:
      //          unit (position chunk = line / 4, row = 4c + lane % 4).
      auto store_chunk = [&](unsigned char* img, int chunk_stride, int c, uint32_t (&e)[EPC]) {
        if constexpr (S::kTransposed) {
          quad_transpose(e, tid & 31);
          *reinterpret_cast<uint4*>(img + (line >> 2) * chunk_stride + (c * 4 + (tid & 3)) * 16) = O::pack(e);
        } else {
          *reinterpret_cast<uint4*>(img + c * chunk_stride + line * 16) = O::pack(e);
        }
      };
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (uint32_t)(i / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int pos = (kb_begin + i) * BKpos + line;
        const PosEntry pe = make_pos_entry<ND>(p, pos);
        unsigned char* a_img = stage_base + s * S::stage_bytes;
        unsigned char* b_img = a_img + S::a_bytes;
        for (int c = phase; c < a_chunks; c += n_phase) {
          uint32_t e[EPC];
#pragma unroll
          for (int j = 0; j < EPC; ++j) e[j] = gather_x<T>(x, g, pe, row_tab[c * EPC + j]);
          store_chunk(a_img, S::a_chunk_stride, c, e);
        }
        if constexpr (MODE == kKfc) {
          for (int c = phase; c < b_chunks; c += n_phase) {
            uint32_t e[EPC];
#pragma unroll
            for (int j = 0; j < EPC; ++j) e[j] = gather_x<T>(x, g, pe, row_tab[kTileM + c * EPC + j]);
            store_chunk(b_img, S::b_chunk_stride, c, e);
          }
        } else {
          // cotangent v[b, co, o]: position -> (b, oflat)
          uint32_t bq = 0, of = 0;
          if (pos < p.n_pos) {
            bq = (uint32_t)pos / (uint32_t)g.out_total;
            of = (uint32_t)pos - bq * (uint32_t)g.out_total;
          }
          const T* vb = vptr + (long long)bq * p.v_batch_stride + of;
          for (int c = phase; c < b_chunks; c += n_phase) {
            uint32_t e[EPC];
#pragma unroll
            for (int j = 0; j < EPC; ++j) {
              const int co = n_tile * BN + c * EPC + j;
              e[j] = (pos < p.n_pos && co < p.c_out_g) ? O::bits(vb + (long long)co * g.out_total) : 0u;
            }
            store_chunk(b_img, S::b_chunk_stride, c, e);
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    }

    // =============================== epilogue ================================================
    if (n_kb > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int m = m_tile * kTileM + tid;  // D row held by this thread
#pragma unroll 1
    for (int n0 = 0; n0 < BN; n0 += 16) {
      uint32_t r[16];
      if (n_kb > 0) {
        tmem_ld16(lane_base + n0, r);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if constexpr (MODE == kFwd) {
        // y[b, grp*Cout_g + n, o]  — consecutive threads = consecutive positions: coalesced
        if (m < p.n_pos) {
          const int b = m / g.out_total, o = m - b * g.out_total;
          T* y = reinterpret_cast<T*>(p.out) + (long long)b * p.v_batch_stride +
                 (long long)grp * p.v_group_stride + o;
          const T* bias = reinterpret_cast<const T*>(p.bias);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n_tile * BN + n0 + j;
            if (n < p.c_out_g) {
              float val = __uint_as_float(r[j]);
              if (bias) val += (float)Cvt<T>::load(bias + grp * p.c_out_g + n);
              Cvt<T>::store(y + (long long)n * g.out_total, val);
            }
          }
        }
      } else {
        // fp32 partials [grp][split][M][N]: this thread owns 16 consecutive n of row m
        const int M = p.n_rows;
        const int N = (MODE == kKfc) ? p.n_rows : p.c_out_g;
        if (m < M) {
          float* dst = reinterpret_cast<float*>(p.out) +
                       (((long long)grp * p.k_splits + split) * M + m) * N + n_tile * BN + n0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n_tile * BN + n0 + j < N) dst[j] = __uint_as_float(r[j]);
        }
      }
    }
    tc_fence_before();
  } else {
    // =============================== MMA issuer (warp 4) =====================================
    constexpr bool kMajorK = (MODE == kFwd) || S::kTransposed;
    constexpr uint32_t idesc = make_idesc(O::FMT, kMajorK ? 0 : 1, kMajorK ? 0 : 1, kTileM, BN);
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % kStages;
      const uint32_t ph = (uint32_t)(i / kStages) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if ((tid & 31) == 0) {
        const uint32_t a_addr = smem_u32(stage_base + s * S::stage_bytes);
        const uint32_t b_addr = a_addr + S::a_bytes;
        if constexpr (kMajorK) {
          // K-major: one MMA consumes 2 chunks; LBO = chunk stride, SBO = 128 B between 8-line groups
          constexpr int n_chunks = (MODE == kFwd) ? kChunksPerStage : S::BKpos / 4;
          constexpr int n_mma = n_chunks / 2;
#pragma unroll
          for (int j = 0; j < n_mma; ++j) {
            const uint64_t da = make_desc(a_addr + j * 2 * S::a_chunk_stride, S::a_chunk_stride, 128);
            const uint64_t db = make_desc(b_addr + j * 2 * S::b_chunk_stride, S::b_chunk_stride, 128);
            umma<O::TF32>(tmem_base, da, db, idesc, (i > 0 || j > 0) ? 1u : 0u);
          }
        } else {
          // MN-major: one MMA consumes UMMA_K positions = UMMA_K lines; SBO = MN chunk stride,
          // LBO = 128 B between groups of 8 lines
          constexpr int n_mma = S::BKpos / UMMA_K;
#pragma unroll
          for (int j = 0; j < n_mma; ++j) {
            const uint64_t da = make_desc(a_addr + j * UMMA_K * 16, 128, S::a_chunk_stride);
            const uint64_t db = make_desc(b_addr + j * UMMA_K * 16, 128, S::b_chunk_stride);
            umma<O::TF32>(tmem_base, da, db, idesc, (i > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);                 // frees the ring slot when these MMAs retire
        if (i == n_kb - 1) umma_commit(accum_bar);  // accumulator complete
      }
      __syncwarp();
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// split-K reduction + scale + cast (and mirror-free: KFC computes the full matrix)
template <typename T>
__global__ void __launch_bounds__(256) tc_reduce_kernel(const float* __restrict__ partial, T* __restrict__ out,
                                                        long long per_group, int k_splits, int groups,
                                                        float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_group * groups) return;
  const long long grp = idx / per_group, i = idx - grp * per_group;
  const float* src = partial + grp * k_splits * per_group + i;
  float s = 0.f;
  for (int k = 0; k < k_splits; ++k) s += src[(long long)k * per_group];
  Cvt<T>::store(out + idx, s * scale);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace tc;

static bool tc_dtype_ok(int dtype, int math) {
  if (math == EINCONV_MATH_TF32) return dtype == EINCONV_F32;
  if (math == EINCONV_MATH_HALF) return dtype == EINCONV_BF16 || dtype == EINCONV_F16;
  return false;
}

static bool fits_short(int v) { return v > -32768 && v < 32767; }

bool tc_supported(int op, const DevGeom& g, int dtype, int math) {
  if (op != EINCONV_OP_CONV_FWD && op != EINCONV_OP_CONV_WEIGHT_VJP && op != EINCONV_OP_KFC) return false;
  if (!tc_dtype_ok(dtype, math)) return false;
  if (g.n_dims > ND) return false;
  const long long x_elems = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
  const long long y_elems = (long long)g.batch * g.groups * std::max(1, g.c_out_g) * g.out_total;
  if (x_elems >= (1ll << 31) || y_elems >= (1ll << 31)) return false;
  if ((long long)g.batch * g.out_total >= (1ll << 31)) return false;
  if ((long long)g.c_in_g * g.k_total >= (1ll << 24)) return false;
  for (int d = 0; d < g.n_dims; ++d) {
    if (!fits_short(g.in[d] + g.pad[d] + g.dil[d] * g.k[d]) || !fits_short(g.stride[d] * g.out[d] + g.pad[d]))
      return false;
  }
  if (g.batch == 0 || g.out_total == 0 || g.c_in_g == 0) return false;
  // tiny problems are launch-bound either way; keep them on the exact fp32 path
  if (op == EINCONV_OP_CONV_FWD && g.c_out_g < 8) return false;
  return true;
}

const char* tc_kernel_name(int op, const DevGeom&, int, int math) {
  const bool tf32 = math == EINCONV_MATH_TF32;
  switch (op) {
    case EINCONV_OP_CONV_FWD: return tf32 ? "tcgen05_fwd_tf32" : "tcgen05_fwd_f16";
    case EINCONV_OP_CONV_WEIGHT_VJP: return tf32 ? "tcgen05_wgrad_tf32" : "tcgen05_wgrad_f16";
    default: return tf32 ? "tcgen05_kfc_tf32" : "tcgen05_kfc_f16";
  }
}

struct TcPlan {
  int bn, m_tiles, n_tiles, k_blocks, k_splits;
  long long M, N;
};

static TcPlan tc_plan(int op, const DevGeom& g, int dtype) {
  TcPlan pl{};
  const int epc = dtype == EINCONV_F32 ? 4 : 8;
  const long long rows = (long long)g.c_in_g * g.k_total;
  const long long pos = (long long)g.batch * g.out_total;
  if (op == EINCONV_OP_CONV_FWD) {
    pl.M = pos; pl.N = g.c_out_g;
    pl.k_blocks = ceil_div(rows, kChunksPerStage * epc);
    pl.k_splits = 1;
  } else {
    pl.M = rows; pl.N = (op == EINCONV_OP_KFC) ? rows : g.c_out_g;
    pl.k_blocks = ceil_div(pos, dtype == EINCONV_F32 ? 32 : 64);
  }
  pl.bn = pl.N > 64 ? 128 : 64;
  pl.m_tiles = ceil_div(pl.M, kTileM);
  pl.n_tiles = ceil_div(pl.N, pl.bn);
  if (op != EINCONV_OP_CONV_FWD) {
    const long long tiles = (long long)pl.m_tiles * pl.n_tiles * g.groups;
    long long want = (2ll * num_sms() + tiles - 1) / tiles;
    long long cap = std::max(1, pl.k_blocks / 8);
    pl.k_splits = (int)std::max<long long>(1, std::min<long long>(std::min(want, cap), 256));
  }
  return pl;
}

int64_t tc_workspace_bytes(int op, const DevGeom& g, int dtype, int) {
  if (op == EINCONV_OP_CONV_FWD) return 0;
  const TcPlan pl = tc_plan(op, g, dtype);
  return (int64_t)g.groups * pl.k_splits * pl.M * pl.N * 4;
}

template <typename T, int MODE, int BN>
static int tc_launch(const Params& p, const TcPlan& pl, int groups, cudaStream_t stream, const char* name) {
  using S = Smem<T, MODE, BN>;
  auto kern = tc_gather_gemm_kernel<T, MODE, BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gather_gemm)");
  dim3 grid(pl.m_tiles, pl.n_tiles, groups * pl.k_splits);
  EINCONV_CHECK_ARG(grid.y < 65536 && grid.z < 65536, "%s: too many tiles", name);
  kern<<<grid, kThreads, S::total, stream>>>(p);
  EINCONV_CHECK_LAUNCH(name);
  return EINCONV_OK;
}

static void fill_params(Params& p, const DevGeom& g, const TcPlan& pl) {
  p.g = g;
  int s = 1;
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      p.in_stride[d] = s;
      s *= g.in[d];
    } else {
      p.in_stride[d] = 0;
    }
  }
  p.n_pos = g.batch * g.out_total;
  p.n_rows = g.c_in_g * g.k_total;
  p.c_out_g = g.c_out_g;
  p.x_batch_stride = g.groups * g.c_in_g * g.in_total;
  p.x_group_stride = g.c_in_g * g.in_total;
  p.v_batch_stride = g.groups * g.c_out_g * g.out_total;
  p.v_group_stride = g.c_out_g * g.out_total;
  p.k_splits = pl.k_splits;
  p.k_blocks = pl.k_blocks;
  p.m_tiles = pl.m_tiles;
  p.n_tiles = pl.n_tiles;
  p.scale = 1.f;
}

template <typename F>
static int tc_dispatch_t(int dtype, F&& f) {
  switch (dtype) {
    case EINCONV_F32: return f((float*)nullptr);
    case EINCONV_BF16: return f((__nv_bfloat16*)nullptr);
    case EINCONV_F16: return f((__half*)nullptr);
    default: set_error("tensor-core kernels: unsupported dtype %d", dtype); return EINCONV_ERR_UNSUPPORTED;
  }
}

int tc_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype, const DevGeom& g, int,
                void*, int64_t, cudaStream_t stream) {
  const TcPlan pl = tc_plan(EINCONV_OP_CONV_FWD, g, dtype);
  Params p{};
  fill_params(p, g, pl);
  p.x = x; p.w = w; p.bias = bias; p.out = y;
  return tc_dispatch_t(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    if (pl.bn == 128) return tc_launch<T, kFwd, 128>(p, pl, g.groups, stream, "tcgen05_fwd");
    return tc_launch<T, kFwd, 64>(p, pl, g.groups, stream, "tcgen05_fwd");
  });
}

template <typename T>
static int tc_reduce(const void* partial, void* out, long long per_group, int splits, int groups, float scale,
                     cudaStream_t stream) {
  const long long total = per_group * groups;
  tc::tc_reduce_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const float*)partial, (T*)out,
                                                                                per_group, splits, groups, scale);
  EINCONV_CHECK_LAUNCH("tc_reduce_kernel");
  return EINCONV_OK;
}

// weight VJP: partial [G][splits][M = (ci,tap)][N = co]  ->  gw [G*co][(ci,tap)] (sum over splits, transposed)
template <typename T>
__global__ void __launch_bounds__(256) tc_reduce_wgrad_kernel(const float* __restrict__ partial,
                                                              T* __restrict__ gw, int M, int N, int k_splits,
                                                              int groups) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_group = (long long)M * N;
  if (idx >= per_group * groups) return;
  const int grp = (int)(idx / per_group);
  const long long rem = idx - (long long)grp * per_group;
  const int co = (int)(rem / M), r = (int)(rem - (long long)co * M);
  const float* src = partial + (long long)grp * k_splits * per_group + (long long)r * N + co;
  float s = 0.f;
  for (int k = 0; k < k_splits; ++k) s += src[(long long)k * per_group];
  Cvt<T>::store(gw + idx, s);
}

template <typename T>
static int tc_reduce_wgrad(const void* partial, void* gw, const DevGeom& g, int splits, cudaStream_t stream) {
  const int M = g.c_in_g * g.k_total, N = g.c_out_g;
  const long long total = (long long)M * N * g.groups;
  tc_reduce_wgrad_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const float*)partial, (T*)gw, M,
                                                                                  N, splits, g.groups);
  EINCONV_CHECK_LAUNCH("tc_reduce_wgrad_kernel");
  return EINCONV_OK;
}

int tc_conv_weight_vjp(const void* x, const void* v, void* gw, int dtype, const DevGeom& g, int, void* workspace,
                       int64_t workspace_bytes, cudaStream_t stream) {
  const TcPlan pl = tc_plan(EINCONV_OP_CONV_WEIGHT_VJP, g, dtype);
  const int64_t need = tc_workspace_bytes(EINCONV_OP_CONV_WEIGHT_VJP, g, dtype, 0);
  if (workspace_bytes < need || !workspace) {
    set_error("conv_weight_vjp: workspace of %lld bytes required, got %lld", (long long)need,
              (long long)workspace_bytes);
    return EINCONV_ERR_WORKSPACE;
  }
  Params p{};
  fill_params(p, g, pl);
  p.x = x; p.v = v; p.out = workspace;
  return tc_dispatch_t(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    int st = (pl.bn == 128) ? tc_launch<T, kWgrad, 128>(p, pl, g.groups, stream, "tcgen05_wgrad")
                            : tc_launch<T, kWgrad, 64>(p, pl, g.groups, stream, "tcgen05_wgrad");
    if (st) return st;
    // partial [G][splits][M=(ci,tap)][N=co]  ->  gw [G*co][ci,tap]: transpose while reducing
    return tc_reduce_wgrad<T>(workspace, gw, g, pl.k_splits, stream);
  });
}

int tc_kfc(const void* x, void* out, int dtype, const DevGeom& g, double scale, int, void* workspace,
           int64_t workspace_bytes, cudaStream_t stream) {
  const TcPlan pl = tc_plan(EINCONV_OP_KFC, g, dtype);
  const int64_t need = tc_workspace_bytes(EINCONV_OP_KFC, g, dtype, 0);
  if (workspace_bytes < need || !workspace) {
    set_error("kfc: workspace of %lld bytes required, got %lld", (long long)need, (long long)workspace_bytes);
    return EINCONV_ERR_WORKSPACE;
  }
  Params p{};
  fill_params(p, g, pl);
  p.x = x; p.out = workspace;
  return tc_dispatch_t(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    int st = (pl.bn == 128) ? tc_launch<T, kKfc, 128>(p, pl, g.groups, stream, "tcgen05_kfc")
                            : tc_launch<T, kKfc, 64>(p, pl, g.groups, stream, "tcgen05_kfc");
    if (st) return st;
    return tc_reduce<T>(workspace, out, pl.M * pl.N, pl.k_splits, g.groups, (float)scale, stream);
  });
}

}  // namespace einconv
