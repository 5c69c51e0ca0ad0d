// 1x1 ("dense") convolutions as plain GEMMs on the channels-FIRST tensors: the specialisation the
// simplifier's dense rule selects (simplifications/opt.py:250-300: a pattern with K = 1, stride 1, no
// padding is an identity and disappears from the network, leaving  y[n,co,p] = sum_ci x[n,ci,p] w[co,ci]).
//
//   forward   : dst[n, co, p] = sum_ci src[n, ci, p] * w[co, ci]  (+ bias[co])     src = x, K = C_in
//   input VJP : dst[n, ci, p] = sum_co src[n, co, p] * w[co, ci]                    src = v, K = C_out
//
// No packed copy of anything.  With positions on M, a tile of the source is [K channels][128 positions]
// with the POSITIONS contiguous: an MN-major A operand, which TMA fetches straight from the tensor as boxes
// of {128 bytes of positions} x {64 channel rows} (3-d map [positions, channels, samples]: rows beyond the
// channel count zero-fill, so a K block never runs into the next sample).  The weights are read in place
// too: w[co][ci] is a K-major B operand for the forward pass (ci contiguous) and an MN-major one for the
// input VJP (K = co on the rows).  Accumulators [128 positions x BN channels] live in TMEM, double-buffered;
// the epilogue stores channels-first, coalesced along positions, exactly like conv_tma.cu's.
//
// Roles (320 threads, persistent CTAs): warp 0 TMA (one box per lane), warp 1 MMA issue, warps 2-9 epilogue
// (with K = 64..256 a tile has only 4..16 MMAs: the epilogue is the long pole, so it gets eight warps and a
// branch-free inner loop; the first version spent 334 instructions per 16-column chunk on one warp per
// scheduler and took 21 K cycles per tile).
// Needs 16-byte aligned rows: prod(spatial) * elem_size % 16 == 0 and C_in * elem_size % 16 == 0.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "dispatch.cuh"
#include "tc_ptx.cuh"

namespace einconv {
namespace dense {

using namespace ptx;

constexpr int kEpilogueWarps = 8;  // two per TMEM lane quarter, alternating 16-column chunks
constexpr int kThreads = 64 + 32 * kEpilogueWarps;
constexpr int kTileM = 128;
constexpr int kKB = 64;          // channel rows (K) per pipeline stage
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;

struct Params {
  int batch, hw, ck, cd;         // samples, positions per sample, reduction channels, destination channels
  int n_mtiles, n_ntiles, bn;    // tiles per sample along positions / along destination channels
  long long total_tiles;
  int n_kblocks;
  int a_blocks, a_block_bytes;   // A stage: a_blocks MN-major blocks of [kKB rows][128 bytes of positions]
  int pos_per_block;             // 64 (16-bit) / 32 (fp32)
  int b_mn;                      // B operand MN-major (input VJP) instead of K-major (forward)
  int b_boxes, b_box_bytes;      // B stage: K-major: kKB * es / 128 boxes of [bn rows][128 B of K];
                                 //          MN-major: bn / pos_per_block blocks of [kKB rows][128 B of N]
  int stage_bytes, n_stages;
  int mn_layout;                 // descriptor layout type of the MN-major operands (2, or 1 for 32-byte atoms)
  const void* bias;
  void* dst;
};

struct Bars {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  float bias[256];
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// MN-major, 128-byte swizzle (cute::UMMA canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO))): 128 bytes of
// M (or N) per K row, 8-row groups SBO = 1024 bytes apart, 128-byte blocks of M / N LBO apart.
// `layout`: 2 = SWIZZLE_128B (16-bit elements): 8-row groups, SBO = 1024 bytes.  1 = SWIZZLE_128B_BASE32B, the
// 128-byte swizzle with 32-byte atoms that MN-major operands of 32-bit elements (tf32) need (TMA:
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): its swizzle has two bits, so the K groups are FOUR rows and SBO = 512
// (measured with one-hot inputs, scripts/probe_dense_tf32.py: with SBO = 1024 the MMA took K rows 4..7 of a
// step from rows 8..11).
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t start16, uint32_t lbo16, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)(start16 & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)((layout == 1 ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr bool TF32 = true; static constexpr int FMT = 2; };
template <> struct Elem<__nv_bfloat16> { static constexpr bool TF32 = false; static constexpr int FMT = 1; };
template <> struct Elem<__half> { static constexpr bool TF32 = false; static constexpr int FMT = 0; };

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
conv_dense_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;
  Bars* bars = reinterpret_cast<Bars*>(ring + (size_t)p.n_stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ES = (int)sizeof(T);
  constexpr int KSTEP = 32 / ES;            // K per MMA: 16 (16-bit) / 8 (tf32)
  constexpr int KSTEPS = kKB / KSTEP;       // MMAs per stage and accumulator

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], kEpilogueWarps); }
    fence_barrier_init();
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.bn) tmem_cols <<= 1;
  if (warp == 1) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const long long tile_begin = p.total_tiles * blockIdx.x / gridDim.x;
  const long long tile_end = p.total_tiles * (blockIdx.x + 1) / gridDim.x;
  const int a_bytes = p.a_blocks * p.a_block_bytes;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one box per lane)
    int st = 0;
    uint32_t ph = 0;
    const uint32_t stage_tx = (uint32_t)(a_bytes + p.b_boxes * p.b_box_bytes);
    for (long long tile = tile_begin; tile < tile_end; ++tile) {
      const int ntile = (int)(tile % p.n_ntiles);
      const long long t2 = tile / p.n_ntiles;
      const int mtile = (int)(t2 % p.n_mtiles);
      const int n = (int)(t2 / p.n_mtiles);
      const int m0 = mtile * kTileM, n0 = ntile * p.bn;
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(&bars->empty[st], ph ^ 1);
        if (lane == 0) mbar_expect_tx(&bars->full[st], stage_tx);
        __syncwarp();
        unsigned char* stage = ring + (size_t)st * p.stage_bytes;
        const int k0 = kb * kKB;
        for (int bx = lane; bx < p.a_blocks + p.b_boxes; bx += 32) {
          if (bx < p.a_blocks) {
            tma_load_3d(stage + (size_t)bx * p.a_block_bytes, &map_a, &bars->full[st], m0 + bx * p.pos_per_block, k0, n);
          } else {
            const int j = bx - p.a_blocks;
            unsigned char* bt = stage + a_bytes + (size_t)j * p.b_box_bytes;
            if (p.b_mn) tma_load_2d(bt, &map_b, &bars->full[st], n0 + j * p.pos_per_block, k0);  // [K rows][128 B of N]
            else tma_load_2d(bt, &map_b, &bars->full[st], k0 + j * p.pos_per_block, n0);        // [N rows][128 B of K]
          }
        }
        __syncwarp();
        if (++st == p.n_stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(Elem<T>::FMT, kTileM, p.bn) | (1u << 15) | (p.b_mn ? (1u << 16) : 0u);
    const uint32_t ring16 = smem_u32(ring) >> 4, stage16 = (uint32_t)p.stage_bytes >> 4;
    const uint32_t a_lbo16 = (uint32_t)p.a_block_bytes >> 4, b_box16 = (uint32_t)p.b_box_bytes >> 4;
    const uint64_t kdesc = make_desc_kmajor(0, 128);
    constexpr uint32_t kMnStep16 = (uint32_t)(KSTEP * 128) >> 4;  // K rows of one MMA in an MN-major tile
    constexpr int kStepsPerBox = 128 / 32;                        // K-major: 32 bytes of K per MMA
    const uint64_t kMnLayout = (uint64_t)p.mn_layout;
    int st = 0;
    uint32_t ph = 0, it = 0;
    for (long long tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.bn;
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(&bars->full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a16 = ring16 + (uint32_t)st * stage16;
          const uint32_t b16 = a16 + ((uint32_t)a_bytes >> 4);
          const uint64_t a_base = make_desc_mnmajor(a16, a_lbo16, kMnLayout);
          for (int k = 0; k < KSTEPS; ++k) {
            uint64_t b_desc;
            if (p.b_mn) b_desc = make_desc_mnmajor(b16 + (uint32_t)k * kMnStep16, b_box16, kMnLayout);
            else b_desc = kdesc + (uint64_t)(b16 + (uint32_t)(k / kStepsPerBox) * b_box16 + 2u * (uint32_t)(k % kStepsPerBox));
            umma<Elem<T>::TF32>(d_tmem, a_base + (uint64_t)((uint32_t)k * kMnStep16), b_desc, idesc, (uint32_t)(kb | k));
          }
          umma_commit(&bars->empty[st]);
          if (kb + 1 == p.n_kblocks) umma_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
        if (++st == p.n_stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;          // TMEM lanes 32 * quarter .. + 31 are visible to this warp
    const int half = (warp - 2) >> 2;      // chunks half, half + 2, ..
    const int etid = threadIdx.x - 64;
    T* dst = reinterpret_cast<T*>(p.dst);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    const long long hw = p.hw;
    const int n_chunks = p.bn >> 4;
    uint32_t it = 0;
    int cur_ntile = -1;
    for (long long tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int ntile = (int)(tile % p.n_ntiles);
      const long long t2 = tile / p.n_ntiles;
      const int mtile = (int)(t2 % p.n_mtiles);
      const int n = (int)(t2 / p.n_mtiles);
      const int n0 = ntile * p.bn;
      if (bias && ntile != cur_ntile) {
        cur_ntile = ntile;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
        for (int c = etid; c < p.bn; c += 32 * kEpilogueWarps)
          bars->bias[c] = n0 + c < p.cd ? (float)Cvt<T>::load(bias + n0 + c) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
      }
      const int pos = mtile * kTileM + quarter * 32 + lane;
      const bool valid = pos < p.hw;
      T* dst_row = dst + ((long long)n * p.cd + n0) * hw + pos;
      mbar_wait_relaxed(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * (uint32_t)p.bn + ((uint32_t)(quarter * 32) << 16);
      for (int ch = half; ch < n_chunks; ch += 2) {
        const int c0 = ch << 4;
        uint32_t r[16];
        tmem_ld16_nowait(taddr + c0, r);
        tmem_ld_wait();
        if (ch + 2 >= n_chunks) {
          // this warp's last accumulator read of the tile: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        }
        if (valid) {
          T* d0 = dst_row + (long long)c0 * hw;
          if (bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + bars->bias[c0 + j]);
          }
          if (n0 + c0 + 16 <= p.cd) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              Cvt<T>::store(d0, __uint_as_float(r[j]));
              d0 += hw;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n0 + c0 + j < p.cd) Cvt<T>::store(d0, __uint_as_float(r[j]));
              d0 += hw;
            }
          }
        }
      }
      if (half >= n_chunks) {
        // (bn = 16: the second warp of the quarter has no chunk but still owes its arrival)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
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
                  (mn_major && dtype == EINCONV_F32 && !getenv("EINCONV_DENSE_TF32_SW128")) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                                                          : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (dense conv, rank %d) failed with CUresult %d", rank, (int)r);
    return EINCONV_ERR_CUDA;
  }
  return EINCONV_OK;
}

bool supported(int op, const DevGeom& g, int dtype, int math) {
  if (op != EINCONV_OP_CONV_FWD && op != EINCONV_OP_CONV_INPUT_VJP) return false;
  if (getenv("EINCONV_NO_CONV_DENSE")) return false;
  if (dtype == EINCONV_F64) return false;
  if (dtype == EINCONV_F32 && (math != EINCONV_MATH_TF32 || getenv("EINCONV_CONV_DENSE_NO_TF32"))) return false;
  if (g.groups != 1 || g.batch == 0) return false;
  for (int d = 0; d < g.n_dims; ++d)
    if (g.k[d] != 1 || g.stride[d] != 1 || g.pad[d] != 0 || g.in[d] != g.out[d]) return false;
  const long long es = elem_size(dtype);
  if (((long long)g.in_total * es) % 16 != 0 || ((long long)g.c_in_g * es) % 16 != 0) return false;
  if (g.c_in_g == 0 || g.c_out_g == 0 || g.in_total == 0) return false;
  return true;
}

template <typename T>
static int run_t(int op, const void* src, const void* w, const void* bias, void* dst, int dtype, const DevGeom& g,
                 cudaStream_t stream) {
  const int es = (int)sizeof(T);
  const bool fwd = op == EINCONV_OP_CONV_FWD;
  Params p{};
  p.batch = g.batch;
  p.hw = g.in_total;
  p.ck = fwd ? g.c_in_g : g.c_out_g;
  p.cd = fwd ? g.c_out_g : g.c_in_g;
  p.pos_per_block = 128 / es;
  p.b_mn = fwd ? 0 : 1;
  p.n_ntiles = (p.cd + 255) / 256;
  const int gran = p.b_mn ? p.pos_per_block : 16;  // MN-major B: whole 128-byte blocks of N
  p.bn = ((p.cd + p.n_ntiles - 1) / p.n_ntiles + gran - 1) / gran * gran;
  p.n_mtiles = (p.hw + kTileM - 1) / kTileM;
  p.total_tiles = (long long)p.batch * p.n_mtiles * p.n_ntiles;
  p.n_kblocks = (p.ck + kKB - 1) / kKB;
  p.a_blocks = kTileM / p.pos_per_block;
  p.a_block_bytes = kKB * 128;
  if (p.b_mn) {
    p.b_boxes = p.bn / p.pos_per_block;
    p.b_box_bytes = kKB * 128;
  } else {
    p.b_boxes = kKB * es / 128;
    p.b_box_bytes = (p.bn * 128 + 1023) / 1024 * 1024;
  }
  p.stage_bytes = p.a_blocks * p.a_block_bytes + p.b_boxes * p.b_box_bytes;
  p.n_stages = std::min(kMaxStages, kSmemBudget / p.stage_bytes);
  EINCONV_CHECK_ARG(p.n_stages >= 2, "conv_dense: stage does not fit twice into shared memory");
  p.mn_layout = (es == 4 && !getenv("EINCONV_DENSE_TF32_SW128")) ? 1 : 2;
  if (const char* e = getenv("EINCONV_DENSE_MN_LAYOUT")) p.mn_layout = atoi(e);
  p.bias = fwd ? bias : nullptr;
  p.dst = dst;

  CUtensorMap map_a, map_b;
  {
    // source [samples][channels][positions], positions innermost
    cuuint64_t dims[3] = {(cuuint64_t)p.hw, (cuuint64_t)p.ck, (cuuint64_t)p.batch};
    cuuint64_t strides[2] = {(cuuint64_t)p.hw * es, (cuuint64_t)p.hw * p.ck * es};
    cuuint32_t box[3] = {(cuuint32_t)p.pos_per_block, (cuuint32_t)kKB, 1};
    int st = encode(&map_a, dtype, src, 3, dims, strides, box, true);
    if (st) return st;
  }
  {
    // weights [C_out][C_in]: rows = co, ci innermost
    cuuint64_t dims[2] = {(cuuint64_t)g.c_in_g, (cuuint64_t)g.c_out_g};
    cuuint64_t strides[1] = {(cuuint64_t)g.c_in_g * es};
    cuuint32_t box[2];
    if (p.b_mn) { box[0] = (cuuint32_t)p.pos_per_block; box[1] = (cuuint32_t)kKB; }  // 128 B of ci (N) x K rows of co
    else { box[0] = (cuuint32_t)p.pos_per_block; box[1] = (cuuint32_t)p.bn; }         // 128 B of ci (K) x N rows of co
    int st = encode(&map_b, dtype, w, 2, dims, strides, box, p.b_mn != 0);
    if (st) return st;
  }
  const size_t smem = (size_t)p.n_stages * p.stage_bytes + sizeof(Bars) + 64;
  auto kern = conv_dense_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_dense)");
  const unsigned grid = (unsigned)std::min<long long>(p.total_tiles, num_sms());
  kern<<<grid, kThreads, smem, stream>>>(map_a, map_b, p);
  EINCONV_CHECK_LAUNCH("conv_dense_kernel");
  return EINCONV_OK;
}

int run(int op, const void* src, const void* w, const void* bias, void* dst, int dtype, const DevGeom& g,
        cudaStream_t stream) {
  switch (dtype) {
    case EINCONV_F32: return run_t<float>(op, src, w, bias, dst, dtype, g, stream);
    case EINCONV_BF16: return run_t<__nv_bfloat16>(op, src, w, bias, dst, dtype, g, stream);
    case EINCONV_F16: return run_t<__half>(op, src, w, bias, dst, dtype, g, stream);
    default: set_error("conv_dense: unsupported dtype"); return EINCONV_ERR_UNSUPPORTED;
  }
}

}  // namespace dense
}  // namespace einconv
