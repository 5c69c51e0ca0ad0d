// TMA throughput micro-benchmark: how fast can one CTA per SM pull {64 x 128 B} boxes through TMA,
// for a 2-d tensor map vs the 5-d maps tc_tma.cu uses, with small / large row strides?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void __launch_bounds__(128, 1) bw(const CUtensorMap* map, int iters, int n_rows_tiles, int depth, int n_w_tiles, int producers) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[32];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 32; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // producers > 0: one issuing thread in each of `producers` warps; producers < 0: -producers lanes of warp 0
  const bool lanes_mode = producers < 0;
  if (lanes_mode) producers = -producers;
  if (lanes_mode ? (int)threadIdx.x < producers : ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < producers)) {
    const int pw = lanes_mode ? threadIdx.x : threadIdx.x >> 5;
    iters /= producers;
    unsigned char* smem_p = smem + pw * depth * 8192;
    uint64_t* bar_p = bar + pw * 8;
    for (int i = 0; i < iters + depth; ++i) {
      if (i >= depth) {  // wait for load i - depth
        int s = (i - depth) % depth; uint32_t ph = ((i - depth) / depth) & 1, ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar_p[s])), "r"(ph) : "memory");
      }
      if (i < iters) {
        int s = i % depth;
        unsigned tile = ((blockIdx.x * 4 + pw) * 7919u + i * 104729u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_p[s])), "r"(8192) : "memory");
        if constexpr (RANK == 2) {
          int row = (tile % n_rows_tiles) * 64, col = ((tile / n_rows_tiles) % n_w_tiles) * 64;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(smem_p + s * 8192)), "l"((uint64_t)map), "r"(smem_u32(&bar_p[s])), "r"(col), "r"(row) : "memory");
        } else {
          int h = tile % n_rows_tiles, col = ((tile / n_rows_tiles) % n_w_tiles) * 64;
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(smem_u32(smem_p + s * 8192)), "l"((uint64_t)map), "r"(smem_u32(&bar_p[s])), "r"(col), "r"(h), "r"(0), "r"(0), "r"(0) : "memory");
        }
      }
    }
  }
}
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0, depth = argc > 2 ? atoi(argv[2]) : 4, ctas = argc > 3 ? atoi(argv[3]) : 148;
  void* fptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fptr, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fptr;
  size_t bytes = (size_t)(argc > 4 ? atoi(argv[4]) : 512) << 20; /* <= 64 MB stays L2-resident */ uint16_t* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
  alignas(64) CUtensorMap map; CUresult r; int n_rows_tiles, n_w_tiles = 1;
  if (variant == 0) {        // 2-d, rows of 128 B contiguous (pitch 128 B): box = contiguous 8 KB
    cuuint64_t dims[2] = {64, bytes / 128}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    n_rows_tiles = (int)(bytes / 128 / 64);
  } else if (variant == 1) { // 2-d, pitch 4096 B (typical GEMM operand, K = 2048 bf16)
    cuuint64_t dims[2] = {2048, bytes / 4096}; cuuint64_t st[1] = {4096}; cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    n_rows_tiles = (int)(bytes / 4096 / 64); n_w_tiles = 32;
  } else {                   // 5-d like tc_tma.cu: (W=128, H, D=1, C=64, B=1), box {64,1,1,64,1}; channel stride = H*256 B
    int H = (int)(bytes / (128 * 2 * 64));
    cuuint64_t dims[5] = {128, (cuuint64_t)H, 1, 64, 1}; cuuint64_t st[4] = {256, 256ull * H, 256ull * H, 256ull * H * 64};
    cuuint32_t box[5] = {64, 1, 1, 64, 1}, es[5] = {1, 1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    n_rows_tiles = H; n_w_tiles = 2;
  }
  CUtensorMap* gmap; cudaMalloc(&gmap, 128); cudaMemcpy(gmap, &map, 128, cudaMemcpyHostToDevice);
  int iters = 4000;
  int producers = argc > 5 ? atoi(argv[5]) : 1;
  const int smem_bytes = abs(producers) * depth * 8192;
  cudaFuncSetAttribute(bw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(bw<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (variant < 2) bw<2><<<ctas, 128, smem_bytes>>>(gmap, iters, n_rows_tiles, depth, n_w_tiles, producers); else bw<5><<<ctas, 128, smem_bytes>>>(gmap, iters, n_rows_tiles, depth, n_w_tiles, producers);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("producers %d MB %d variant %d depth %d ctas %d encode %d err %s: %.3f ms, %.1f GB/s total, %.1f cycles/box/SM @1.9GHz\n", producers, (int)(bytes >> 20), variant, depth, ctas, (int)r, cudaGetErrorString(cudaGetLastError()), ms, (double)ctas * iters * 8192 / ms / 1e6, ms * 1e-3 * 1.9e9 / iters);
  return 0;
}
