// Bring-up probe: does a TMA tiled load work at all with the way tc_tma.cu issues it?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu   (run: ./tma_probe <variant>)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, bool PARAM>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, uint16_t* out, int bytes, int variant, int c0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* map = PARAM ? &pmap : gmap;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (variant != 9) {
      if constexpr (RANK == 2) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(&bar)), "r"(0), "r"(0) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(&bar)), "r"(c0), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
      }
    } else {
      // no TMA: complete the transaction by hand
      asm volatile("mbarrier.complete_tx.shared::cta.relaxed.cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    }
  }
  // everybody waits (bounded)
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (clock64() - t0 > 2000000000ll) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA\n"); return; }
  }
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  int c0 = argc > 2 ? atoi(argv[2]) : -1;
  int Dext = argc > 3 ? atoi(argv[3]) : 1;
  int Next = argc > 4 ? atoi(argv[4]) : 1;
  void* fptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fptr, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fptr;
  printf("variant %d encode fn %p q=%d\n", variant, fptr, (int)q);
  const int W = 64, H = 16, C = 8;
  std::vector<uint16_t> h((size_t)W * H * C * Dext * Next);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)i;
  uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 65536); cudaMemset(o, 0xff, 65536);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  alignas(64) CUtensorMap map;
  CUresult r;
  int bytes;
  bool rank5 = (variant == 1 || variant == 3 || variant == 5);
  CUtensorMapSwizzle swz = (variant >= 2 && variant <= 3) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (!rank5) {
    cuuint64_t dims[2] = {W, (cuuint64_t)H * C}; cuuint64_t strides[1] = {W * 2};
    cuuint32_t box[2] = {64, 8}; cuuint32_t es[2] = {1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 64 * 8 * 2;
  } else {
    cuuint64_t dims[5] = {W, H, (cuuint64_t)Dext, C, (cuuint64_t)Next}; cuuint64_t strides[4] = {W * 2, W * H * 2, (cuuint64_t)W * H * Dext * 2, (cuuint64_t)W * H * Dext * C * 2};
    cuuint32_t box[5] = {16, 4, 1, 8, 1}; cuuint32_t es[5] = {1, 1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 16 * 4 * 8 * 2;
  }
  printf("encode result %d\n", (int)r);
  CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(map)); cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
  bool param = variant != 4 && variant != 5;
  if (!rank5) { if (param) probe<2, true><<<1, 128, 16384>>>(map, gmap, o, bytes, variant, c0); else probe<2, false><<<1, 128, 16384>>>(map, gmap, o, bytes, variant, c0); }
  else { if (param) probe<5, true><<<1, 128, 16384>>>(map, gmap, o, bytes, variant, c0); else probe<5, false><<<1, 128, 16384>>>(map, gmap, o, bytes, variant, c0); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint16_t> res(bytes / 2); cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost);
    printf("first values:"); for (int i = 0; i < 24; ++i) printf(" %u", res[i]); printf("\n");
  }
  return 0;
}
