// Micro-benchmark: issue rate of tcgen05.mma (SS mode, K-major SW128 operands in shared memory, one accumulator
// or two alternating ones), with and without a tcgen05.commit every `per_commit` instructions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_rate scripts/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major A operand: `layout` 2 = SWIZZLE_128B (16-bit), 1 = SWIZZLE_128B_BASE32B (tf32: K groups of 4 rows)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((layout == 1 ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// AMN: 0 = K-major A, 1 = MN-major A (4 blocks of [32 K rows][128 B of M], as conv_rows.cu / conv_dense.cu stage it)
template <int FMT, int N, int AMN = 0>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int per_commit, int two_acc, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint64_t a = AMN ? make_desc_mn(smem_u32(smem), 4096, FMT == 2 ? 1 : 2) : make_desc(smem_u32(smem));
    const uint64_t b = make_desc(smem_u32(smem) + 128 * 128);
    constexpr uint32_t idesc = make_idesc(FMT, 128, N) | (AMN ? (1u << 15) : 0u);
    constexpr uint32_t a_step = AMN ? ((FMT == 2 ? 8 : 16) * 128) >> 4 : 2;  // one K step in 16-byte units
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
        const uint32_t d = tmem + ((two_acc && (i & 1)) ? 256u : 0u);
        const uint64_t ad = a + a_step * (i & (AMN && FMT != 2 ? 1 : 3)), bd = b + 2 * (i & 3);
        if (FMT == 2)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(i > 1)) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(i > 1)) : "memory");
        if (per_commit > 0 && (i % per_commit) == per_commit - 1 && i != iters - 1) {
          // commit to a barrier nobody waits on until the end (count arrivals instead of phases: re-init not needed for timing)
        }
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int FMT, int N, int AMN = 0>
void run(const char* name, int two_acc, int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = (128 + 256) * 128 + 1024;
  cudaFuncSetAttribute(rate_kernel<FMT, N, AMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  rate_kernel<FMT, N, AMN><<<grid, 128, smem>>>(iters, 0, two_acc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-10s N=%3d accumulators=%d grid=%3d: %7.1f cycles per MMA (%s)\n", name, N, two_acc ? 2 : 1, grid, (double)mx / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<2, 128>("tf32", 0, grid);
    run<2, 256>("tf32", 0, grid);
    run<2, 256>("tf32", 1, grid);
    run<0, 128>("f16", 0, grid);
    run<0, 256>("f16", 0, grid);
    run<0, 256>("f16", 1, grid);
    run<2, 64>("tf32", 0, grid);
    run<0, 64>("f16", 0, grid);
    run<2, 192>("tf32", 0, grid);
    run<2, 192, 1>("tf32 A-MN", 0, grid);
    run<2, 192, 1>("tf32 A-MN", 1, grid);
    run<2, 64, 1>("tf32 A-MN", 0, grid);
    run<2, 128, 1>("tf32 A-MN", 0, grid);
    run<2, 256, 1>("tf32 A-MN", 0, grid);
    run<0, 192, 1>("f16 A-MN", 0, grid);
    run<0, 64, 1>("f16 A-MN", 0, grid);
    run<0, 256, 1>("f16 A-MN", 0, grid);
  }
  return 0;
}
