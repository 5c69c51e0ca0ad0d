// Micro-benchmark: the producer <-> MMA-issuer barrier ring of kfc_super_kernel WITHOUT any data movement.
// Warp 0 ("producer") waits for empty[s] and arrives on full[s]; warp 1 waits for full[s], issues `per_stage`
// tcgen05.mma (N = 128 or 256) and commits to empty[s].  Reports cycles per stage for several ring depths and
// wait flavours, against per_stage x (64 | 128) cycles of tensor work.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_pipe scripts/mma_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_spin(uint64_t* bar, uint32_t parity) { while (!try_wait(bar, parity)) {} }
// the kernels' flavour: clock64() watchdog inside the loop
__device__ __forceinline__ void wait_watchdog(uint64_t* bar, uint32_t parity) {
  if (try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
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

template <int N>
__global__ void __launch_bounds__(256, 1) pipe_kernel(int iters, int stages, int per_stage, int flavour, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[16], empty[16], done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (flavour & 16) break;  // no producer at all
      if (flavour & 1) wait_watchdog(&empty[s], ph ^ 1u); else wait_spin(&empty[s], ph ^ 1u);
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1 && (flavour & 64)) {
    // variant: ONE thread runs the whole issue loop (no elect, no warp sync); waits + commits as in flavour 0
    if (threadIdx.x == 32) {
      const uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem) + 128 * 128);
      constexpr uint32_t idesc = make_idesc(2, 128, N);
      long long t0 = clock64();
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) {
        if (!(flavour & 16)) wait_spin(&full[s], ph);
        for (int j = 0; j < per_stage; ++j)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(a + 2 * j), "l"(b + 2 * j), "r"(idesc), "r"((uint32_t)(i | j)) : "memory");
        if (!(flavour & 32))
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (i == iters - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      wait_spin(&done, 0);
      long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  } else if (warp == 1) {
    const uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem) + 128 * 128);
    constexpr uint32_t idesc = make_idesc(2, 128, N);
    long long t0 = clock64();
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (!(flavour & 16)) { if (flavour & 1) wait_watchdog(&full[s], ph); else wait_spin(&full[s], ph); }
      if (flavour & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        for (int j = 0; j < per_stage; ++j)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(a + 2 * j), "l"(b + 2 * j), "r"(idesc), "r"((uint32_t)(i | j)) : "memory");
        if (flavour & 8)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        else if (!(flavour & 32))
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (i == iters - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    wait_spin(&done, 0);
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  } else if (warp >= 4 && (flavour & 4)) {
    // the epilogue warps of the real kernel poll the accumulator barrier with nanosleep
    while (!try_wait(&done, 0)) __nanosleep(256);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N>
void run(int stages, int per_stage, int flavour) {
  long long* d;
  const int grid = 148;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = (128 + 256) * 128 + 1024;
  cudaFuncSetAttribute(pipe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  pipe_kernel<N><<<grid, 256, smem>>>(iters, stages, per_stage, flavour, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d stages=%2d mma/stage=%d flavour=%d (1 watchdog wait, 2 tcgen05 fence, 4 polling epilogue warps): %7.1f cycles per stage, tensor work %d (%s)\n",
         N, stages, per_stage, flavour, (double)mx / iters, per_stage * N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  // 8: plain arrive instead of tcgen05.commit; 16: no producer / no full wait (commits go nowhere); 32: no commit at all
  for (int flavour : {0, 16 + 32, 64, 64 + 16 + 32, 64 + 16}) {
    run<128>(12, 4, flavour);
    run<128>(12, 16, flavour);
    run<256>(12, 8, flavour);
  }
  return 0;
}
