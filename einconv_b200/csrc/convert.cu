// fp32 -> scaled fp16 operand copy for the factor networks in TF32 mode.
//
// TF32 keeps 10 explicit mantissa bits of an fp32 operand; so does fp16.  Scaling a tensor by a power
// of two s such that max|x| * s lies in [2^13, 2^14) and rounding to fp16 (round-to-nearest-even;
// the tensor core TRUNCATES fp32 operands to TF32) therefore loses no more precision than kind::tf32
// for every element within 2^-27 of the maximum, and smaller elements contribute < 2^-54 max^2 to a
// Gram sum.  In exchange kind::f16 runs at twice the kind::tf32 rate on half the operand bytes, which
// is what the L2-bound KFC kernel needs.  The factor is un-scaled by 1 / s^2 (exact) in the finish
// kernel, read from device memory so that no host synchronisation is needed.
#include <cuda_fp16.h>

#include <algorithm>

#include "dispatch.cuh"

namespace einconv {

__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n,
                                                     unsigned* __restrict__ out_bits) {
  float m = 0.f;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const long long step = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread (one per iteration was latency-bound: 2 TB/s)
  for (; i + 3 * step < n4; i += 4 * step) {
    const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + step), c = __ldg(x4 + i + 2 * step), d = __ldg(x4 + i + 3 * step);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
  }
  for (; i < n4; i += step) {
    const float4 v = __ldg(x4 + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (long long j = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += step)
    m = fmaxf(m, fabsf(__ldg(x + j)));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  // non-negative floats (and +inf / NaN payloads) order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

__global__ void __launch_bounds__(256) scale_to_half_kernel(const float* __restrict__ x, __half* __restrict__ y,
                                                            long long n, const unsigned* __restrict__ absmax_bits,
                                                            float* __restrict__ inv_s2) {
  const float s = operand_scale(*absmax_bits);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_s2 = 1.f / (s * s);
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  uint2* y4 = reinterpret_cast<uint2*>(y);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    const __half2 lo = __floats2half2_rn(v.x * s, v.y * s), hi = __floats2half2_rn(v.z * s, v.w * s);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned*>(&lo);
    o.y = *reinterpret_cast<const unsigned*>(&hi);
    y4[i] = o;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(__ldg(x + i) * s);
}

// two tensors in one launch: blocks [0, blocks_a) reduce a into out_bits[0], the others b into out_bits[1]
__global__ void __launch_bounds__(256) absmax2_kernel(const float* __restrict__ a, long long na, unsigned blocks_a,
                                                      const float* __restrict__ b, long long nb,
                                                      unsigned* __restrict__ out_bits) {
  const bool first = blockIdx.x < blocks_a;
  const float* x = first ? a : b;
  const long long n = first ? na : nb;
  const long long bid = first ? blockIdx.x : blockIdx.x - blocks_a;
  const long long nblk = first ? blocks_a : gridDim.x - blocks_a;
  float m = 0.f;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = bid * blockDim.x + threadIdx.x; i < n4; i += nblk * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (long long i = n4 * 4 + bid * blockDim.x + threadIdx.x; i < n; i += nblk * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits + (first ? 0 : 1), __float_as_uint(m));
}

int launch_absmax2(const float* a, long long na, const float* b, long long nb, unsigned* out_bits,
                   cudaStream_t stream) {
  EINCONV_CHECK_ARG(((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0, "absmax: misaligned tensor");
  const unsigned blocks_a = (unsigned)std::min<long long>((na / 4 + 255) / 256 + 1, 8ll * num_sms());
  const unsigned blocks_b = (unsigned)std::min<long long>((nb / 4 + 255) / 256 + 1, 2ll * num_sms());
  absmax2_kernel<<<blocks_a + blocks_b, 256, 0, stream>>>(a, na, blocks_a, b, nb, out_bits);
  EINCONV_CHECK_LAUNCH("absmax2_kernel");
  return EINCONV_OK;
}

// max |x| of an fp32 tensor as float bits, atomically maxed into *out_bits (caller zeroes it first)
int launch_absmax(const float* x, long long n, unsigned* out_bits, cudaStream_t stream) {
  if (n == 0) return EINCONV_OK;
  EINCONV_CHECK_ARG(((uintptr_t)x % 16) == 0, "absmax: misaligned tensor");
  const unsigned blocks = (unsigned)std::min<long long>((n / 4 + 255) / 256 + 1, 8ll * num_sms());
  absmax_kernel<<<blocks, 256, 0, stream>>>(x, n, out_bits);
  EINCONV_CHECK_LAUNCH("absmax_kernel");
  return EINCONV_OK;
}

// header: [0] absmax bits, [1] 1 / s^2 (float); x must be 16-byte aligned (torch allocations are)
int scaled_half_copy(const float* x, __half* y, long long n, void* header, cudaStream_t stream) {
  EINCONV_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 8) == 0, "scaled_half_copy: misaligned tensor");
  cudaError_t e = cudaMemsetAsync(header, 0, 8, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(absmax)");
  if (n == 0) return EINCONV_OK;
  const unsigned blocks = (unsigned)std::min<long long>((n / 4 + 255) / 256 + 1, 8ll * num_sms());
  absmax_kernel<<<blocks, 256, 0, stream>>>(x, n, (unsigned*)header);
  EINCONV_CHECK_LAUNCH("absmax_kernel");
  scale_to_half_kernel<<<blocks, 256, 0, stream>>>(x, y, n, (const unsigned*)header, (float*)header + 1);
  EINCONV_CHECK_LAUNCH("scale_to_half_kernel");
  return EINCONV_OK;
}

// first half of scaled_half_copy: header[0] = absmax bits; the consumer (flat_copy_kernel<.., CONVERT>) scales,
// converts and writes header[1] = 1 / s^2 itself
int absmax_header(const float* x, long long n, void* header, cudaStream_t stream) {
  EINCONV_CHECK_ARG(((uintptr_t)x % 16) == 0, "absmax_header: misaligned tensor");
  cudaError_t e = cudaMemsetAsync(header, 0, 8, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(absmax)");
  if (n == 0) return EINCONV_OK;
  const unsigned blocks = (unsigned)std::min<long long>((n / 4 + 255) / 256 + 1, 8ll * num_sms());
  absmax_kernel<<<blocks, 256, 0, stream>>>(x, n, (unsigned*)header);
  EINCONV_CHECK_LAUNCH("absmax_kernel");
  return EINCONV_OK;
}

}  // namespace einconv
