// Shared device/host helpers for the einconv_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/einconv_b200.h"

namespace einconv {

constexpr int kMaxDims = EINCONV_MAX_DIMS;
// SMs of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on B200): grid sizing and
// the split-K heuristics count in units of it
int num_sms();

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local message, C status codes)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
void count_launch(int n = 1);

#define EINCONV_CHECK_ARG(cond, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ::einconv::set_error(__VA_ARGS__);    \
      return EINCONV_ERR_INVALID;           \
    }                                       \
  } while (0)

#define EINCONV_CHECK_LAUNCH(what)                                   \
  do {                                                               \
    cudaError_t err__ = cudaGetLastError();                          \
    if (err__ != cudaSuccess) return ::einconv::cuda_fail(err__, what); \
    ::einconv::count_launch();                                       \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Division by a runtime constant: q = (n * mul) >> 32 >> shift  (n < 2^31)
// ---------------------------------------------------------------------------------------------
struct FastDiv {
  uint32_t div = 1, mul = 0, shift = 0;
  FastDiv() = default;
  explicit FastDiv(uint32_t d) : div(d) {
    if (d <= 1) {
      div = 1;
      mul = 0;
      shift = 0;
      return;
    }
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;  // ceil(log2 d)
    shift = s - 1;
    // mul = ceil(2^(32+shift) / d); fits in 32 bits for d >= 2 with shift = ceil(log2 d) - 1
    uint64_t one = 1ull << (32 + shift);
    mul = (uint32_t)((one + d - 1) / d);
  }
  __host__ __device__ __forceinline__ uint32_t quot(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return div == 1 ? n : (__umulhi(n, mul) >> shift);
#else
    return div == 1 ? n : (uint32_t)(((uint64_t)n * mul) >> 32 >> shift);
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = quot(n);
    r = n - q * div;
  }
};

// ---------------------------------------------------------------------------------------------
// Device-side geometry.  Spatial dimension d has
//   in[d], k[d], stride[d], dil[d], pad[d] (left), out[d]
// and the index law  i = stride * o + dil * k - pad   (conv_index_pattern.py:131-134).
// ---------------------------------------------------------------------------------------------
struct DevGeom {
  int n_dims;
  int groups;
  int batch;
  int c_in_g;
  int c_out_g;
  int in[kMaxDims];
  int k[kMaxDims];
  int stride[kMaxDims];
  int dil[kMaxDims];
  int pad[kMaxDims];
  int out[kMaxDims];
  int in_total;   // prod(in)
  int out_total;  // prod(out)
  int k_total;    // prod(k)
  FastDiv out_div[kMaxDims];
  FastDiv k_div[kMaxDims];
  FastDiv in_div[kMaxDims];
  FastDiv stride_div[kMaxDims];
  // 0: `in` are the convolution's inputs (window sums ask "is input i under tap k of some output").
  // 1: transpose-convolution view (swap_io): `in` are the positions o of the tensor being summed and
  //    a position belongs to tap k iff the convolution input index s*o + dil*k - pad lies in [0, out).
  int member_mode;
};

// Geometry of the associated convolution seen from the transpose convolution: in <-> out.
inline DevGeom swap_io(const DevGeom& g) {
  DevGeom t = g;
  for (int d = 0; d < kMaxDims; ++d) {
    t.in[d] = g.out[d];
    t.out[d] = g.in[d];
    t.in_div[d] = g.out_div[d];
    t.out_div[d] = g.in_div[d];
  }
  t.in_total = g.out_total;
  t.out_total = g.in_total;
  t.member_mode = 1;
  return t;
}

// Validates and converts the ABI struct.  Returns 0 or a negative status (message set).
int make_dev_geom(const einconv_geom* g, DevGeom* out, bool need_c_out);

inline int64_t elem_size(int dtype) {
  switch (dtype) {
    case EINCONV_F32: return 4;
    case EINCONV_BF16: return 2;
    case EINCONV_F16: return 2;
    case EINCONV_F64: return 8;
    default: return 0;
  }
}

// ---------------------------------------------------------------------------------------------
// dtype <-> compute type conversion
// ---------------------------------------------------------------------------------------------
template <typename T> struct Cvt;
template <> struct Cvt<float> {
  using acc_t = float;
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <> struct Cvt<double> {
  using acc_t = double;
  static __device__ __forceinline__ double load(const double* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(double* p, double v) { *p = v; }
};
template <> struct Cvt<__nv_bfloat16> {
  using acc_t = float;
  // NB: read the bits through the builtin __ldg(unsigned short*).  The bf16/half __ldg overloads
  // are non-volatile inline asm, which the compiler may hoist above the bounds predicate of a
  // `ok ? load(p) : 0` select and fault on padding addresses.
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) {
    const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
    return __uint_as_float(static_cast<unsigned int>(u) << 16);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
  }
};
template <> struct Cvt<__half> {
  using acc_t = float;
  static __device__ __forceinline__ float load(const __half* p) {
    const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
    return __half2float(__ushort_as_half(u));
  }
  static __device__ __forceinline__ void store(__half* p, float v) { *p = __float2half_rn(v); }
};

template <typename F>
inline int dispatch_dtype(int dtype, F&& f) {
  switch (dtype) {
    case EINCONV_F32: return f((float*)nullptr);
    case EINCONV_BF16: return f((__nv_bfloat16*)nullptr);
    case EINCONV_F16: return f((__half*)nullptr);
    case EINCONV_F64: return f((double*)nullptr);
    default:
      set_error("unknown dtype %d", dtype);
      return EINCONV_ERR_INVALID;
  }
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Power-of-two operand scale for the scaled-fp16 copies of fp32 tensors (convert.cu): max|x| * s lies in
// [2^13, 2^14), far below the fp16 maximum; 1 for empty / all-zero / non-finite tensors.
__device__ __forceinline__ float operand_scale(unsigned absmax_bits) {
  const float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
  const int k = max(-60, min(60, 14 - e));
  return ldexpf(1.f, k);
}

}  // namespace einconv
