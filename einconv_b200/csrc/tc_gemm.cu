// tcgen05 / TMEM tensor-core kernels (placeholder translation unit while the FFMA paths are
// brought up: reports "unsupported" so the dispatcher keeps to the CUDA-core kernels).
#include "dispatch.cuh"

namespace einconv {
bool tc_supported(int, const DevGeom&, int, int) { return false; }
const char* tc_kernel_name(int, const DevGeom&, int, int) { return "tcgen05_gather_gemm"; }
int64_t tc_workspace_bytes(int, const DevGeom&, int, int) { return 0; }
int tc_conv_fwd(const void*, const void*, const void*, void*, int, const DevGeom&, int, void*,
                int64_t, cudaStream_t) {
  set_error("tensor-core forward not built");
  return EINCONV_ERR_UNSUPPORTED;
}
int tc_conv_weight_vjp(const void*, const void*, void*, int, const DevGeom&, int, void*, int64_t,
                       cudaStream_t) {
  set_error("tensor-core weight VJP not built");
  return EINCONV_ERR_UNSUPPORTED;
}
int tc_kfc(const void*, void*, int, const DevGeom&, double, int, void*, int64_t, cudaStream_t) {
  set_error("tensor-core KFC not built");
  return EINCONV_ERR_UNSUPPORTED;
}
}  // namespace einconv
