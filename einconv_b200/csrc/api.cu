// extern "C" boundary of libeinconv_b200.so: argument validation, kernel selection, workspace
// accounting.  See include/einconv_b200.h for the contract of each entry point and the reference
// call site it replaces.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include <cstdlib>

#include "dispatch.cuh"

namespace einconv {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(err));
  return EINCONV_ERR_CUDA;
}

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int make_dev_geom(const einconv_geom* g, DevGeom* out, bool need_c_out) {
  EINCONV_CHECK_ARG(g != nullptr, "geometry is NULL");
  EINCONV_CHECK_ARG(g->n_dims >= 1 && g->n_dims <= kMaxDims, "n_dims must be in [1, %d], got %d",
                    kMaxDims, g->n_dims);
  EINCONV_CHECK_ARG(g->groups >= 1, "groups must be >= 1, got %d", g->groups);
  EINCONV_CHECK_ARG(g->batch >= 0 && g->batch < (1ll << 31), "bad batch size %lld", (long long)g->batch);
  EINCONV_CHECK_ARG(g->c_in_g >= 0 && g->c_in_g < (1ll << 31), "bad c_in_g %lld", (long long)g->c_in_g);
  if (need_c_out)
    EINCONV_CHECK_ARG(g->c_out_g >= 0 && g->c_out_g < (1ll << 31), "bad c_out_g %lld", (long long)g->c_out_g);
  DevGeom d{};
  d.n_dims = g->n_dims;
  d.groups = g->groups;
  d.batch = (int)g->batch;
  d.c_in_g = (int)g->c_in_g;
  d.c_out_g = need_c_out ? (int)g->c_out_g : 0;
  long long in_total = 1, out_total = 1, k_total = 1;
  for (int i = 0; i < g->n_dims; ++i) {
    EINCONV_CHECK_ARG(g->in[i] >= 0 && g->k[i] >= 1 && g->stride[i] >= 1 && g->dilation[i] >= 1 &&
                          g->pad_left[i] >= 0 && g->out[i] >= 0,
                      "invalid geometry in dimension %d (in=%d k=%d stride=%d dilation=%d pad=%d out=%d)",
                      i, g->in[i], g->k[i], g->stride[i], g->dilation[i], g->pad_left[i], g->out[i]);
    // every output must be able to reach its last tap without running past what padding allows;
    // the right padding is implied by `out`, so only the left edge can be checked here.
    d.in[i] = g->in[i];
    d.k[i] = g->k[i];
    d.stride[i] = g->stride[i];
    d.dil[i] = g->dilation[i];
    d.pad[i] = g->pad_left[i];
    d.out[i] = g->out[i];
    d.out_div[i] = FastDiv((uint32_t)g->out[i]);
    d.k_div[i] = FastDiv((uint32_t)g->k[i]);
    d.in_div[i] = FastDiv((uint32_t)g->in[i]);
    d.stride_div[i] = FastDiv((uint32_t)g->stride[i]);
    in_total *= g->in[i];
    out_total *= g->out[i];
    k_total *= g->k[i];
    EINCONV_CHECK_ARG(in_total < (1ll << 31) && out_total < (1ll << 31) && k_total < (1ll << 24),
                      "spatial extent too large");
  }
  for (int i = g->n_dims; i < kMaxDims; ++i) {
    d.in[i] = 1; d.k[i] = 1; d.stride[i] = 1; d.dil[i] = 1; d.pad[i] = 0; d.out[i] = 1;
  }
  d.in_total = (int)in_total;
  d.out_total = (int)out_total;
  d.k_total = (int)k_total;
  *out = d;
  return EINCONV_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel selection
// ---------------------------------------------------------------------------------------------
enum class Path { FFMA, DEPTHWISE, TENSOR, TMA, CONV_TMA, WGRAD_SLAB, DENSE, CONV_ROWS };

static int resolve_math(int dtype, int math) {
  if (math == EINCONV_MATH_AUTO)
    return (dtype == EINCONV_BF16 || dtype == EINCONV_F16) ? EINCONV_MATH_HALF : EINCONV_MATH_FP32;
  return math;
}

// `aligned`: every tensor pointer of the call is 16-byte aligned.  The tensor-core routes move data with
// TMA boxes and 16-byte vector accesses; a misaligned base (e.g. the batch-shard view x_all[1:] of a
// tensor with an odd per-sample element count) takes the CUDA-core path instead of failing.
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static Path select_path(int op, const DevGeom& g, int dtype, int math, int kernels, bool aligned = true) {
  if (kernels == EINCONV_KERNELS_GENERIC) return Path::FFMA;
  if ((op == EINCONV_OP_CONV_FWD || op == EINCONV_OP_CONV_INPUT_VJP) && depthwise_supported(g, dtype))
    return Path::DEPTHWISE;
  const int m = resolve_math(dtype, math);
  if (aligned && (m == EINCONV_MATH_TF32 || m == EINCONV_MATH_HALF)) {
    // reduction networks: TMA-fed tcgen05 kernel when the layout allows it (16-byte aligned rows);
    // the software-gather tensor-core kernel is kept for the forward pass only (for the reduction
    // networks it is slower than the CUDA-core kernel)
    // 1x1 kernels on fp32 tensors: the TMA reduction kernel reads the channels-first tensors as they are
    // (K-major: positions contiguous), where the slab route would write two scaled fp16 packed copies first
    // (measured 41 vs 102 us on [32,512,28,28] -> 128, 95 vs 129 us on [32,256,56,56] -> 64)
    if (op == EINCONV_OP_CONV_WEIGHT_VJP && dtype == EINCONV_F32 && g.k_total == 1 && tma::tma_supported(op, g, dtype))
      return Path::TMA;
    if (op == EINCONV_OP_CONV_WEIGHT_VJP && wslab::supported(g, dtype)) return Path::WGRAD_SLAB;
    if (op == EINCONV_OP_CONV_WEIGHT_VJP || op == EINCONV_OP_KFC)
      return tma::tma_supported(op, g, dtype) ? Path::TMA : Path::FFMA;
    // 1x1 / stride 1 / no padding (the simplifier's dense rule): a GEMM on the channels-first tensors themselves
    if ((op == EINCONV_OP_CONV_FWD || op == EINCONV_OP_CONV_INPUT_VJP) && dense::supported(op, g, dtype, m))
      return Path::DENSE;
    // stride-1 2-d forward on fp32 tensors: MN-major operand from the channels-first tensor, kw taps in the epilogue
    if ((op == EINCONV_OP_CONV_FWD || op == EINCONV_OP_CONV_INPUT_VJP) && crows::supported(op, g, dtype, m))
      return Path::CONV_ROWS;
    // forward / input VJP: TMA-fed implicit GEMM over a zero-padded channels-last copy (stride 1)
    if ((op == EINCONV_OP_CONV_FWD || op == EINCONV_OP_CONV_INPUT_VJP) && ctma::supported(op, g, dtype, m))
      return Path::CONV_TMA;
    if (tc_supported(op, g, dtype, m)) return Path::TENSOR;
  }
  return Path::FFMA;
}

// KFAC-reduce on tensor cores: stage 1 (window sums, HBM-bound) writes u as a channels-first 1-d
// "image" [G*A channels][B positions] in fp32; stage 2 (sum_b u u^T) is then exactly the KFC network
// of a 1x1 convolution over that image and runs on the TMA-fed tcgen05 kernel (TF32).
static bool kfac_stage2_geom(const DevGeom& g, DevGeom* g2) {
  const long long A = (long long)g.c_in_g * g.k_total;
  if (A >= (1ll << 24) || g.batch < 1) return false;
  einconv_geom e{};
  e.n_dims = 1; e.groups = g.groups; e.batch = 1; e.c_in_g = A; e.c_out_g = 0;
  e.in[0] = g.batch; e.k[0] = 1; e.stride[0] = 1; e.dilation[0] = 1; e.pad_left[0] = 0; e.out[0] = g.batch;
  return make_dev_geom(&e, g2, false) == EINCONV_OK;
}

static bool kfac_reduce_tensor(const DevGeom& g, int dtype, int math, int kernels, DevGeom* g2) {
  if (kernels == EINCONV_KERNELS_GENERIC || dtype == EINCONV_F64) return false;
  const int m = resolve_math(dtype, math);
  if (m != EINCONV_MATH_TF32 && m != EINCONV_MATH_HALF) return false;
  if ((long long)g.c_in_g * g.k_total < 256) return false;  // tiny factors: the CUDA-core GEMM is as fast
  return kfac_stage2_geom(g, g2) && tma::tma_supported(EINCONV_OP_KFC, *g2, EINCONV_F32);
}

// KFC of a layer the TMA kernel cannot read directly (e.g. C_in = 3: channel blocks below 32): write
// the unfolded input once (bit-exact gather) and run the factor as the KFC of a 1x1 convolution over
// the "image" [B, G*A channels, prod(out) positions].  Costs one extra pass over U in HBM but moves
// the A x A x (B*O) contraction from CUDA cores to the tensor cores (ResNet-50 stem: 14.4 -> ~3 ms).
static bool kfc_via_unfold(const DevGeom& g, int dtype, int math, int kernels, DevGeom* g1, int64_t* u_bytes) {
  if (kernels == EINCONV_KERNELS_GENERIC || dtype == EINCONV_F64) return false;
  const int m = resolve_math(dtype, math);
  if (m != EINCONV_MATH_TF32 && m != EINCONV_MATH_HALF) return false;
  if (tma::tma_supported(EINCONV_OP_KFC, g, dtype)) return false;
  const long long A = (long long)g.c_in_g * g.k_total;
  if (A < 32 || A >= (1ll << 20) || g.k_total == 1) return false;
  const long long bytes = (long long)g.batch * g.groups * A * g.out_total * elem_size(dtype);
  if (bytes > (8ll << 30)) return false;
  einconv_geom e{};
  e.n_dims = 1; e.groups = g.groups; e.batch = g.batch; e.c_in_g = A; e.c_out_g = 0;
  e.in[0] = g.out_total; e.k[0] = 1; e.stride[0] = 1; e.dilation[0] = 1; e.pad_left[0] = 0; e.out[0] = g.out_total;
  if (make_dev_geom(&e, g1, false) != EINCONV_OK) return false;
  if (!tma::tma_supported(EINCONV_OP_KFC, *g1, dtype)) return false;
  *u_bytes = ((bytes + 1023) / 1024) * 1024;
  return true;
}

// KFC of fp32 tensors in TF32 mode runs on scaled fp16 operand copies (same 10-bit mantissa, see
// convert.cu) whenever a tensor-core path exists for the fp16 problem.  EINCONV_KFC_TF32=1 keeps
// kind::tf32 on the fp32 data.
static bool kfc_half_route(const DevGeom& g, int dtype, int math, int kernels, int64_t* x16_bytes) {
  if (dtype != EINCONV_F32 || kernels == EINCONV_KERNELS_GENERIC) return false;
  if (resolve_math(dtype, math) != EINCONV_MATH_TF32) return false;
  if (getenv("EINCONV_KFC_TF32")) return false;
  // measured on the ResNet-50 layer table (scripts/time_factors.py): the copy pays off for kernels with
  // several taps (the operand is re-read per tap) up to A = 2304; 1x1 layers and the A = 4608 layers are
  // faster on kind::tf32 without the extra pass
  long long max_a = 2304;
  if (const char* e = getenv("EINCONV_KFC_HALF_MAXA")) max_a = atoll(e);  // A/B measurements
  // narrow images (rows shorter than a 128-byte box in fp16) use the flat-plane operand layout of tc_tma.cu,
  // where fp16 positions are as densely packed as fp32 ones: the copy pays off for every A
  const bool flat = g.n_dims == 2 && g.k_total > 1 && g.k_total <= 25 && g.out[1] * 2 < 128 &&
                    !getenv("EINCONV_KFC_NO_FLAT");
  if (g.k_total == 1 || (!flat && (long long)g.c_in_g * g.k_total > max_a)) return false;
  DevGeom g1;
  int64_t ub;
  if (!tma::tma_supported(EINCONV_OP_KFC, g, EINCONV_F16) &&
      !kfc_via_unfold(g, EINCONV_F16, EINCONV_MATH_HALF, kernels, &g1, &ub))
    return false;
  const long long n = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
  *x16_bytes = 256 + ((n * 2 + 255) / 256) * 256;  // [header: absmax, 1/s^2][fp16 copy]
  return true;
}

// Dense rule of the simplifier (simplifications/opt.py:196-240: kernel size 1, stride 1, no padding => the
// index pattern is the identity): a 1x1 KFC factor is the plain Gram matrix X X^T over ALL positions of a
// sample, so the spatial dimensions collapse into one.  The TMA kernel then reads whole 128-byte lines of the
// contiguous (H W) axis instead of one short box per image row (14 x 14: 56-byte rows needed shifted copies).
static DevGeom kfc_canonical(const DevGeom& g) {
  if (g.k_total != 1 || g.n_dims == 1) return g;
  for (int i = 0; i < g.n_dims; ++i)
    if (g.stride[i] != 1 || g.pad[i] != 0 || g.out[i] != g.in[i]) return g;
  einconv_geom e{};
  e.n_dims = 1; e.groups = g.groups; e.batch = g.batch; e.c_in_g = g.c_in_g; e.c_out_g = 0;
  e.in[0] = g.in_total; e.k[0] = 1; e.stride[0] = 1; e.dilation[0] = 1; e.pad_left[0] = 0; e.out[0] = g.in_total;
  DevGeom d;
  return make_dev_geom(&e, &d, false) == EINCONV_OK ? d : g;
}

static int64_t kfac_u_bytes(const DevGeom& g) {
  return (((int64_t)g.batch * g.groups * g.c_in_g * g.k_total * 4 + 255) / 256) * 256;
}

static int check_math(int dtype, int math) {
  EINCONV_CHECK_ARG(math >= EINCONV_MATH_AUTO && math <= EINCONV_MATH_HALF, "unknown math mode %d", math);
  EINCONV_CHECK_ARG(elem_size(dtype) > 0, "unknown dtype %d", dtype);
  if (math == EINCONV_MATH_TF32 && dtype != EINCONV_F32) {
    set_error("EINCONV_MATH_TF32 requires f32 tensors");
    return EINCONV_ERR_UNSUPPORTED;
  }
  if (math == EINCONV_MATH_HALF && dtype != EINCONV_BF16 && dtype != EINCONV_F16) {
    set_error("EINCONV_MATH_HALF requires bf16/f16 tensors");
    return EINCONV_ERR_UNSUPPORTED;
  }
  return EINCONV_OK;
}

}  // namespace einconv

using namespace einconv;

extern "C" {

int einconv_abi_version(void) { return EINCONV_ABI_VERSION; }

const char* einconv_last_error(void) { return g_error; }

int64_t einconv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* einconv_selected_kernel(int op, const einconv_geom* geom, int dtype, int math,
                                    int kernels) {
  DevGeom g;
  const bool no_cout = op == EINCONV_OP_UNFOLD || op == EINCONV_OP_FOLD || op == EINCONV_OP_UNFOLD_TRANSPOSE ||
                       op == EINCONV_OP_KFAC_REDUCE_TRANSPOSE;
  if (make_dev_geom(geom, &g, !no_cout)) return nullptr;
  if (check_math(dtype, math)) return nullptr;
  switch (op) {
    case EINCONV_OP_UNFOLD: return "unfold_plane";
    case EINCONV_OP_FOLD: return "fold_gather";
    case EINCONV_OP_UNFOLD_TRANSPOSE: return "unfold_transpose_gather";
    default: break;
  }
  if (op == EINCONV_OP_KFC) {
    g = kfc_canonical(g);
    DevGeom g1;
    int64_t ub, hb;
    if (kfc_half_route(g, dtype, math, kernels, &hb))
      return kfc_via_unfold(g, EINCONV_F16, EINCONV_MATH_HALF, kernels, &g1, &ub) ? "f16scaled+unfold+tcgen05_tma_kfc"
                                                                                   : "f16scaled+tcgen05_tma_kfc";
    if (kfc_via_unfold(g, dtype, math, kernels, &g1, &ub)) return "unfold+tcgen05_tma_kfc";
  }
  if (op == EINCONV_OP_KFAC_REDUCE || op == EINCONV_OP_KFAC_REDUCE_TRANSPOSE) {
    DevGeom g2;
    return kfac_reduce_tensor(op == EINCONV_OP_KFAC_REDUCE ? g : swap_io(g), dtype, math, kernels, &g2)
               ? "window_sums+tcgen05_tma_kfc"
               : "window_sums+ffma_outer";
  }
  switch (select_path(op, g, dtype, math, kernels)) {
    case Path::DEPTHWISE: return "depthwise_stencil";
    case Path::TENSOR: return tc_kernel_name(op, g, dtype, resolve_math(dtype, math));
    case Path::TMA: return op == EINCONV_OP_KFC ? "tcgen05_tma_kfc" : "tcgen05_tma_wgrad";
    case Path::CONV_TMA: return op == EINCONV_OP_CONV_FWD ? "tcgen05_tma_conv_fwd" : "tcgen05_tma_conv_input_vjp";
    case Path::DENSE: return op == EINCONV_OP_CONV_FWD ? "tcgen05_dense_conv_fwd" : "tcgen05_dense_conv_input_vjp";
    case Path::CONV_ROWS: return op == EINCONV_OP_CONV_FWD ? "tcgen05_rows_conv_fwd" : "tcgen05_rows_conv_input_vjp";
    case Path::WGRAD_SLAB: return "tcgen05_slab_wgrad";
    default: return "ffma_gather_gemm";
  }
}

static int64_t workspace_for(int op, DevGeom g, int dtype, int math, int kernels, bool aligned);

int64_t einconv_workspace_bytes(int op, const einconv_geom* geom, int dtype, int math, int kernels) {
  DevGeom g;
  const bool no_cout = op == EINCONV_OP_UNFOLD || op == EINCONV_OP_FOLD || op == EINCONV_OP_UNFOLD_TRANSPOSE ||
                       op == EINCONV_OP_KFAC_REDUCE_TRANSPOSE;
  if (make_dev_geom(geom, &g, !no_cout)) return -1;
  if (check_math(dtype, math)) return -1;
  // the route depends on the alignment of the pointers, which this query does not see: enough for either
  return std::max(workspace_for(op, g, dtype, math, kernels, true), workspace_for(op, g, dtype, math, kernels, false));
}

static int64_t workspace_for(int op, DevGeom g, int dtype, int math, int kernels, bool aligned) {
  if (op == EINCONV_OP_UNFOLD || op == EINCONV_OP_FOLD || op == EINCONV_OP_UNFOLD_TRANSPOSE) return 0;
  if (op == EINCONV_OP_KFAC_REDUCE_TRANSPOSE) {
    // same two stages as KFAC-reduce, on the transpose-convolution view of the geometry
    op = EINCONV_OP_KFAC_REDUCE;
    g = swap_io(g);
  }
  if (op == EINCONV_OP_KFC) g = kfc_canonical(g);
  if (op == EINCONV_OP_KFC && aligned) {
    DevGeom g1;
    int64_t ub, hb;
    if (kfc_half_route(g, dtype, math, kernels, &hb)) {
      if (kfc_via_unfold(g, EINCONV_F16, EINCONV_MATH_HALF, kernels, &g1, &ub))
        return hb + ub + tma::tma_workspace_bytes(op, g1, EINCONV_F16);
      return hb + tma::tma_workspace_bytes(op, g, EINCONV_F16);
    }
    if (kfc_via_unfold(g, dtype, math, kernels, &g1, &ub)) return ub + tma::tma_workspace_bytes(op, g1, dtype);
  }
  const Path path = select_path(op, g, dtype, math, kernels, aligned);
  if (path == Path::DEPTHWISE || path == Path::DENSE) return 0;
  if (path == Path::TENSOR) return tc_workspace_bytes(op, g, dtype, resolve_math(dtype, math));
  if (path == Path::TMA) return tma::tma_workspace_bytes(op, g, dtype);
  if (path == Path::CONV_TMA) return ctma::workspace_bytes(op, g, dtype);
  if (path == Path::CONV_ROWS) return crows::workspace_bytes(op, g, dtype, resolve_math(dtype, math));
  if (path == Path::WGRAD_SLAB) return wslab::workspace_bytes(g, dtype);
  switch (op) {
    case EINCONV_OP_KFAC_REDUCE: {
      DevGeom g2;
      // stage 1 writes u into the (256-byte aligned) workspace, so stage 2 never sees the caller's pointer
      if (kfac_reduce_tensor(g, dtype, math, kernels, &g2))
        return kfac_u_bytes(g) + tma::tma_workspace_bytes(EINCONV_OP_KFC, g2, EINCONV_F32);
      return ffma_splitk_workspace(op, g, dtype, nullptr);
    }
    case EINCONV_OP_CONV_WEIGHT_VJP:
    case EINCONV_OP_KFC: return ffma_splitk_workspace(op, g, dtype, nullptr);
    default: return 0;
  }
}

int einconv_unfold(const void* x, const int64_t* x_strides_host, void* out, int dtype,
                   const einconv_geom* g, void* stream) {
  EINCONV_CHECK_ARG(x && out, "unfold: NULL tensor pointer");
  return launch_unfold(x, x_strides_host, out, dtype, g, (cudaStream_t)stream);
}

int einconv_fold(const void* gu, void* gx, int dtype, const einconv_geom* g, void* stream) {
  EINCONV_CHECK_ARG(gu && gx, "fold: NULL tensor pointer");
  return launch_fold(gu, gx, dtype, g, (cudaStream_t)stream);
}

int einconv_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                     const einconv_geom* geom, int math, int kernels, void* workspace,
                     int64_t workspace_bytes, void* stream) {
  EINCONV_CHECK_ARG(x && w && y, "conv_fwd: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, true);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  if ((long long)g.batch * g.groups * g.c_out_g * g.out_total == 0) return EINCONV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool al = aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias);
  switch (select_path(EINCONV_OP_CONV_FWD, g, dtype, math, kernels, al)) {
    case Path::DEPTHWISE: return dw_conv_fwd(x, w, bias, y, dtype, g, s);
    case Path::DENSE: return dense::run(EINCONV_OP_CONV_FWD, x, w, bias, y, dtype, g, s);
    case Path::CONV_ROWS:
      return crows::run(EINCONV_OP_CONV_FWD, x, w, bias, y, dtype, g, resolve_math(dtype, math), workspace,
                        workspace_bytes, s);
    case Path::CONV_TMA:
      return ctma::run(EINCONV_OP_CONV_FWD, x, w, bias, y, dtype, g, workspace, workspace_bytes, s);
    case Path::TENSOR:
      return tc_conv_fwd(x, w, bias, y, dtype, g, resolve_math(dtype, math), workspace, workspace_bytes, s);
    default: return ffma_conv_fwd(x, w, bias, y, dtype, g, s);
  }
}

int einconv_conv_input_vjp(const void* w, const void* v, void* gx, int dtype,
                           const einconv_geom* geom, int math, int kernels, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  EINCONV_CHECK_ARG(w && v && gx, "conv_input_vjp: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, true);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  if ((long long)g.batch * g.groups * g.c_in_g * g.in_total == 0) return EINCONV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool al = aligned16(w) && aligned16(v) && aligned16(gx);
  switch (select_path(EINCONV_OP_CONV_INPUT_VJP, g, dtype, math, kernels, al)) {
    case Path::DEPTHWISE: return dw_conv_input_vjp(w, v, gx, dtype, g, s);
    case Path::DENSE: return dense::run(EINCONV_OP_CONV_INPUT_VJP, v, w, nullptr, gx, dtype, g, s);
    case Path::CONV_ROWS:
      return crows::run(EINCONV_OP_CONV_INPUT_VJP, v, w, nullptr, gx, dtype, g, resolve_math(dtype, math), workspace,
                        workspace_bytes, s);
    case Path::CONV_TMA:
      return ctma::run(EINCONV_OP_CONV_INPUT_VJP, v, w, nullptr, gx, dtype, g, workspace, workspace_bytes, s);
    default: return ffma_conv_input_vjp(w, v, gx, dtype, g, s);
  }
}

int einconv_conv_weight_vjp(const void* x, const void* v, void* gw, int dtype,
                            const einconv_geom* geom, int math, int kernels, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  EINCONV_CHECK_ARG(x && v && gw, "conv_weight_vjp: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, true);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  if ((long long)g.groups * g.c_out_g * g.c_in_g * g.k_total == 0) return EINCONV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool al = aligned16(x) && aligned16(v) && aligned16(gw);
  switch (select_path(EINCONV_OP_CONV_WEIGHT_VJP, g, dtype, math, kernels, al)) {
    case Path::WGRAD_SLAB: return wslab::run(x, v, gw, dtype, g, workspace, workspace_bytes, s);
    case Path::TMA:
      return tma::tma_run(EINCONV_OP_CONV_WEIGHT_VJP, x, v, gw, dtype, g, 1.0, workspace, workspace_bytes, s);
    case Path::TENSOR:
      return tc_conv_weight_vjp(x, v, gw, dtype, g, resolve_math(dtype, math), workspace, workspace_bytes, s);
    default: return ffma_conv_weight_vjp(x, v, gw, dtype, g, workspace, workspace_bytes, s);
  }
}

int einconv_kfc(const void* x, void* out, int dtype, const einconv_geom* geom, double scale,
                int math, int kernels, void* workspace, int64_t workspace_bytes, void* stream) {
  EINCONV_CHECK_ARG(x && out, "kfc: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, false);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  if ((long long)g.groups * g.c_in_g * g.k_total == 0) return EINCONV_OK;
  g = kfc_canonical(g);
  cudaStream_t s = (cudaStream_t)stream;
  // the factor itself is written with scalar stores: only the input pointer matters
  const bool al = aligned16(x);
  if (al) {
    int64_t hb;
    if (kfc_half_route(g, dtype, math, kernels, &hb)) {
      DevGeom g1;
      int64_t ub = 0;
      const bool via_unfold = kfc_via_unfold(g, EINCONV_F16, EINCONV_MATH_HALF, kernels, &g1, &ub);
      const int64_t need = hb + ub + tma::tma_workspace_bytes(EINCONV_OP_KFC, via_unfold ? g1 : g, EINCONV_F16);
      if (workspace_bytes < need || !workspace) {
        set_error("kfc: workspace of %lld bytes required, got %lld", (long long)need, (long long)workspace_bytes);
        return EINCONV_ERR_WORKSPACE;
      }
      char* ws = (char*)workspace;
      __half* x16 = (__half*)(ws + 256);
      const float* inv_s2 = (const float*)ws + 1;
      const long long n = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
      // flat-plane plans read nothing but their copies of x, so the copy pass can convert by itself (16 launches
      // less per ResNet-50 sweep).  Measured neutral (16.99 vs 16.84 ms at batch 256, 4.235 vs 4.246 ms at batch 32:
      // every kernel-column copy re-reads x, now as fp32), hence opt-in
      if (!via_unfold && getenv("EINCONV_KFC_FUSED_CONVERT") && tma::flat_reads_copies_only(EINCONV_OP_KFC, g, EINCONV_F16)) {
        st = absmax_header((const float*)x, n, ws, s);
        if (st) return st;
        // (no fp16 tensor exists on this route: a null x makes any stray read of it fail loudly)
        return tma::tma_run_typed_out(EINCONV_OP_KFC, nullptr, nullptr, out, EINCONV_F16, EINCONV_F32, g, scale, ws + hb,
                                      workspace_bytes - hb, s, inv_s2, 0, (const float*)x, (const unsigned*)ws,
                                      (float*)ws + 1);
      }
      st = scaled_half_copy((const float*)x, x16, n, ws, s);
      if (st) return st;
      if (via_unfold) {
        st = launch_unfold(x16, nullptr, ws + hb, EINCONV_F16, geom, s);
        if (st) return st;
        return tma::tma_run_typed_out(EINCONV_OP_KFC, ws + hb, nullptr, out, EINCONV_F16, EINCONV_F32, g1, scale,
                                      ws + hb + ub, workspace_bytes - hb - ub, s, inv_s2);
      }
      return tma::tma_run_typed_out(EINCONV_OP_KFC, x16, nullptr, out, EINCONV_F16, EINCONV_F32, g, scale, ws + hb,
                                    workspace_bytes - hb, s, inv_s2);
    }
  }
  if (al) {
    DevGeom g1;
    int64_t ub;
    if (kfc_via_unfold(g, dtype, math, kernels, &g1, &ub)) {
      const int64_t need = ub + tma::tma_workspace_bytes(EINCONV_OP_KFC, g1, dtype);
      if (workspace_bytes < need || !workspace) {
        set_error("kfc: workspace of %lld bytes required, got %lld", (long long)need, (long long)workspace_bytes);
        return EINCONV_ERR_WORKSPACE;
      }
      st = launch_unfold(x, nullptr, workspace, dtype, geom, s);
      if (st) return st;
      return tma::tma_run(EINCONV_OP_KFC, workspace, nullptr, out, dtype, g1, scale, (char*)workspace + ub,
                          workspace_bytes - ub, s);
    }
  }
  switch (select_path(EINCONV_OP_KFC, g, dtype, math, kernels, al)) {
    case Path::TMA:
      return tma::tma_run(EINCONV_OP_KFC, x, nullptr, out, dtype, g, scale, workspace, workspace_bytes, s);
    case Path::TENSOR:
      return tc_kfc(x, out, dtype, g, scale, resolve_math(dtype, math), workspace, workspace_bytes, s);
    default: return ffma_kfc(x, out, dtype, g, scale, workspace, workspace_bytes, s);
  }
}

static int kfac_reduce_impl(const void* x, void* out, int dtype, const DevGeom& g, double scale, int math,
                            int kernels, void* workspace, int64_t workspace_bytes, void* stream) {
  if ((long long)g.groups * g.c_in_g * g.k_total == 0) return EINCONV_OK;
  int st;
  DevGeom g2;
  if (kfac_reduce_tensor(g, dtype, math, kernels, &g2)) {
    const int64_t u_bytes = kfac_u_bytes(g);
    const int64_t need = u_bytes + tma::tma_workspace_bytes(EINCONV_OP_KFC, g2, EINCONV_F32);
    if (workspace_bytes < need || !workspace) {
      set_error("kfac_reduce: workspace of %lld bytes required, got %lld", (long long)need,
                (long long)workspace_bytes);
      return EINCONV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    st = kfac_window_sums(x, workspace, dtype, g, /*transposed=*/1, s);
    if (st) return st;
    // the finish kernel of the tensor-core path casts to `dtype`; u itself is fp32
    return tma::tma_run_typed_out(EINCONV_OP_KFC, workspace, nullptr, out, EINCONV_F32, dtype, g2, scale,
                                  (char*)workspace + u_bytes, workspace_bytes - u_bytes, s);
  }
  return ffma_kfac_reduce(x, out, dtype, g, scale, workspace, workspace_bytes, (cudaStream_t)stream);
}

int einconv_kfac_reduce(const void* x, void* out, int dtype, const einconv_geom* geom, double scale,
                        int math, int kernels, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  EINCONV_CHECK_ARG(x && out, "kfac_reduce: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, false);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  return kfac_reduce_impl(x, out, dtype, g, scale, math, kernels, workspace, workspace_bytes, stream);
}

// stage-2 geometry over `n_shards` gathered blocks: a 1x1 "convolution" over 1-d images of B_shard positions
static bool kfac_shard_geom(const DevGeom& g, int n_shards, DevGeom* g2) {
  const long long A = (long long)g.c_in_g * g.k_total;
  if (A >= (1ll << 24) || g.batch < 1 || n_shards < 1) return false;
  einconv_geom e{};
  e.n_dims = 1; e.groups = g.groups; e.batch = n_shards; e.c_in_g = A; e.c_out_g = 0;
  e.in[0] = g.batch; e.k[0] = 1; e.stride[0] = 1; e.dilation[0] = 1; e.pad_left[0] = 0; e.out[0] = g.batch;
  return make_dev_geom(&e, g2, false) == EINCONV_OK;
}

static bool kfac_from_sums_route(const DevGeom& g, int n_shards, int dtype, int math, int kernels, DevGeom* g2) {
  DevGeom tmp;
  if (!kfac_reduce_tensor(g, dtype, math, kernels, &tmp)) return false;  // same conditions as the fused call
  if (!kfac_shard_geom(g, n_shards, g2)) return false;
  if ((g.batch * 4) % 16 != 0) return false;  // rows of u must be directly readable by TMA
  return tma::tma_supported(EINCONV_OP_KFC, *g2, EINCONV_F32);
}

int einconv_kfac_reduce_sums(const void* x, void* u, int dtype, const einconv_geom* geom, void* stream) {
  EINCONV_CHECK_ARG(x && u, "kfac_reduce_sums: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, false);
  if (st) return st;
  EINCONV_CHECK_ARG(elem_size(dtype) > 0, "unknown dtype %d", dtype);
  if ((long long)g.batch * g.groups * g.c_in_g * g.k_total == 0) return EINCONV_OK;
  return kfac_window_sums(x, u, dtype, g, /*transposed=*/1, (cudaStream_t)stream);
}

int64_t einconv_kfac_reduce_from_sums_workspace(int n_shards, int dtype, const einconv_geom* geom, int math,
                                                int kernels) {
  DevGeom g, g2;
  if (make_dev_geom(geom, &g, false)) return -1;
  if (check_math(dtype, math)) return -1;
  if (!kfac_from_sums_route(g, n_shards, dtype, math, kernels, &g2)) return -1;
  return tma::tma_workspace_bytes(EINCONV_OP_KFC, g2, EINCONV_F32);
}

int einconv_kfac_reduce_from_sums(const void* u, int n_shards, int64_t shard_stride, void* out, int dtype,
                                  const einconv_geom* geom, double scale, int math, int kernels, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  EINCONV_CHECK_ARG(u && out, "kfac_reduce_from_sums: NULL tensor pointer");
  DevGeom g, g2;
  int st = make_dev_geom(geom, &g, false);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  if (!kfac_from_sums_route(g, n_shards, dtype, math, kernels, &g2) || !aligned16(u) || (shard_stride * 4) % 16 != 0) {
    set_error("kfac_reduce_from_sums: no tensor-core route for this geometry / math mode / alignment");
    return EINCONV_ERR_UNSUPPORTED;
  }
  const long long block = (long long)g.groups * g.c_in_g * g.k_total * g.batch;
  EINCONV_CHECK_ARG(n_shards == 1 || shard_stride >= block, "kfac_reduce_from_sums: shard_stride %lld < block of %lld",
                    (long long)shard_stride, block);
  return tma::tma_run_typed_out(EINCONV_OP_KFC, u, nullptr, out, EINCONV_F32, dtype, g2, scale, workspace,
                                workspace_bytes, (cudaStream_t)stream, nullptr,
                                (n_shards > 1 && shard_stride != block) ? shard_stride : 0);
}

int einconv_unfold_transpose(const void* x, void* out, int dtype, const einconv_geom* g, void* stream) {
  EINCONV_CHECK_ARG(x && out, "unfold_transpose: NULL tensor pointer");
  return launch_unfold_transpose(x, out, dtype, g, (cudaStream_t)stream);
}

int einconv_kfac_reduce_transpose(const void* x, void* out, int dtype, const einconv_geom* geom, double scale,
                                  int math, int kernels, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
  EINCONV_CHECK_ARG(x && out, "kfac_reduce_transpose: NULL tensor pointer");
  DevGeom g;
  int st = make_dev_geom(geom, &g, false);
  if (st) return st;
  if ((st = check_math(dtype, math))) return st;
  // x lives on the convolution's OUTPUT grid: swap the roles and use the transpose membership rule
  return kfac_reduce_impl(x, out, dtype, swap_io(g), scale, math, kernels, workspace, workspace_bytes, stream);
}

}  // extern "C"
