/*
 * einconv_b200 — C ABI of the sm_100a evaluator for einconv's convolution tensor networks.
 *
 * The reference (f-dangel/einconv, pure Python) has no FFI layer: its hot path is the call
 *     torch.einsum(equation, *operands).reshape(shape)
 * at einconv/functionals/conv.py:61, einconv/functionals/unfold.py:56 and, for the VJPs and
 * Kronecker factors, on the user side (README.md:64-73).  Each entry point below replaces that
 * call for one of the six networks einconv/expressions/*.py emit.  The one-hot index-pattern
 * operands (einconv/conv_index_pattern.py:12-85) are never passed: the kernels evaluate
 *     i_d = stride_d * o_d + dilation_d * k_d - pad_left_d ,  valid iff 0 <= i_d < in_d
 * (conv_index_pattern.py:131-134, einconv/utils.py:97-160) in registers.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *  - the caller owns all memory (inputs, outputs, workspace); the library never allocates,
 *    frees or keeps device memory and never synchronises the stream;
 *  - tensors are dense, channels-first, innermost spatial dimension contiguous, exactly the
 *    layout of the reference's operands: x [B, G*Cin_g, in_0..in_{N-1}], weight
 *    [G*Cout_g, Cin_g, k_0..k_{N-1}], y / v [B, G*Cout_g, out_0..out_{N-1}];
 *  - every function returns 0 on success or a negative einconv_status; a message for the last
 *    failure of the calling thread is available from einconv_last_error();
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default).
 */
#ifndef EINCONV_B200_H_
#define EINCONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EINCONV_MAX_DIMS 8
#define EINCONV_ABI_VERSION 1

typedef enum {
  EINCONV_OK = 0,
  EINCONV_ERR_INVALID = -1,      /* bad argument (NULL pointer, size <= 0, unsupported combination) */
  EINCONV_ERR_UNSUPPORTED = -2,  /* valid request, no kernel for it (e.g. fp64 on tensor cores) */
  EINCONV_ERR_WORKSPACE = -3,    /* workspace too small; see einconv_workspace_bytes */
  EINCONV_ERR_CUDA = -4          /* a CUDA runtime / driver call failed */
} einconv_status;

typedef enum {
  EINCONV_F32 = 0,
  EINCONV_BF16 = 1,
  EINCONV_F16 = 2,
  EINCONV_F64 = 3
} einconv_dtype;

/* How the multiply-accumulate of a contraction is carried out.
 *  AUTO   : FP32 (f32/f64 I/O) or native 16-bit tensor cores (bf16/f16 I/O)
 *  FP32   : CUDA-core FMA in fp32 (fp64 for f64 I/O); meets the reference's rtol 1e-5
 *  TF32   : tcgen05 kind::tf32 on f32 I/O, fp32 accumulate in TMEM.  The KFC factor and the weight
 *           VJP may instead run kind::f16 on operand copies scaled by a power of two and rounded to
 *           fp16 — the same 10 explicit mantissa bits as TF32, round-to-nearest instead of truncation —
 *           whenever that is faster (half the operand bytes, twice the MMA rate); the stated 1e-3
 *           tolerance holds.
 *  HALF   : tcgen05 kind::f16 on bf16/f16 I/O, fp32 accumulate in TMEM
 */
typedef enum {
  EINCONV_MATH_AUTO = 0,
  EINCONV_MATH_FP32 = 1,
  EINCONV_MATH_TF32 = 2,
  EINCONV_MATH_HALF = 3
} einconv_math;

/* Which kernel family may be used; GENERIC corresponds to the reference's simplify=False
 * (one code path for every geometry), SPECIALIZED to simplify=True (1x1/dense, depthwise,
 * down-sampling variants picked from the predicates of simplifications/opt.py:196-300). */
typedef enum {
  EINCONV_KERNELS_SPECIALIZED = 0,
  EINCONV_KERNELS_GENERIC = 1
} einconv_kernels;

/* Geometry of one N-d convolution; the C counterpart of what the reference stores per
 * dimension in pattern._pattern_hyperparams (conv_index_pattern.py:77-83) plus the channel
 * bookkeeping of expressions/convNd_forward.py:69-77.  `out` must equal
 * 1 + floor((in + pad_left + pad_right - (k + (k-1)(dilation-1))) / stride) (utils.py:158-160);
 * pad_right enters only through `out`. */
typedef struct {
  int32_t n_dims;   /* number of spatial dimensions N, 1..EINCONV_MAX_DIMS */
  int32_t groups;   /* G */
  int64_t batch;    /* B */
  int64_t c_in_g;   /* input channels per group  */
  int64_t c_out_g;  /* output channels per group (ignored by unfold / kfc / kfac_reduce) */
  int32_t in[EINCONV_MAX_DIMS];
  int32_t k[EINCONV_MAX_DIMS];
  int32_t stride[EINCONV_MAX_DIMS];
  int32_t dilation[EINCONV_MAX_DIMS];
  int32_t pad_left[EINCONV_MAX_DIMS];
  int32_t out[EINCONV_MAX_DIMS];
} einconv_geom;

typedef enum {
  EINCONV_OP_UNFOLD = 0,
  EINCONV_OP_CONV_FWD = 1,
  EINCONV_OP_CONV_INPUT_VJP = 2,
  EINCONV_OP_CONV_WEIGHT_VJP = 3,
  EINCONV_OP_KFC = 4,
  EINCONV_OP_KFAC_REDUCE = 5,
  EINCONV_OP_FOLD = 6,
  EINCONV_OP_UNFOLD_TRANSPOSE = 7,
  EINCONV_OP_KFAC_REDUCE_TRANSPOSE = 8
} einconv_op;

int einconv_abi_version(void);
const char* einconv_last_error(void);

/* Bytes of scratch an op needs for this geometry / dtype / math mode (0 if none). */
int64_t einconv_workspace_bytes(int op, const einconv_geom* g, int dtype, int math, int kernels);

/* unfoldNd (im2col).  Replaces torch.einsum at einconv/functionals/unfold.py:56 for the network
 * of expressions/convNd_unfold.py:42-53:  out[b, c, k.., o..] = x[b, c, i(o,k)] or 0.
 * x: [B, C, in..] with element strides x_strides[0..N+1] (NULL = contiguous);
 * out: contiguous [B, C*prod(k), prod(out)].  Bit-exact copy; g->groups and c_out_g ignored,
 * C = g->groups * g->c_in_g. */
int einconv_unfold(const void* x, const int64_t* x_strides_host, void* out, int dtype,
                   const einconv_geom* g, void* stream);

/* Adjoint of unfold (col2im): gx[b,c,i] = sum_{k,o : i(o,k)=i} gu[b, c, k.., o..].
 * Backward of einconv_unfold for autograd; gu contiguous [B, C*prod(k), prod(out)]. */
int einconv_fold(const void* gu, void* gx, int dtype, const einconv_geom* g, void* stream);

/* convNd forward.  Replaces torch.einsum at einconv/functionals/conv.py:61 (network of
 * expressions/convNd_forward.py:48-55) and the bias add at conv.py:63-67.
 * y[b, g*Cout_g+co, o..] = bias[.] + sum_{ci,k..} x[b, g*Cin_g+ci, i(o,k)..] * w[g*Cout_g+co, ci, k..]
 * bias may be NULL. */
int einconv_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                     const einconv_geom* g, int math, int kernels, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* convNd input VJP (network of expressions/convNd_input_vjp.py:53-60).
 * gx[b, g*Cin_g+ci, i..] = sum_{co,k..,o.. : i(o,k)=i} v[b, g*Cout_g+co, o..] * w[g*Cout_g+co, ci, k..] */
int einconv_conv_input_vjp(const void* w, const void* v, void* gx, int dtype,
                           const einconv_geom* g, int math, int kernels, void* workspace,
                           int64_t workspace_bytes, void* stream);

/* convNd weight VJP (network of expressions/convNd_weight_vjp.py:51-58).
 * gw[g*Cout_g+co, ci, k..] = sum_{b,o..} x[b, g*Cin_g+ci, i(o,k)..] * v[b, g*Cout_g+co, o..] */
int einconv_conv_weight_vjp(const void* x, const void* v, void* gw, int dtype,
                            const einconv_geom* g, int math, int kernels, void* workspace,
                            int64_t workspace_bytes, void* stream);

/* KFC factor (network of expressions/convNd_kfc.py:75-106), A = Cin_g * prod(k):
 * out[g, (ci,k..), (ci',k'..)] = scale * sum_{b,o..} U[b,g,ci,k..,o..] * U[b,g,ci',k'..,o..]
 * with U the (never materialised) unfolded input.  The reference uses scale = 1/B
 * (convNd_kfc.py:105); a rank holding a batch shard passes the GLOBAL 1/B so that a plain SUM
 * all-reduce of the per-rank outputs reproduces it.  out: contiguous fp32/fp64/bf16/f16 [G, A, A]
 * in `dtype`. */
int einconv_kfc(const void* x, void* out, int dtype, const einconv_geom* g, double scale,
                int math, int kernels, void* workspace, int64_t workspace_bytes, void* stream);

/* KFAC-reduce factor (network of expressions/convNd_kfac_reduce.py:77-108):
 * out[g, a, a'] = scale * sum_b u[b,g,a] * u[b,g,a'],  u[b,g,(ci,k..)] = sum_{o..} U[b,g,ci,k..,o..];
 * reference scale = 1 / (B * prod(out)^2). */
int einconv_kfac_reduce(const void* x, void* out, int dtype, const einconv_geom* g, double scale,
                        int math, int kernels, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* ---- KFAC-reduce in two stages, for batch-sharded evaluation.  The factor is
 *   out = scale * sum_b u_b u_b^T,   u[b, g, (ci,k..)] = sum_{o..} U[b,g,ci,k..,o..]
 * (expressions/convNd_kfac_reduce.py:77-108: the second pattern copy shares no index with the first, so
 * the network factorises over the window sums u).  u is tiny (B x A numbers against A x A for the factor):
 * ranks exchange u (all-gather) instead of all-reducing factors, and every rank then forms the full-batch
 * factor — bit-identical on all ranks and to a single-process evaluation of the concatenated batch.
 *
 * Stage 1: u[g*A + a][b] (fp32; fp64 for f64 input), contiguous [G*A][B] — positions b innermost. */
int einconv_kfac_reduce_sums(const void* x, void* u, int dtype, const einconv_geom* g, void* stream);

/* Stage 2 over `n_shards` blocks of window sums, block s at u + s * shard_stride (elements), each laid out
 * [G*A][B_shard] with B_shard = g->batch:  out[g,a,a'] = scale * sum_s sum_b u_s[g,a,b] u_s[g,a',b].
 * Tensor-core route only (math TF32 / HALF, A >= 256, 4 * B_shard and 4 * shard_stride multiples of 16
 * bytes); otherwise EINCONV_ERR_UNSUPPORTED and the caller sums factors instead.  out: [G, A, A] in `dtype`. */
int einconv_kfac_reduce_from_sums(const void* u, int n_shards, int64_t shard_stride, void* out, int dtype,
                                  const einconv_geom* g, double scale, int math, int kernels, void* workspace,
                                  int64_t workspace_bytes, void* stream);
/* Workspace of stage 2 (bytes), or -1 if the tensor-core route does not apply. */
int64_t einconv_kfac_reduce_from_sums_workspace(int n_shards, int dtype, const einconv_geom* g, int math,
                                                int kernels);

/* ---- transpose convolutions.  The geometry always describes the ASSOCIATED
 * convolution: `in` = its input sizes (the transpose convolution's output), `out` = its output sizes
 * (the spatial sizes of the tensor x passed here), with out = 1 + floor((in + 2p - d(k-1) - 1) / s);
 * output_padding is already folded into `in` (einconv/utils.py:163-205). ---- */

/* Input unfolding of a transpose convolution.  Replaces torch.einsum at
 * einconv/functionals/unfold_transpose.py:70 for the network of
 * expressions/conv_transposeNd_unfold.py:81-91:
 *   out[b, c, k.., i..] = x[b, c, o..]  if  i_d = s_d o_d + d_d k_d - p_d  for every d,  else 0.
 * x: contiguous [B, C, out..], C = g->groups * g->c_in_g; out: contiguous [B, C*prod(k), prod(in)].
 * Bit-exact copy. */
int einconv_unfold_transpose(const void* x, void* out, int dtype, const einconv_geom* g, void* stream);

/* Input-based KFAC-reduce factor of a transpose convolution (network of
 * expressions/conv_transposeNd_kfac_reduce.py:95-110), A = C_g * prod(k):
 *   out[g, a, a'] = scale * sum_b u[b,g,a] u[b,g,a'],
 *   u[b,g,(c,k..)] = sum_{o.. : 0 <= s o + d k - p < in} x[b, g*C_g + c, o..];
 * reference scale = 1 / (B * prod(in)^2).  x: [B, G*C_g, out..] with C_g = g->c_in_g. */
int einconv_kfac_reduce_transpose(const void* x, void* out, int dtype, const einconv_geom* g, double scale,
                                  int math, int kernels, void* workspace, int64_t workspace_bytes,
                                  void* stream);

/* Name of the kernel family the dispatcher would pick (static string; for tests, logs and the
 * launch counter in bench.py).  Returns NULL on invalid arguments. */
const char* einconv_selected_kernel(int op, const einconv_geom* g, int dtype, int math, int kernels);

/* Number of kernel launches issued by this library in the calling process so far. */
int64_t einconv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EINCONV_B200_H_ */
