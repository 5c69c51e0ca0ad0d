// Internal prototypes shared between api.cu and the kernel translation units.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace einconv {

// unfold.cu
int launch_unfold(const void* x, const int64_t* x_strides, void* out, int dtype,
                  const einconv_geom* geom, cudaStream_t stream);
int launch_fold(const void* gu, void* gx, int dtype, const einconv_geom* geom, cudaStream_t stream);
int launch_unfold_transpose(const void* x, void* out, int dtype, const einconv_geom* geom, cudaStream_t stream);

// contract_ffma.cu
int ffma_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                  const DevGeom& g, cudaStream_t stream);
int ffma_conv_input_vjp(const void* w, const void* v, void* gx, int dtype, const DevGeom& g,
                        cudaStream_t stream);
int ffma_conv_weight_vjp(const void* x, const void* v, void* gw, int dtype, const DevGeom& g,
                         void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int ffma_kfc(const void* x, void* out, int dtype, const DevGeom& g, double scale, void* workspace,
             int64_t workspace_bytes, cudaStream_t stream);
int ffma_kfac_reduce(const void* x, void* out, int dtype, const DevGeom& g, double scale,
                     void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int64_t ffma_splitk_workspace(int op, const DevGeom& g, int dtype, int* splits_out);
int kfac_window_sums(const void* x, void* u_out, int dtype, const DevGeom& g, int transposed,
                     cudaStream_t stream);

// depthwise.cu
bool depthwise_supported(const DevGeom& g, int dtype);
int dw_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                const DevGeom& g, cudaStream_t stream);
int dw_conv_input_vjp(const void* w, const void* v, void* gx, int dtype, const DevGeom& g,
                      cudaStream_t stream);

// tc_gemm.cu (tcgen05 / TMEM tensor-core kernels)
bool tc_supported(int op, const DevGeom& g, int dtype, int math);
const char* tc_kernel_name(int op, const DevGeom& g, int dtype, int math);
int64_t tc_workspace_bytes(int op, const DevGeom& g, int dtype, int math);
int tc_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                const DevGeom& g, int math, void* workspace, int64_t workspace_bytes,
                cudaStream_t stream);
int tc_conv_weight_vjp(const void* x, const void* v, void* gw, int dtype, const DevGeom& g,
                       int math, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int tc_kfc(const void* x, void* out, int dtype, const DevGeom& g, double scale, int math,
           void* workspace, int64_t workspace_bytes, cudaStream_t stream);


// convert.cu
int scaled_half_copy(const float* x, __half* y, long long n, void* header, cudaStream_t stream);
int absmax_header(const float* x, long long n, void* header, cudaStream_t stream);  // header[0] only
int launch_absmax(const float* x, long long n, unsigned* out_bits, cudaStream_t stream);
int launch_absmax2(const float* a, long long na, const float* b, long long nb, unsigned* out_bits,
                   cudaStream_t stream);

// tc_tma.cu (TMA-fed tcgen05 kernels for the reduction networks)
namespace tma {
bool tma_supported(int op, const DevGeom& g, int dtype);
int64_t tma_workspace_bytes(int op, const DevGeom& g, int dtype);
int tma_run(int op, const void* x, const void* v, void* out, int dtype, const DevGeom& g, double scale,
            void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int tma_run_typed_out(int op, const void* x, const void* v, void* out, int dtype, int out_dtype, const DevGeom& g,
                      double scale, void* workspace, int64_t workspace_bytes, cudaStream_t stream,
                      const float* scale_dev = nullptr, long long x_batch_stride = 0, const float* x_f32 = nullptr,
                      const unsigned* absmax_bits = nullptr, float* inv_s2 = nullptr);
// KFC plans whose kernel reads only the flat-plane copies of x (the copy pass may then convert fp32 -> scaled fp16)
bool flat_reads_copies_only(int op, const DevGeom& g, int dtype);
}  // namespace tma

// conv_tma.cu (TMA-fed tcgen05 implicit GEMM for forward / input VJP, stride 1)
namespace ctma {
bool supported(int op, const DevGeom& g, int dtype, int math);
int64_t workspace_bytes(int op, const DevGeom& g, int dtype);
int run(int op, const void* src, const void* w, const void* bias, void* dst, int dtype, const DevGeom& g,
        void* workspace, int64_t workspace_bytes, cudaStream_t stream);
// shared with wgrad_slab.cu
struct WeightJob {  // optional weight packing fused into the same launch (conv_tma.cu)
  const void* src;
  void* dst;
  int groups, taps, np, cgp, cog, cg, transpose;
  long long total;
};
int pack_channels_last(const void* src, void* xp, int dtype, int n_dims, int batch, int ctot, int cp,
                       const int* S, const long long* P, const int* pad, cudaStream_t stream,
                       const WeightJob* wj = nullptr, const int* sstep = nullptr, const int* soff = nullptr);
// same layout, fp32 source -> fp16 destination scaled by operand_scale(*absmax_bits) (convert.cu)
// an optional weight job is converted with the scale of absmax_bits[1]
int pack_channels_last_f32_to_f16(const float* src, void* xp, int n_dims, int batch, int ctot, int cp, const int* S,
                                  const long long* P, const int* pad, const unsigned* absmax_bits,
                                  cudaStream_t stream, const WeightJob* wj = nullptr);
int encode_2d(CUtensorMap* map, int dtype, const void* base, long long rows, long long cols, int box_cols,
              int box_rows, int swizzle_bytes);
int encode_tiled_raw(CUtensorMap* map, const void* base, int elem_bytes, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box, const unsigned* elem_strides);
}  // namespace ctma

// conv_dense.cu (1x1 convolutions as GEMMs on the channels-first tensors, no packed copies)
namespace dense {
bool supported(int op, const DevGeom& g, int dtype, int math);
int run(int op, const void* src, const void* w, const void* bias, void* dst, int dtype, const DevGeom& g,
        cudaStream_t stream);
}  // namespace dense

// conv_rows.cu (stride-1 2-d forward from the channels-first tensor: kh as TMA boxes, kw in the epilogue)
namespace crows {
bool supported(int op, const DevGeom& g, int dtype, int math);
int64_t workspace_bytes(int op, const DevGeom& g, int dtype, int math);
int run(int op, const void* x, const void* w, const void* bias, void* y, int dtype, const DevGeom& g, int math,
        void* workspace, int64_t workspace_bytes, cudaStream_t stream);
}  // namespace crows

// wgrad_slab.cu (weight VJP from channels-last padded copies, MN-major tcgen05, stride 1, 16-bit)
namespace wslab {
bool supported(const DevGeom& g, int dtype);
int64_t workspace_bytes(const DevGeom& g, int dtype);
int run(const void* x, const void* v, void* gw, int dtype, const DevGeom& g, void* workspace,
        int64_t workspace_bytes, cudaStream_t stream);
}  // namespace wslab

}  // namespace einconv
