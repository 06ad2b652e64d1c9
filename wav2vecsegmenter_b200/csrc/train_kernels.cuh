// Launchers of the head-only training step's CUDA-core kernels (train_kernels.cu). Device pointers + stream;
// 0 / W2VSEG_ERR_* return codes. Reference: train.py:381-480 (loss, backward) around lib/models.py:279-319.
#pragma once
#include "common.h"
#include "dropout.cuh"

namespace w2v {

// final LayerNorm + Linear(1024 -> 1) + masked BCE-with-logits (pos_weight) and the backward down to dx2.
// loss = sum over unmasked frames / B (reference: loss_per_point.sum(dim=1).mean()).
int head_loss_backward_launch(const float* x2, int B, int R, const float* gamma, const float* beta, float eps,
                              const float* w_out, const float* b_out, const int32_t* out_len, const float* target,
                              float pos_weight, float* dx2, __nv_bfloat16* dx2_bf, float* dlogit, float2* stats,
                              float* logits, float* loss_rows, float* loss, cudaStream_t s);
// LayerNorm(1024) backward wrt the input; dx_out = dx_in + ..., either may be null (statistics only)
int layernorm_bwd_launch(const float* x, const __nv_bfloat16* dy, int64_t rows, const float* gamma, float eps,
                         const float* dx_in, float* dx_out, __nv_bfloat16* dx_bf, float2* stats, cudaStream_t s);
// out[c] = sum_r a[r, c] (bias gradients); a bf16 or fp32 with leading dimension lda
int colsum_launch(const void* a, bool a_bf16, int64_t lda, int64_t rows, int C, float* scratch,
                  size_t scratch_floats, float* out, cudaStream_t s);
// d_gamma[c] = sum_r dy[r, c] * xhat[r, c], d_beta[c] = sum_r dy[r, c]
int ln_param_grads_launch(const __nv_bfloat16* dy, const float* x, const float2* stats, int64_t rows, int C,
                          float* scratch, size_t scratch_floats, float* d_gamma, float* d_beta, cudaStream_t s);
// gradients of output_layer.{weight,bias} and layer_norm.{weight,bias} from the per-row dlogit
int final_param_grads_launch(const float* dlogit, const float* x2, const float2* stats, int64_t rows, int C,
                             const float* gamma, const float* beta, const float* w_out, float* scratch,
                             size_t scratch_floats, float* tmpA, float* tmpS, float* d_w, float* d_gamma,
                             float* d_beta, float* d_b, cudaStream_t s);
// out = dropout(gelu(z)) / dz = dm * mask * gelu'(z); the mask index is the flat element index (d.thresh 0: none)
int gelu_fwd_launch(const __nv_bfloat16* z, __nv_bfloat16* out, int64_t n, const DropSite& d, cudaStream_t s);
int gelu_bwd_launch(const __nv_bfloat16* z, const __nv_bfloat16* dm, __nv_bfloat16* dz, int64_t n, const DropSite& d,
                    cudaStream_t s);
// x_out = x_in + dropout(y), fp32 [n]
int resid_dropout_launch(const float* x_in, const float* y, float* x_out, int64_t n, const DropSite& d, cudaStream_t s);
// dst (bf16) = src (fp32) * dropout mask: the gradient entering a dropped residual branch
int mask_cast_launch(const float* src, __nv_bfloat16* dst, int64_t n, const DropSite& d, cudaStream_t s);
// gather_rows + dropout (init_dropout on the encoder output)
int gather_dropout_launch(const float* src, int64_t batch_stride, int B, int T, int C, float* dst, const DropSite& d,
                          cudaStream_t s);
// dst[c, r] = src[r, c] (r < rows), zero for rows <= r < rows_pad; dst leading dimension rows_pad
int transpose_bf16_launch(const __nv_bfloat16* src, int64_t ld_src, int64_t rows, int cols, __nv_bfloat16* dst,
                          int64_t rows_pad, cudaStream_t s);
int copy_f32_launch(const float* src, float* dst, int64_t n, cudaStream_t s);

}  // namespace w2v
