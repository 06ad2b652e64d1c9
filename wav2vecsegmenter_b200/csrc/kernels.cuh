// Launchers for the memory-bound / CUDA-core kernels of the SFC path (kernels.cu) and the fused
// attention kernel (attention.cu). All take device pointers and a stream; all return 0 / W2VSEG_ERR_*.
#pragma once
#include "common.h"
#include "dropout.cuh"

namespace w2v {

// per-window normalisation statistics (lib/datautils.py:122-125) + valid encoder frames.
// stats[b] = (mean, 1/std) over the zero-padded row of norm_len[b] samples; (0, 1) if norm_len==0.
int window_stats_launch(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                        const int32_t* norm_len, int B, double2* partial /*[B*64] scratch*/,
                        float2* stats, int32_t* enc_len, int32_t* included, cudaStream_t s);

// conv layer 0 (1->512, k=10, s=5) + LayerNorm(512) + GELU, input normalisation fused (HF:281-299).
// out bf16 [B*R0, 512], channels-last.
int conv0_ln_gelu_launch(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                         const float2* stats, const float* w_t /*[10][512]*/, const float* bias,
                         const float* gamma, const float* beta, float eps, __nv_bfloat16* out,
                         int B, int R0, cudaStream_t s);

// GroupNorm variant of layer 0 (HF feat_extract_norm="group", HF:302-323): conv -> GroupNorm(512 groups: per
// channel over the frames of the padded row, norm_len[b] samples, or l_max if norm_len[b] == 0) -> GELU.
// scratch: conv0_gn_scratch_floats(B, R0) floats. bias may be null (config.conv_bias = False).
size_t conv0_gn_scratch_floats(int B, int R0);
int conv0_gn_gelu_launch(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                         const int32_t* norm_len, int l_max, const float2* stats, const float* w_t,
                         const float* bias, const float* gamma, const float* beta, float eps,
                         float* scratch, __nv_bfloat16* out, int B, int R0, cudaStream_t s);

// same layer on the tensor cores (conv0_tc.cu): LayerNorm folded into ONE K=16 fp16 tcgen05.mma per
// 128-frame x 256-channel tile, GELU in the epilogue. `pack` = conv0_tc_pack_bytes() device bytes
// filled once per weight set by conv0_tc_pack_launch (w element (c,k) at w[c*sc + k*sk]).
size_t conv0_tc_pack_bytes();
int conv0_tc_pack_launch(const float* w, int sc, int sk, const float* bias, const float* gamma,
                         const float* beta, void* pack, cudaStream_t s);
int conv0_tc_launch(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                    const float2* stats, const void* pack, float eps, __nv_bfloat16* out, int B,
                    int R0, cudaStream_t s);

// LayerNorm over C (512 or 1024) per row, optional GELU; in fp32 or bf16, out bf16. in==out allowed
// for bf16 input.
int layernorm_launch(const void* in, bool in_f32, int64_t rows, int C, const float* gamma,
                     const float* beta, float eps, int act, __nv_bfloat16* out, cudaStream_t s);

// post-LayerNorm encoders: h (fp32 [rows, 1024]) = LayerNorm(h) in place, plus a bf16 copy for the next GEMM
int layernorm_dual_launch(float* h, int64_t rows, const float* gamma, const float* beta, float eps,
                          __nv_bfloat16* out_bf16, cudaStream_t s);

// h fp32 [B*R, C] -> zpad bf16 [B*(R+pad2), C] interior rows (64-row zero halo each side is
// memset by the caller). Feeds the positional conv (HF:360-368).
int cast_to_padded_launch(const float* h, int B, int R, int C, int halo, __nv_bfloat16* zpad,
                          cudaStream_t s);

// gather a strided fp32 [B, T, C] view into contiguous [B*T, C] (head entry point)
int gather_rows_launch(const float* src, int64_t batch_stride, int B, int T, int C, float* dst,
                       cudaStream_t s);

// final LayerNorm(1024) + Linear(1024->1) + sigmoid + masking (lib/models.py:317-319,
// lib/evaluate.py:82-91). logits/probs [B*R] (either may be null).
int head_final_launch(const float* y, int B, int R, int C, const float* gamma, const float* beta,
                      float eps, const float* w_out, const float* b_out, const int32_t* out_len,
                      float* logits, float* probs, int64_t prob_stride, int row_cols, int flag_col,
                      const int32_t* included, cudaStream_t s);

// effective SM clock in MHz, one value per block (measurement utility, see kernels.cu)
int clock_probe_launch(float* mhz_out, int n_blocks, int spin_us, cudaStream_t s);

// fused non-causal attention, key-length masked (HF:500-549; torch MHA in lib/models.py:291-300)
// lse (optional, fp32 [B, heads, R]): per-row log-sum-exp in the log2 domain, for attention_bwd_launch
// drop (optional, training only): dropout on the attention weights, mask index ((b*heads + h)*R + q)*R + k
int attention_launch(const __nv_bfloat16* qkv, int B, int R, int heads, int head_dim,
                     const int32_t* kv_len, float scale, __nv_bfloat16* ctx, cudaStream_t s,
                     float* lse = nullptr, const DropSite* drop = nullptr);
// backward of the same attention (head training step, attention_bwd.cu): dqkv bf16 [B*R, 3*D] (dQ | dK | dV)
// from dctx bf16 [B*R, D], the forward's qkv / ctx and lse. delta: fp32 [B, heads, R] scratch.
int attention_bwd_launch(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                         const float* lse, float* delta, int B, int R, int heads, int head_dim,
                         const int32_t* kv_len, float scale, __nv_bfloat16* dqkv, const DropSite& drop,
                         cudaStream_t s);

// same contract, tcgen05 / TMEM / TMA implementation (attention_tc.cu) — the product path
int attention_tc_launch(const __nv_bfloat16* qkv, int B, int R, int heads, int head_dim,
                        const int32_t* kv_len, float scale, __nv_bfloat16* ctx, cudaStream_t s);
// head_dim 64 specialisation with 8 column-split softmax warps (attention_tc64.cu); reached through
// attention_tc_launch unless W2VSEG_ATT64=v15 selects the 4-warp kernel for A/B measurements
int attention_tc64_launch(const __nv_bfloat16* qkv, int B, int R, int heads, const int32_t* kv_len,
                          float scale, __nv_bfloat16* ctx, cudaStream_t s);

// ---- weight packing -----------------------------------------------------------------------
// dst_bf16[r*ld_dst + c] = src[r*cols + c] * scale
int pack_matrix_launch(const float* src, int rows, int cols, float scale, __nv_bfloat16* dst,
                       int64_t ld_dst, cudaStream_t s);
// conv weight [O, I, J] fp32 -> [O, J*I] bf16 (K index = j*I + i), optional per-tap scale[j]
int pack_conv_launch(const float* src, int O, int I, int J, const float* tap_scale,
                     __nv_bfloat16* dst, cudaStream_t s);
// ---- bias correction for bf16 weight rounding (engine.cu: w2vseg_calibrate / w2vseg_correct_bias)
// xmean[k] = mean over num_groups x rows_per_group rows of the bf16 view A (row g*group_stride + t, stride row_stride)
// (two deterministic passes through `scratch`, >= K floats; more scratch = more row slabs in parallel)
int colmean_launch(const __nv_bfloat16* A, int64_t row_stride, int K, int num_groups, int rows_per_group,
                   int64_t group_stride, float* xmean, float* scratch, size_t scratch_floats, cudaStream_t s);
// bias[n] -= sum_k (packed_bf16[n, k] - exact_fp32[n, k]) * xmean[...]  (layouts: see kernels.cu)
int bias_correct_launch(const float* src, const __nv_bfloat16* packed, int64_t ld, int N, int K, float scale,
                        int layout, int I, int J, const float* tap_scale, const float* xmean,
                        int x_tap_stride, int gc, float* bias, cudaStream_t s);
// [O, J] fp32 -> [J, O] fp32 (conv layer 0 taps)
int transpose_f32_launch(const float* src, int O, int J, float* dst, cudaStream_t s);
// dst[i] = a[i] * sa + (b ? b[i] * sb : 0)
int axpby_launch(const float* a, float sa, const float* b, float sb, float* dst, int n,
                 cudaStream_t s);
// weight-norm tap scale: scale[j] = g[j] / sqrt(sum_{o,i} v[o,i,j]^2)   (HF:355, dim=2)
int weightnorm_scale_launch(const float* v, const float* g, int OI, int J, float* scale,
                            cudaStream_t s);

// ---- talk-level reductions (see w2vseg.h) ---------------------------------------------------
int scatter_rows_launch(const float* rows, int64_t row_stride, const int32_t* start,
                        const int32_t* count, int n_rows, double* talk, int64_t n_frames,
                        int flag_col, cudaStream_t s);
int nanfill_launch(double* talk, int64_t n_frames, const int32_t* idx, int n_idx, cudaStream_t s);
int overlap_average_launch(const double* tilings, int n_tilings, int64_t n_frames, double* out,
                           cudaStream_t s);
int moving_average_launch(const double* arr, int64_t n, int window, double* out, cudaStream_t s);

}  // namespace w2v
