/*
 * w2vseg.h — C ABI of libw2vseg.so, the B200 (sm_100a) implementation of the Wav2VecSegmenter
 * segmentation-frame-classifier (SFC) forward pass over sliding audio windows.
 *
 * The reference (ahclab/Wav2VecSegmenter) is pure Python and has no FFI; the boundary this
 * library replaces is the pair of calls made per batch by lib/evaluate.py:59 and :72
 *     _, hidden = model.wav2vec_model(audio, in_mask)       (lib/models.py:367-368 / :484-485,
 *                                                            transformers Wav2Vec2Model.forward)
 *     logits    = model.seg_model(hidden, out_mask)         (lib/models.py:307-319)
 * followed by sigmoid + masking (lib/evaluate.py:82-91), the scatter of window rows into the
 * per-talk probability vector (lib/evaluate.py:100-111), the average over shifted tilings
 * (segment.py:101-108) and the trailing moving average (lib/segment.py:508-522).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C"; every pointer parameter documented "device" is a CUDA device pointer owned by
 *     the caller and must stay alive until the work queued on `stream` has completed.
 *   - No exceptions, no host allocation on the hot path; scratch comes from a caller-provided
 *     workspace sized by w2vseg_workspace_bytes(). The handle owns only its packed weights.
 *   - Every function returns 0 on success or a negative W2VSEG_ERR_* code;
 *     w2vseg_last_error() returns a thread-local description of the last failure.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - There is NO CPU fallback: without an sm_100 device every compute entry point fails.
 *
 * Geometry. A batch holds B windows. All windows of a batch share one frame stride
 *     R = w2vseg_frame_stride(l_max) = ceil(l_max / 320)
 * (l_max = longest window of the batch in samples). Frame t of window b lives at row b*R + t of
 * every [B*R, C] activation. A window of `len` samples has
 *     T(len) = w2vseg_num_frames(len)   (the 7 strided convs, HF:1005-1024)
 * frames that the encoder treats as valid keys; T(l_max) <= R - 1 always. Rows t >= T(len) are
 * still computed exactly as the reference computes its padded frames (zeroed before the
 * positional conv, HF:753-756, and kept as queries), because the head may consume one frame
 * more than the encoder mask admits (lib/evaluate.py:63-70).
 */
#ifndef W2VSEG_H_
#define W2VSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2VSEG_ABI_VERSION 4

#define W2VSEG_OK 0
#define W2VSEG_ERR_ARG (-1)    /* bad argument / shape */
#define W2VSEG_ERR_CUDA (-2)   /* CUDA runtime or driver error (incl. no sm_100 device) */
#define W2VSEG_ERR_STATE (-3)  /* weights missing / handle not finalised */

typedef struct w2vseg_handle w2vseg_handle;

/* Architecture of one SFC model: lib/models.py:173-212 (SHAS ctor) + the XLS-R-300m
 * Wav2Vec2Config the reference downloads (lib/models.py:334,443). */
typedef struct w2vseg_config {
  int32_t n_layers;          /* wav2vec_keep_layers: encoder layers kept (16 / 24)            */
  int32_t n_adapter_layers;  /* last n layers carry a ScaledParallelAdapter (0 = none)         */
  int32_t hidden;            /* 1024                                                          */
  int32_t heads;             /* 16  (head_dim must be 64)                                     */
  int32_t ffn;               /* 4096                                                          */
  int32_t adapter_dim;       /* 512 (lib/models.py:400-402)                                   */
  float adapter_scale;       /* 4.0                                                           */
  int32_t conv_dim;          /* 512                                                           */
  int32_t pos_kernel;        /* 128                                                           */
  int32_t pos_groups;        /* 16                                                            */
  int32_t head_layers;       /* n_transformer_enc_layers: 0 or 1                              */
  int32_t head_heads;        /* n_transformer_enc_heads: 8 (head_dim must be 128 or 64)       */
  int32_t head_ffn;          /* 2048 (torch.nn.TransformerEncoderLayer default)               */
  float ln_eps;              /* 1e-5                                                          */
  /* feature-extractor variant (HF Wav2Vec2Config.feat_extract_norm / conv_bias):
   *   feat_group_norm = 0: LayerNorm + GELU after each of the 7 convs (XLS-R, HF:275-299)
   *   feat_group_norm = 1: GroupNorm(512 groups, i.e. per channel over time) + GELU after conv 0,
   *                        GELU only after conv 1..6 (HF:302-323, 249-272); then fe.conv{1..6}.ln.* do not exist
   *   conv_bias = 0: the convs have no bias (fe.conv{l}.bias do not exist)                     */
  int32_t feat_group_norm;
  int32_t conv_bias;
  /* encoder layer variant (HF Wav2Vec2Config.do_stable_layer_norm):
   *   post_layer_norm = 0: pre-LN "stable layer norm" layers (XLS-R, wav2vec2-large-lv60; HF:632-655)
   *   post_layer_norm = 1: h = LN(h + Attn(h)); h = LN(h + FFN(h)) (wav2vec2-base / -large-960h; HF:575-609);
   *                        no FFN adapters (n_adapter_layers must be 0)                               */
  int32_t post_layer_norm;
} w2vseg_config;

/* ---- library ------------------------------------------------------------------------------ */
int32_t w2vseg_abi_version(void);
const char* w2vseg_last_error(void);
/* kernels launched by this library since process start (bench.py's gpu_launches) */
int64_t w2vseg_launch_count(void);
/* 0 if the current CUDA device is an sm_100 part this library can run on, else W2VSEG_ERR_CUDA */
int32_t w2vseg_device_ok(void);

/* Per-kernel device timing for the roofline report: when enabled every kernel this library
 * launches is bracketed by CUDA events on its stream. collect() synchronises the device and
 * writes "<kernel-name> <launches> <total_ms>\n" lines into buf; returns bytes written. */
int32_t w2vseg_profile_enable(int32_t on);
int64_t w2vseg_profile_collect(char* buf, size_t cap);

/* ---- geometry (pure host arithmetic) ------------------------------------------------------- */
/* conv-stack output length for `n_samples` input samples (HF:1005-1024); 0 if n_samples < 400 */
int32_t w2vseg_num_frames(int64_t n_samples);
/* frame stride R shared by all windows of a batch whose longest window has l_max samples */
int32_t w2vseg_frame_stride(int64_t l_max);

/* ---- model handle --------------------------------------------------------------------------
 * replaces: hydra.utils.instantiate(config.task.model).to(device) + load_state_dict
 * (segment.py:41-52). */
int32_t w2vseg_create(const w2vseg_config* cfg, w2vseg_handle** out);
void w2vseg_destroy(w2vseg_handle* h);

/* Upload one parameter tensor. `src` = device pointer to contiguous fp32 in the PyTorch
 * state-dict layout of that parameter; `numel` is checked against the expected size.
 * Canonical names (i = layer index, l = conv layer 0..6):
 *   fe.conv{l}.weight [512,Cin,k]  fe.conv{l}.bias  fe.conv{l}.ln.weight  fe.conv{l}.ln.bias
 *     (fe.conv0.ln.* = the GroupNorm affine when feat_group_norm = 1)
 *   fp.ln.weight  fp.ln.bias  fp.proj.weight [1024,512]  fp.proj.bias
 *   pos.weight_g [1,1,128]  pos.weight_v [1024,64,128]  (or pos.weight, already folded)  pos.bias
 *   enc.{i}.ln1.{weight,bias}  enc.{i}.{q,k,v,o}.{weight,bias}  enc.{i}.ln2.{weight,bias}
 *   enc.{i}.ff1.{weight,bias}  enc.{i}.ff2.{weight,bias}
 *   enc.{i}.ad_down.{weight,bias}  enc.{i}.ad_up.{weight,bias}          (adapter layers only)
 *   head.ln1.*  head.in_proj.{weight[3072,1024],bias}  head.o.*  head.ln2.*  head.ff1.*  head.ff2.*
 *   head.ln_f.{weight,bias}  head.out.{weight[1,1024],bias[1]}
 * The library folds weight-norm, concatenates Q/K/V, folds the adapter into the FFN matrices and
 * casts matrices to bf16 (vectors stay fp32). */
int32_t w2vseg_set_weight(w2vseg_handle* h, const char* name, const float* src_device,
                          int64_t numel, void* stream);
/* Must be called after all weights are set (checks completeness, runs the folding kernels). */
int32_t w2vseg_finalize_weights(w2vseg_handle* h, void* stream);

/* ---- bias correction for the bf16 weight rounding (optional, after finalize) -------------------
 * The matrices are stored in bf16; y = W x + b then carries the error (bf16(W) - W) x, whose mean over
 * frames is NOT zero: activations have a large frame-independent component, so every layer adds a constant
 * vector and the logits of a 24-layer model end up shifted by a frame-independent offset (+0.033 for the
 * random-init large model, profiles/parity_r02.md). Standard post-training-quantisation bias correction
 * removes it: b' = b - (bf16(W) - W) E[x], with E[x] measured on a calibration signal.
 *   1. w2vseg_calibrate: one forward over `audio` (any speech-like signal; same arguments as
 *      w2vseg_sfc_forward) that records the mean input row of every GEMM in the handle. The positional
 *      conv, whose fp32 weight-norm factors the handle keeps, is corrected here.
 *   2. w2vseg_correct_bias(name, src): once per matrix tensor, with the SAME fp32 tensor that was given to
 *      w2vseg_set_weight (names without a bf16 matrix are accepted and ignored).
 * w2vseg_finalize_weights resets every bias to its checkpoint value; correcting a matrix twice without a
 * finalize in between is an error (W2VSEG_ERR_STATE). */
int32_t w2vseg_calibrate(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                         const int32_t* sample_len, const int32_t* norm_len, const int32_t* out_len,
                         int32_t B, int64_t l_max, void* workspace, size_t workspace_bytes, void* stream);
int32_t w2vseg_correct_bias(w2vseg_handle* h, const char* name, const float* src_device, int64_t numel,
                            void* stream);

/* ---- SFC forward ---------------------------------------------------------------------------- */
/* scratch bytes needed by encode / head / sfc_forward for B windows of at most l_max samples */
size_t w2vseg_workspace_bytes(const w2vseg_handle* h, int32_t B, int64_t l_max);

/* replaces model.wav2vec_model(audio, in_mask) (lib/evaluate.py:59) INCLUDING the per-row
 * normalisation of CollateFn (lib/datautils.py:122-125).
 *   audio        device fp32 [B, audio_stride] raw (un-normalised) samples, row b valid for
 *                sample_len[b] samples; values beyond are ignored
 *   sample_len   device int32 [B]   valid samples per window (sum of in_mask row)
 *   norm_len     device int32 [B]   length the reference's mean/std were taken over = padded
 *                length of the reference batch the window belonged to (>= sample_len[b]);
 *                0 = do not normalise this row (audio already normalised, or a silent window,
 *                lib/datautils.py:88)
 *   l_max        host: max sample_len over the batch (defines R)
 *   hidden_out   device fp32 [B, R, hidden]: encoder output incl. rows >= T(len)
 *   enc_len_out  device int32 [B] or NULL: T(sample_len[b])
 *   included_out device int32 [B] or NULL: CollateFn's `included` (lib/datautils.py:88) decided
 *                on the device: 0 iff norm_len[b] > 0 and the window's samples sum to zero (such a
 *                window is not normalised; lib/evaluate.py:109-111 reports its frames as 0) */
int32_t w2vseg_encode(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                      const int32_t* sample_len, const int32_t* norm_len, int32_t B, int64_t l_max,
                      float* hidden_out, int32_t* enc_len_out, int32_t* included_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* replaces model.seg_model(hidden, out_mask) + sigmoid + masking (lib/evaluate.py:72-91).
 *   hidden        device fp32, window b frame t at hidden + b*batch_stride + t*hidden_dim
 *   T             frames per window presented to the head (hidden.shape[1])
 *   out_len       device int32 [B]: number of leading true entries of out_mask row b (<= T)
 *   logits_out    device fp32 [B, T] (0 where masked, like lib/evaluate.py:91) or NULL
 *   probs_out     device fp32 [B, T] (0 where masked) or NULL */
int32_t w2vseg_head(w2vseg_handle* h, const float* hidden, int64_t batch_stride, int32_t T,
                    const int32_t* out_len, int32_t B, float* logits_out, float* probs_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* encode + head without materialising the hidden state for the caller. Outputs are [B, R]
 * (R = w2vseg_frame_stride(l_max)); the head sees all R rows of each window as queries and
 * out_len[b] of them as keys, which yields the same valid-frame values as the reference's
 * T-row call. */
int32_t w2vseg_sfc_forward(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                           const int32_t* sample_len, const int32_t* norm_len,
                           const int32_t* out_len, int32_t B, int64_t l_max, float* logits_out,
                           float* probs_out, int32_t* included_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* w2vseg_sfc_forward writing straight into the row matrix w2vseg_scatter_rows consumes (what
 * lib/evaluate.py:82-111 keeps per window: probabilities + the `included` flag):
 *   rows_out[b * row_stride + t]        = probability of frame t (0 where masked), t < R
 *   rows_out[b * row_stride + R .. row_cols) = 0
 *   rows_out[b * row_stride + flag_col] = 1.0f / 0.0f: CollateFn's `included` (flag_col < 0: not written)
 * Requires R <= row_cols <= row_stride and flag_col in [R, row_cols) or negative. */
int32_t w2vseg_sfc_forward_rows(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                                const int32_t* sample_len, const int32_t* norm_len,
                                const int32_t* out_len, int32_t B, int64_t l_max, float* rows_out,
                                int64_t row_stride, int32_t row_cols, int32_t flag_col,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- head-only training step (frozen encoder) --------------------------------------------------
 * replaces, for the frozen-encoder setting (finetune_wav2vec=False: middle 0/16, large 0/24), the body of the
 * reference's training loop train.py:381-480 after the encoder forward: SegmentationFrameClassifier forward
 * (lib/models.py:307-319), BCEWithLogitsLoss(pos_weight, reduction none) -> masked -> .sum(dim=1).mean()
 * (train.py:416-459, ma_window unset), and the backward pass down to every seg_model parameter. The encoder
 * output comes from w2vseg_encode (no gradient flows into it). The optimiser stays with the caller: gradients are
 * returned in fp32 in the PyTorch shape of each parameter, packed into one buffer:
 *   w2vseg_head_grad_floats(h)                 total floats of the gradient buffer
 *   w2vseg_head_grad_offset(h, name, &numel)   float offset of a parameter (canonical head.* names of
 *                                              w2vseg_set_weight), -1 if unknown
 * After the optimiser step the caller re-uploads the head parameters with w2vseg_set_weight("head....");
 * that does not un-finalise the handle.
 *   hidden      device fp32, window b frame t at hidden + b*batch_stride + t*hidden_dim (as w2vseg_head)
 *   target      device fp32 [B, T] labels in [0, 1]
 *   loss_out    device fp32 [1]; logits_out device fp32 [B, T] or NULL (0 where masked)
 * Dropout (the head runs in train() mode in the reference: init_dropout on the encoder output,
 * lib/models.py:309, and the TransformerEncoderLayer's own dropout on the attention weights, after the attention
 * block, inside the FFN and after the FFN, lib/models.py:291-300): init_dropout / layer_dropout in [0, 1); masks
 * are a counter-based hash of (seed, site, element index) — csrc/dropout.cuh — regenerated in the backward, never
 * stored, so a step is reproducible from its seed and a test can rebuild the masks on the host. Pass a new seed
 * every step. With both 0 the step is the deterministic gradient of the eval-mode head.
 * Arithmetic: bf16 GEMM operands (tcgen05, fp32 accumulate) for forward, dgrad and wgrad; attention backward
 * recomputes S / P from the forward's row log-sum-exp (mma.sync, attention_bwd.cu); LayerNorm, GELU', loss and
 * all reductions in fp32 with fixed summation orders (bit-reproducible). */
int64_t w2vseg_head_grad_floats(const w2vseg_handle* h);
int64_t w2vseg_head_grad_offset(const w2vseg_handle* h, const char* name, int64_t* numel_out);
size_t w2vseg_head_train_workspace_bytes(const w2vseg_handle* h, int32_t B, int32_t T);
int32_t w2vseg_head_train_step(w2vseg_handle* h, const float* hidden, int64_t batch_stride, int32_t T,
                               const int32_t* out_len, const float* target, float pos_weight, int32_t B,
                               float* loss_out, float* logits_out, float* grads, size_t grads_floats,
                               float init_dropout, float layer_dropout, uint32_t seed,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- measurement utility ------------------------------------------------------------------------
 * Effective SM clock: n_blocks single-thread blocks spin for spin_us microseconds of %globaltimer and write
 * clock64 ticks per microsecond (MHz) to mhz_out[block] (device fp32). Enqueued right after a run of forward
 * steps it reports the clock the run was held at by the power cap, which nvidia-smi's clocks.sm does not show. */
int32_t w2vseg_clock_probe(float* mhz_out, int32_t n_blocks, int32_t spin_us, void* stream);

/* ---- talk-level reductions (all device pointers) --------------------------------------------- */
/* talk[0..n_frames) = NaN, then for each window row w: talk[start[w] .. start[w]+count[w]) =
 * (double) rows[w*row_stride .. +count[w]) ; count[w] < 0 writes zeros over -count[w] frames
 * (silent windows). If flag_col >= 0, column flag_col of each row holds the window's `included`
 * flag as a float and a zero there turns the row into a zero-writing one.
 * lib/evaluate.py:21-22,100-111. */
int32_t w2vseg_scatter_rows(const float* rows, int64_t row_stride, const int32_t* start,
                            const int32_t* count, int32_t n_rows, double* talk, int64_t n_frames,
                            int32_t flag_col, void* stream);
/* in-place sequential fill of the listed NaN frames with the nan-mean of talk[j-2 .. j+2]
 * (lib/evaluate.py:118-125); idx must be sorted ascending. */
int32_t w2vseg_nanfill(double* talk, int64_t n_frames, const int32_t* idx, int32_t n_idx,
                       void* stream);
/* out[j] = (tilings[0][j] + tilings[1][j] + ... ) / n_tilings, summed in that order in fp64
 * (segment.py:101-108). tilings = [n_tilings, n_frames] contiguous. out may alias tilings. */
int32_t w2vseg_overlap_average(const double* tilings, int32_t n_tilings, int64_t n_frames,
                               double* out, void* stream);
/* out[i] = (arr[max(0,i-window+1)] + ... + arr[i]) / count, left-to-right fp64
 * (lib/segment.py:508-522). window >= 1. out must not alias arr. */
int32_t w2vseg_moving_average(const double* arr, int64_t n, int32_t window, double* out,
                              void* stream);

/* ---- single kernels (unit tests, ncu) ------------------------------------------------------- */
/* out = epilogue(A[M,K] * W[N,K]^T): bf16 operands, fp32 accumulate on tcgen05.
 * act: 0 none, 1 erf-GELU, 2 ReLU. resid (fp32 [M,N]) requires out_f32 = 1. block_n 64/128/256
 * selects the single-CTA tile width; block_n 512 selects the CTA-pair (cta_group::2) 256x256 kernel. */
int32_t w2vseg_gemm(const void* A_bf16, const void* W_bf16, int32_t M, int32_t N, int32_t K,
                    const float* bias, int32_t act, const float* resid, void* out,
                    int32_t out_f32, int32_t block_n, void* stream);
/* Strided conv as implicit GEMM: x channels-last bf16 [rows_in, C], out bf16 [rows_out, N],
 * out[r] = bias + W[N, kw*C] . x[r*stride .. r*stride+kw) ; x must have kw extra rows of slack. */
int32_t w2vseg_conv_gemm(const void* x_bf16, int64_t rows_out, int32_t C, int32_t kw,
                         int32_t stride, const void* W_bf16, int32_t N, const float* bias,
                         void* out_bf16, void* stream);
/* Grouped positional convolution (HF:326-379 + HF:764-765) on a zero-padded channels-last input:
 *   h[b*R + t, o] += gelu(bias[o] + sum_{j<taps, i<64} W[o, j*64 + i] * zpad[b*(R+2*halo) + t + j, (o/64)*64 + i])
 * zpad bf16 [B*(R+2*halo)+2*halo, D] (halo = taps/2 zero rows around every window), W bf16 [D, taps*64],
 * h fp32 [B*R, D] updated in place. impl 0: the forward pass's resident-A kernel (posconv_tc.cu),
 * impl 1: the generic shifted-row GEMM (gemm_tc.cu, a_mode 1) kept as a second implementation. */
int32_t w2vseg_posconv(const void* zpad_bf16, const void* W_bf16, const float* bias, int32_t B,
                       int32_t R, int32_t D, int32_t taps, float* h, int32_t impl, void* stream);
/* Conv layer 0 (Conv1d(1, 512, k=10, stride=5, bias) -> LayerNorm(512) -> GELU, HF:281-299) with the
 * window normalisation (x - mean) * rstd applied on the fly (lib/datautils.py:122-125).
 * audio fp32 [B, audio_stride]; samples at or beyond sample_len[b] read as 0; stats fp32 [B][2] =
 * (mean, 1/std) per window; w fp32 [512, 10]; out bf16 [B*R0, 512] channels-last (frame t of window b
 * in row b*R0 + t). impl 0: the forward pass's tcgen05 kernel (LayerNorm folded into a K=16 fp16 MMA,
 * conv0_tc.cu); impl 1: the CUDA-core kernel kept as a second implementation. scratch: >= 64 KiB of
 * device memory for the packed weights. */
int32_t w2vseg_conv0(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                     const float* stats, const float* w, const float* bias, const float* gamma,
                     const float* beta, float eps, void* out_bf16, int32_t B, int32_t R0,
                     int32_t impl, void* scratch, size_t scratch_bytes, void* stream);
/* LayerNorm over the last dim (C = 512 or 1024). in: fp32 or bf16; out: bf16; act 0/1 (GELU). */
int32_t w2vseg_layernorm(const void* in, int32_t in_f32, int64_t rows, int32_t C,
                         const float* gamma, const float* beta, float eps, int32_t act,
                         void* out_bf16, void* stream);
/* Non-causal multi-head attention with a per-window key-length mask.
 * qkv bf16 [B*R, 3*heads*head_dim] (Q | K | V column blocks); ctx bf16 [B*R, heads*head_dim]. */
int32_t w2vseg_attention(const void* qkv_bf16, int32_t B, int32_t R, int32_t heads,
                         int32_t head_dim, const int32_t* kv_len, float scale, void* ctx_bf16,
                         void* stream);
/* same contract on the legacy warp-level mma.sync path: kept as an independent second
 * implementation for the parity tests of the tcgen05 kernel, not used by the forward pass */
int32_t w2vseg_attention_mma(const void* qkv_bf16, int32_t B, int32_t R, int32_t heads,
                             int32_t head_dim, const int32_t* kv_len, float scale, void* ctx_bf16,
                             void* stream);

/* training forward of the head attention: w2vseg_attention_mma that also returns the per-row log-sum-exp
 * (fp32 [B, heads, R], log2 domain) the backward re-exponentiates with. dropout in [0, 1) on the attention
 * weights: element (b, h, q, k) is kept iff lowbias32((((b*heads + h)*R + q)*R + k) ^ key) >= dropout * 2^32,
 * key = lowbias32(seed * 0x9E3779B9 + 1) (site 1 of csrc/dropout.cuh), and scaled by 1 / (1 - dropout). */
int32_t w2vseg_attention_train(const void* qkv_bf16, int32_t B, int32_t R, int32_t heads, int32_t head_dim,
                               const int32_t* kv_len, float scale, void* ctx_bf16, float* lse, float dropout,
                               uint32_t seed, void* stream);
/* attention backward: dqkv bf16 [B*R, 3*heads*head_dim] (dQ | dK | dV) from dctx bf16 [B*R, heads*head_dim],
 * the forward's qkv / ctx / lse; delta_scratch: fp32 [B, heads, R]. Rows >= kv_len[b] get zero dK / dV.
 * dropout / seed: the forward's. */
int32_t w2vseg_attention_bwd(const void* qkv_bf16, const void* ctx_bf16, const void* dctx_bf16, const float* lse,
                             float* delta_scratch, int32_t B, int32_t R, int32_t heads, int32_t head_dim,
                             const int32_t* kv_len, float scale, void* dqkv_bf16, float dropout, uint32_t seed,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2VSEG_H_ */
