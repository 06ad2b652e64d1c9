// Model handle, weight packing and the forward-pass orchestration behind the C ABI (w2vseg.h).
//
// One handle = one SFC model replica on the current device (one process per GPU). The forward
// pass is a fixed sequence of launches on the caller's stream; it allocates nothing and performs
// no host synchronisation, so the host can capture it into a CUDA graph or run it ahead.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "train_kernels.cuh"

using namespace w2v;
typedef __nv_bfloat16 bf16;

namespace {

constexpr int kConvK[7] = {10, 3, 3, 3, 3, 2, 2};
constexpr int kHalo = 64;  // pos-conv padding (128 // 2, HF:343)
constexpr size_t kCalibScratch = 64 * 8192;  // floats: row-slab partial sums of the calibration pass

struct LNW { float* g = nullptr; float* b = nullptr; };

// mean input (over the rows of the calibration batch) of every GEMM of a transformer layer
struct LayerX { float* ln1 = nullptr; float* ctx = nullptr; float* ln2 = nullptr; float* mid = nullptr; };

struct EncLayerW {
  LNW ln1, ln2;
  LayerX x;
  bf16* wqkv = nullptr; float* bqkv = nullptr;
  bf16* wo = nullptr;   float* bo = nullptr;
  bf16* w1 = nullptr;   float* b1 = nullptr;   // [F1, D]: FFN-up rows, then adapter-down rows
  bf16* w2 = nullptr;                          // [D, F1]: FFN-down cols, then scale*adapter-up cols
  float* b2_raw = nullptr; float* bu_raw = nullptr; float* b2 = nullptr;
  int F1 = 0;
  bool adapter = false;
};

struct HeadW {
  LNW ln1, ln2, lnf;
  LayerX x;
  bf16* win = nullptr; float* bin = nullptr;
  bf16* wo = nullptr;  float* bo = nullptr;
  bf16* w1 = nullptr;  float* b1 = nullptr;
  bf16* w2 = nullptr;  float* b2 = nullptr;
  float* wout = nullptr; float* bout = nullptr;
};

enum SlotKind { SLOT_VEC, SLOT_MAT, SLOT_CONV, SLOT_CONV0, SLOT_RAW };

struct Slot {
  SlotKind kind;
  int64_t numel;
  // VEC / RAW
  float* fdst = nullptr;
  float scale = 1.f;
  // MAT
  bf16* bdst = nullptr;
  int rows = 0, cols = 0;
  int64_t ld = 0;
  // CONV [O, I, J]
  int O = 0, I = 0, J = 0;
  bool optional = false;
  int alt_group = 0;  // slots sharing a non-zero alt_group: group satisfied by its 'primary' set
  // bias correction (w2vseg_correct_bias): effective bias of this matrix's output rows, mean input of its
  // columns (x_tap_stride > 0: grouped positional conv table [tap][channel])
  float* c_bias = nullptr;
  const float* c_x = nullptr;
  int c_x_tap_stride = 0;
};

struct BiasPair { float* raw; float* eff; int n; };

struct Workspace {
  float2* stats; int32_t* enc_len; int32_t* included; double2* stat_partial;
  bf16* conv[7];
  bf16* feat; float* h; bf16* zpad; bf16* xn; bf16* qkv; bf16* ctx; bf16* mid;
  float* gn_scratch;   // GroupNorm extractor only: per-block channel sums + per-window folded taps
  size_t bytes;
};

}  // namespace

struct w2vseg_handle {
  w2vseg_config cfg;
  int D, DH, F1max;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0, arena_used = 0;
  // weights
  float* conv0_wt = nullptr;
  uint8_t* conv0_pack = nullptr;   // fp16 LayerNorm-folded weights + variance factor (conv0_tc.cu)
  bool conv0_cuda_cores = false;   // W2VSEG_CONV0=cuda: the CUDA-core kernel (A/B measurements)
  float* conv_b[7] = {};
  LNW conv_ln[7];
  bf16* conv_w[7] = {};
  LNW fp_ln; bf16* fp_w = nullptr; float* fp_b = nullptr;
  float* pos_g = nullptr; float* pos_v = nullptr; float* pos_scale = nullptr;
  bf16* pos_w = nullptr; float* pos_b = nullptr;
  std::vector<EncLayerW> enc;
  HeadW head;
  std::map<std::string, Slot> slots;
  std::map<std::string, bool> was_set;
  bool finalized = false;
  // bias correction state: GEMM biases exist as (raw, effective) pairs; finalize() resets eff = raw
  std::vector<BiasPair> bias_pairs;
  std::map<std::string, bool> corrected;
  bool calibrated = false;
  float* conv_x[7] = {};           // mean im2col row of conv layer l (k*512 floats)
  float* fp_x = nullptr;           // mean input of the feature projection (512)
  float* pos_x = nullptr;          // [taps][D]: mean of zpad rows shifted by tap
  float* calib_scratch = nullptr;  // kCalibScratch floats

  template <typename T>
  T* alloc(size_t n) {
    size_t off = (arena_used + 255) & ~(size_t)255;
    arena_used = off + n * sizeof(T);
    if (arena == nullptr) return nullptr;  // sizing pass
    return reinterpret_cast<T*>(arena + off);
  }
};

namespace {

void add_vec(w2vseg_handle* h, const std::string& name, float** dst, int n, bool optional = false) {
  *dst = h->alloc<float>(n);
  Slot s; s.kind = SLOT_VEC; s.numel = n; s.fdst = *dst; s.optional = optional;
  h->slots[name] = s;
}
// GEMM bias: `raw` receives the checkpoint values, the forward reads `eff` (= raw minus bias corrections)
float* add_bias(w2vseg_handle* h, const std::string& name, float** eff, int n) {
  float* raw = h->alloc<float>(n);
  *eff = h->alloc<float>(n);
  if (!name.empty()) {
    Slot s; s.kind = SLOT_VEC; s.numel = n; s.fdst = raw;
    h->slots[name] = s;
  }
  h->bias_pairs.push_back({raw, *eff, n});
  return raw;
}
void add_ln(w2vseg_handle* h, const std::string& prefix, LNW* ln, int n) {
  add_vec(h, prefix + ".weight", &ln->g, n);
  add_vec(h, prefix + ".bias", &ln->b, n);
}
// matrix [rows, cols] packed into dst (+col offset / row offset already applied), leading dim ld
void add_mat(w2vseg_handle* h, const std::string& name, bf16* dst, int rows, int cols, int64_t ld,
             float scale = 1.f, float* c_bias = nullptr, const float* c_x = nullptr) {
  Slot s; s.kind = SLOT_MAT; s.numel = (int64_t)rows * cols; s.bdst = dst; s.rows = rows;
  s.cols = cols; s.ld = ld; s.scale = scale; s.c_bias = c_bias; s.c_x = c_x;
  h->slots[name] = s;
}

// Lays out every parameter in the arena and registers its upload slot. Called twice: once with
// arena == nullptr to size the arena, once for real.
void build_layout(w2vseg_handle* h) {
  const w2vseg_config& c = h->cfg;
  const int D = h->D, CD = c.conv_dim;
  h->arena_used = 0;
  h->slots.clear();
  h->bias_pairs.clear();

  h->calib_scratch = h->alloc<float>(kCalibScratch);
  // feature extractor
  h->conv0_wt = h->alloc<float>((size_t)10 * CD);
  h->conv0_pack = h->alloc<uint8_t>(conv0_tc_pack_bytes());
  {
    Slot s; s.kind = SLOT_CONV0; s.numel = (int64_t)CD * 10; s.fdst = h->conv0_wt; s.O = CD; s.J = 10;
    h->slots["fe.conv0.weight"] = s;
  }
  for (int l = 0; l < 7; ++l) {
    const std::string p = "fe.conv" + std::to_string(l);
    const bool gn = c.feat_group_norm != 0;
    if (!c.conv_bias) {                  // no conv biases in this architecture: a zero vector (arena is zero-filled)
      if (l == 0) h->conv_b[l] = h->alloc<float>(CD);
      else add_bias(h, "", &h->conv_b[l], CD);
    } else if (l == 0) {
      add_vec(h, p + ".bias", &h->conv_b[l], CD);   // folded into the fp16 pack of conv0_tc / the GroupNorm taps
    } else {
      add_bias(h, p + ".bias", &h->conv_b[l], CD);
    }
    if (!gn || l == 0) add_ln(h, p + ".ln", &h->conv_ln[l], CD);   // "group": GroupNorm affine after conv 0 only
    if (l > 0) {
      h->conv_w[l] = h->alloc<bf16>((size_t)CD * CD * kConvK[l]);
      h->conv_x[l] = h->alloc<float>((size_t)CD * kConvK[l]);
      Slot s; s.kind = SLOT_CONV; s.numel = (int64_t)CD * CD * kConvK[l]; s.bdst = h->conv_w[l];
      s.O = CD; s.I = CD; s.J = kConvK[l];
      // Bias correction needs E[x] to carry over from the calibration signal to the data. After a per-frame
      // LayerNorm it does; in the GroupNorm variant the conv inputs are only normalised per channel over TIME,
      // their mean follows the signal's loudness profile, and a calibrated correction does more harm than good
      // (measured: hidden-state error 0.7 % -> 2.4 %): conv layers are left uncorrected there.
      if (!gn) { s.c_bias = h->conv_b[l]; s.c_x = h->conv_x[l]; }
      h->slots[p + ".weight"] = s;
    }
  }
  // feature projection
  add_ln(h, "fp.ln", &h->fp_ln, CD);
  h->fp_w = h->alloc<bf16>((size_t)D * CD);
  h->fp_x = h->alloc<float>(CD);
  add_bias(h, "fp.proj.bias", &h->fp_b, D);
  add_mat(h, "fp.proj.weight", h->fp_w, D, CD, CD, 1.f, h->fp_b, h->fp_x);
  // positional conv: either (weight_g, weight_v) or an already folded weight
  const int gc = D / c.pos_groups;  // channels per group
  h->pos_g = h->alloc<float>(c.pos_kernel);
  h->pos_v = h->alloc<float>((size_t)D * gc * c.pos_kernel);
  h->pos_scale = h->alloc<float>(c.pos_kernel);
  h->pos_w = h->alloc<bf16>((size_t)D * gc * c.pos_kernel);
  h->pos_x = h->alloc<float>((size_t)c.pos_kernel * D);
  add_bias(h, "pos.bias", &h->pos_b, D);
  {
    Slot s; s.kind = SLOT_RAW; s.numel = c.pos_kernel; s.fdst = h->pos_g; s.alt_group = 1;
    h->slots["pos.weight_g"] = s;
    Slot v; v.kind = SLOT_RAW; v.numel = (int64_t)D * gc * c.pos_kernel; v.fdst = h->pos_v;
    v.alt_group = 1;
    h->slots["pos.weight_v"] = v;
    Slot w; w.kind = SLOT_CONV; w.numel = v.numel; w.bdst = h->pos_w; w.O = D; w.I = gc;
    w.J = c.pos_kernel; w.alt_group = 2; w.c_bias = h->pos_b; w.c_x = h->pos_x; w.c_x_tap_stride = D;
    h->slots["pos.weight"] = w;
  }

  // encoder layers
  h->enc.assign(c.n_layers, EncLayerW());
  for (int i = 0; i < c.n_layers; ++i) {
    EncLayerW& L = h->enc[i];
    const std::string p = "enc." + std::to_string(i);
    L.adapter = i >= c.n_layers - c.n_adapter_layers;
    L.F1 = c.ffn + (L.adapter ? c.adapter_dim : 0);
    add_ln(h, p + ".ln1", &L.ln1, D);
    add_ln(h, p + ".ln2", &L.ln2, D);
    L.x.ln1 = h->alloc<float>(D); L.x.ctx = h->alloc<float>(D); L.x.ln2 = h->alloc<float>(D);
    L.x.mid = h->alloc<float>(L.F1);
    L.wqkv = h->alloc<bf16>((size_t)3 * D * D);
    float* bqkv_raw = add_bias(h, "", &L.bqkv, 3 * D);
    const char* qkvn[3] = {".q", ".k", ".v"};
    // post-LN encoder, layer 0: q/k/v read the un-normalised output of the positional conv (no per-frame
    // LayerNorm in front of them): a calibrated mean does not carry over, no bias correction (cf. the GroupNorm convs)
    const bool qkv_corr = !(c.post_layer_norm && i == 0);
    for (int j = 0; j < 3; ++j) {
      add_mat(h, p + qkvn[j] + ".weight", L.wqkv + (size_t)j * D * D, D, D, D, 1.f,
              qkv_corr ? L.bqkv + (size_t)j * D : nullptr, L.x.ln1);
      Slot s; s.kind = SLOT_VEC; s.numel = D; s.fdst = bqkv_raw + (size_t)j * D;
      h->slots[p + qkvn[j] + ".bias"] = s;
    }
    L.wo = h->alloc<bf16>((size_t)D * D);
    add_bias(h, p + ".o.bias", &L.bo, D);
    add_mat(h, p + ".o.weight", L.wo, D, D, D, 1.f, L.bo, L.x.ctx);
    L.w1 = h->alloc<bf16>((size_t)L.F1 * D);
    float* b1_raw = add_bias(h, "", &L.b1, L.F1);
    L.w2 = h->alloc<bf16>((size_t)D * L.F1);
    L.b2_raw = h->alloc<float>(D);
    L.bu_raw = h->alloc<float>(D);
    L.b2 = h->alloc<float>(D);
    add_mat(h, p + ".ff1.weight", L.w1, c.ffn, D, D, 1.f, L.b1, L.x.ln2);
    { Slot s; s.kind = SLOT_VEC; s.numel = c.ffn; s.fdst = b1_raw; h->slots[p + ".ff1.bias"] = s; }
    add_mat(h, p + ".ff2.weight", L.w2, D, c.ffn, L.F1, 1.f, L.b2, L.x.mid);
    { Slot s; s.kind = SLOT_VEC; s.numel = D; s.fdst = L.b2_raw; h->slots[p + ".ff2.bias"] = s; }
    if (L.adapter) {
      // y + s*(Wu relu(Wd u + bd) + bu)  ==  extra FFN hidden units with ReLU and weights s*Wu
      add_mat(h, p + ".ad_down.weight", L.w1 + (size_t)c.ffn * D, c.adapter_dim, D, D, 1.f, L.b1 + c.ffn, L.x.ln2);
      { Slot s; s.kind = SLOT_VEC; s.numel = c.adapter_dim; s.fdst = b1_raw + c.ffn;
        h->slots[p + ".ad_down.bias"] = s; }
      add_mat(h, p + ".ad_up.weight", L.w2 + c.ffn, D, c.adapter_dim, L.F1, c.adapter_scale, L.b2, L.x.mid + c.ffn);
      { Slot s; s.kind = SLOT_VEC; s.numel = D; s.fdst = L.bu_raw; h->slots[p + ".ad_up.bias"] = s; }
    }
  }

  // head
  if (c.head_layers > 0) {
    HeadW& H = h->head;
    add_ln(h, "head.ln1", &H.ln1, D);
    add_ln(h, "head.ln2", &H.ln2, D);
    H.x.ln1 = h->alloc<float>(D); H.x.ctx = h->alloc<float>(D); H.x.ln2 = h->alloc<float>(D);
    H.x.mid = h->alloc<float>(c.head_ffn);
    H.win = h->alloc<bf16>((size_t)3 * D * D);
    add_bias(h, "head.in_proj.bias", &H.bin, 3 * D);
    add_mat(h, "head.in_proj.weight", H.win, 3 * D, D, D, 1.f, H.bin, H.x.ln1);
    H.wo = h->alloc<bf16>((size_t)D * D);
    add_bias(h, "head.o.bias", &H.bo, D);
    add_mat(h, "head.o.weight", H.wo, D, D, D, 1.f, H.bo, H.x.ctx);
    H.w1 = h->alloc<bf16>((size_t)c.head_ffn * D);
    add_bias(h, "head.ff1.bias", &H.b1, c.head_ffn);
    add_mat(h, "head.ff1.weight", H.w1, c.head_ffn, D, D, 1.f, H.b1, H.x.ln2);
    H.w2 = h->alloc<bf16>((size_t)D * c.head_ffn);
    add_bias(h, "head.ff2.bias", &H.b2, D);
    add_mat(h, "head.ff2.weight", H.w2, D, c.head_ffn, c.head_ffn, 1.f, H.b2, H.x.mid);
  }
  add_ln(h, "head.ln_f", &h->head.lnf, D);
  add_vec(h, "head.out.weight", &h->head.wout, D);
  add_vec(h, "head.out.bias", &h->head.bout, 1);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves the workspace; with base == nullptr only computes the size.
Workspace carve(const w2vseg_handle* h, uint8_t* base, int B, int R) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    off = align_up(off, 1024);
    uint8_t* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const size_t M = (size_t)B * R;
  const int D = h->D, CD = h->cfg.conv_dim;
  w.stats = reinterpret_cast<float2*>(take(sizeof(float2) * B));
  w.enc_len = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * B));
  w.included = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * B));
  w.stat_partial = reinterpret_cast<double2*>(take(sizeof(double2) * 64 * B));
  for (int l = 0; l < 7; ++l) {
    const size_t rows = M << (6 - l);
    w.conv[l] = reinterpret_cast<bf16*>(take((rows + 4) * CD * sizeof(bf16)));
  }
  w.feat = reinterpret_cast<bf16*>(take(M * CD * sizeof(bf16)));
  w.h = reinterpret_cast<float*>(take(M * D * sizeof(float)));
  w.zpad = reinterpret_cast<bf16*>(take(((size_t)B * (R + 2 * kHalo) + 2 * kHalo) * D * sizeof(bf16)));
  w.xn = reinterpret_cast<bf16*>(take(M * D * sizeof(bf16)));
  w.qkv = reinterpret_cast<bf16*>(take(M * 3 * D * sizeof(bf16)));
  w.ctx = reinterpret_cast<bf16*>(take(M * D * sizeof(bf16)));
  w.mid = reinterpret_cast<bf16*>(take(M * (size_t)h->F1max * sizeof(bf16)));
  w.gn_scratch = h->cfg.feat_group_norm
                     ? reinterpret_cast<float*>(take(conv0_gn_scratch_floats(B, R << 6) * sizeof(float)))
                     : nullptr;
  w.bytes = align_up(off, 1024);
  return w;
}

GemmProblem linear(const bf16* A, int64_t M, int K, const bf16* W, int N, const float* bias) {
  GemmProblem g = {};
  g.A = A; g.a_rows = M; g.a_row_stride = K; g.a_cols = K;
  g.W = W; g.N = N; g.K = K;
  g.num_groups = 1; g.rows_per_group = (int)M; g.a_group_rows = 0; g.o_group_rows = 0;
  g.a_mode = 0;
  g.bias = bias;
  g.act_split = N; g.act_lo = ACT_NONE; g.act_hi = ACT_NONE;
  g.resid = nullptr; g.ld_resid = 0;
  g.out = nullptr; g.ld_out = N; g.out_f32 = 0;
  g.mask_len = nullptr; g.mask_period = 1;
  return g;
}

// grouped positional conv over the zero-padded input (halo = taps/2 rows around every window), h += gelu(..)
static GemmProblem posconv_problem(const bf16* zpad, const bf16* W, const float* bias, int B, int R, int D,
                                   int taps, float* h) {
  const int halo = taps / 2;
  GemmProblem g = {};
  g.A = zpad; g.a_rows = (int64_t)B * (R + 2 * halo) + 2 * halo; g.a_row_stride = D; g.a_cols = D;
  g.W = W; g.N = D; g.K = 64 * taps;
  g.num_groups = B; g.rows_per_group = R; g.a_group_rows = R + 2 * halo; g.o_group_rows = R;
  g.a_mode = 1;
  g.bias = bias; g.act_split = D; g.act_lo = ACT_GELU; g.act_hi = ACT_GELU;
  g.resid = h; g.ld_resid = D; g.out = h; g.ld_out = D; g.out_f32 = 1;
  g.mask_len = nullptr; g.mask_period = 1;
  return g;
}

int check_ready(const w2vseg_handle* h) {
  if (h == nullptr) { set_error("null handle"); return W2VSEG_ERR_ARG; }
  if (!h->finalized) {
    set_error("weights are not finalised: call w2vseg_set_weight for every tensor, then "
              "w2vseg_finalize_weights");
    return W2VSEG_ERR_STATE;
  }
  return 0;
}

// ---- encoder: audio -> h (fp32 [B*R, D]) -------------------------------------------------------
// calib: additionally record the mean input row of every GEMM (w2vseg_calibrate)
int run_encoder(w2vseg_handle* h, const Workspace& w, const float* audio, int64_t audio_stride,
                const int32_t* sample_len, const int32_t* norm_len, int B, int R, cudaStream_t st,
                bool calib = false, int64_t l_max_samples = 0) {
  const w2vseg_config& c = h->cfg;
  const int D = h->D, CD = c.conv_dim;
  const int64_t M = (int64_t)B * R;

  W2V_TRY(window_stats_launch(audio, audio_stride, sample_len, norm_len, B, w.stat_partial, w.stats,
                              w.enc_len, w.included, st));

  // conv feature extractor (HF:382-419). Layer l activations: channels-last bf16 [B*R*2^(6-l), 512].
  const int R0 = R << 6;
  const bool gn = c.feat_group_norm != 0;
  if (gn)   // GroupNorm over time: statistics over the padded row of the reference batch (norm_len)
    W2V_TRY(conv0_gn_gelu_launch(audio, audio_stride, sample_len, norm_len, (int)l_max_samples, w.stats, h->conv0_wt,
                                 c.conv_bias ? h->conv_b[0] : nullptr, h->conv_ln[0].g, h->conv_ln[0].b, c.ln_eps,
                                 w.gn_scratch, w.conv[0], B, R0, st));
  else if (h->conv0_cuda_cores)
    W2V_TRY(conv0_ln_gelu_launch(audio, audio_stride, sample_len, w.stats, h->conv0_wt, h->conv_b[0],
                                 h->conv_ln[0].g, h->conv_ln[0].b, c.ln_eps, w.conv[0], B, R0, st));
  else
    W2V_TRY(conv0_tc_launch(audio, audio_stride, sample_len, w.stats, h->conv0_pack, c.ln_eps,
                            w.conv[0], B, R0, st));
  for (int l = 1; l < 7; ++l) {
    const int64_t rows_in = M << (7 - l), rows_out = M << (6 - l);
    // the last output row's im2col view runs (k-2) rows past the input: keep that slack finite
    W2V_CHECK_CUDA(cudaMemsetAsync(w.conv[l - 1] + rows_in * CD, 0, (size_t)4 * CD * sizeof(bf16), st));
    GemmProblem g = linear(w.conv[l - 1], rows_out, kConvK[l] * CD, h->conv_w[l], CD, h->conv_b[l]);
    g.a_row_stride = 2 * CD;  // stride-2 conv: consecutive output frames start 2 input rows apart
    g.out = w.conv[l]; g.ld_out = CD;
    if (gn) { g.act_lo = ACT_GELU; g.act_hi = ACT_GELU; }   // no norm after conv 1..6: GELU in the GEMM epilogue
    if (calib) W2V_TRY(colmean_launch(w.conv[l - 1], 2 * CD, kConvK[l] * CD, 1, (int)rows_out, 0, h->conv_x[l], h->calib_scratch, kCalibScratch, st));
    prof_tag(kConvK[l] == 3 ? "gemm.conv_k3" : "gemm.conv_k2");
    W2V_TRY(gemm_tc2_launch(g, st));
    if (!gn) {
      prof_tag("ln_gelu.conv");
      W2V_TRY(layernorm_launch(w.conv[l], false, rows_out, CD, h->conv_ln[l].g, h->conv_ln[l].b,
                               c.ln_eps, /*gelu*/ 1, w.conv[l], st));
    }
  }

  // feature projection (HF:429-434) with the frame mask fused (HF:753-756)
  W2V_TRY(layernorm_launch(w.conv[6], false, M, CD, h->fp_ln.g, h->fp_ln.b, c.ln_eps, 0, w.feat, st));
  {
    GemmProblem g = linear(w.feat, M, CD, h->fp_w, D, h->fp_b);
    g.out = w.h; g.ld_out = D; g.out_f32 = 1;
    g.mask_len = w.enc_len; g.mask_period = R;
    if (calib) W2V_TRY(colmean_launch(w.feat, CD, CD, 1, (int)M, 0, h->fp_x, h->calib_scratch, kCalibScratch, st));
    prof_tag("gemm.feat_proj");
    W2V_TRY(gemm_tc2_launch(g, st));
  }

  // positional conv embedding (HF:360-368) + residual (HF:764-765)
  W2V_CHECK_CUDA(cudaMemsetAsync(w.zpad, 0, ((size_t)B * (R + 2 * kHalo) + 2 * kHalo) * D * sizeof(bf16), st));
  W2V_TRY(cast_to_padded_launch(w.h, B, R, D, kHalo, w.zpad, st));
  {
    const int gc = D / c.pos_groups;
    W2V_REQUIRE(gc == 64, "positional conv: %d channels per group unsupported (64 only)", gc);
    GemmProblem g = posconv_problem(w.zpad, h->pos_w, h->pos_b, B, R, D, c.pos_kernel, w.h);
    if (calib)   // tap j reads the rows shifted by j: one mean row per tap
      for (int j = 0; j < c.pos_kernel; ++j)
        W2V_TRY(colmean_launch(w.zpad + (size_t)j * D, D, D, B, R, R + 2 * kHalo, h->pos_x + (size_t)j * D, h->calib_scratch, kCalibScratch, st));
    prof_tag("gemm.pos_conv");
    W2V_TRY(posconv_tc_launch(g, st));
  }

  // transformer layers (pre-LN "stable layer norm" variant, HF:632-655; adapter lib/models.py:404-428)
  const bool post_ln = c.post_layer_norm != 0;
  // post-LN encoder (do_stable_layer_norm = False, HF Wav2Vec2EncoderLayer): h = LN1(h + Attn(h)); h = LN2(h + FFN(h)).
  // The attention / FFN inputs are the residual stream itself, so every LayerNorm writes fp32 (in place) AND the
  // bf16 copy the next GEMM reads; only the first layer needs a separate cast (encoder.layer_norm is an Identity in
  // the reference, lib/models.py:349).
  if (post_ln && c.n_layers > 0) W2V_TRY(cast_to_padded_launch(w.h, B, R, D, 0, w.xn, st));
  for (int i = 0; i < c.n_layers; ++i) {
    const EncLayerW& L = h->enc[i];
    if (!post_ln) W2V_TRY(layernorm_launch(w.h, true, M, D, L.ln1.g, L.ln1.b, c.ln_eps, 0, w.xn, st));
    {
      GemmProblem g = linear(w.xn, M, D, L.wqkv, 3 * D, L.bqkv);
      g.out = w.qkv; g.ld_out = 3 * D;
      if (calib) W2V_TRY(colmean_launch(w.xn, D, D, 1, (int)M, 0, L.x.ln1, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.qkv");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    W2V_TRY(attention_tc_launch(w.qkv, B, R, c.heads, h->DH, w.enc_len, 1.0f / sqrtf((float)h->DH),
                             w.ctx, st));
    {
      GemmProblem g = linear(w.ctx, M, D, L.wo, D, L.bo);
      g.resid = w.h; g.ld_resid = D; g.out = w.h; g.ld_out = D; g.out_f32 = 1;
      if (calib) W2V_TRY(colmean_launch(w.ctx, D, D, 1, (int)M, 0, L.x.ctx, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.attn_out");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    if (post_ln) W2V_TRY(layernorm_dual_launch(w.h, M, L.ln1.g, L.ln1.b, c.ln_eps, w.xn, st));
    else W2V_TRY(layernorm_launch(w.h, true, M, D, L.ln2.g, L.ln2.b, c.ln_eps, 0, w.xn, st));
    {
      GemmProblem g = linear(w.xn, M, D, L.w1, L.F1, L.b1);
      g.act_split = c.ffn; g.act_lo = ACT_GELU; g.act_hi = ACT_RELU;
      g.out = w.mid; g.ld_out = L.F1;
      if (calib) W2V_TRY(colmean_launch(w.xn, D, D, 1, (int)M, 0, L.x.ln2, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.ffn_up");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    {
      GemmProblem g = linear(w.mid, M, L.F1, L.w2, D, L.b2);
      g.resid = w.h; g.ld_resid = D; g.out = w.h; g.ld_out = D; g.out_f32 = 1;
      if (calib) W2V_TRY(colmean_launch(w.mid, L.F1, L.F1, 1, (int)M, 0, L.x.mid, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.ffn_down");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    if (post_ln) W2V_TRY(layernorm_dual_launch(w.h, M, L.ln2.g, L.ln2.b, c.ln_eps, w.xn, st));
  }
  return 0;
}

// ---- head: y (fp32 [B*R, D], updated in place) -> logits / probs --------------------------------
int run_head(w2vseg_handle* h, const Workspace& w, float* y, int B, int R, const int32_t* out_len,
             float* logits, float* probs, cudaStream_t st, int64_t prob_stride = 0, int row_cols = 0,
             int flag_col = -1, bool calib = false) {
  const w2vseg_config& c = h->cfg;
  const int D = h->D;
  const int64_t M = (int64_t)B * R;
  if (c.head_layers > 0) {
    const HeadW& H = h->head;
    const int hd = D / c.head_heads;
    W2V_TRY(layernorm_launch(y, true, M, D, H.ln1.g, H.ln1.b, c.ln_eps, 0, w.xn, st));
    {
      GemmProblem g = linear(w.xn, M, D, H.win, 3 * D, H.bin);
      g.out = w.qkv; g.ld_out = 3 * D;
      if (calib) W2V_TRY(colmean_launch(w.xn, D, D, 1, (int)M, 0, H.x.ln1, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.head");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    W2V_TRY(attention_tc_launch(w.qkv, B, R, c.head_heads, hd, out_len, 1.0f / sqrtf((float)hd), w.ctx, st));
    {
      GemmProblem g = linear(w.ctx, M, D, H.wo, D, H.bo);
      g.resid = y; g.ld_resid = D; g.out = y; g.ld_out = D; g.out_f32 = 1;
      if (calib) W2V_TRY(colmean_launch(w.ctx, D, D, 1, (int)M, 0, H.x.ctx, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.head");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    W2V_TRY(layernorm_launch(y, true, M, D, H.ln2.g, H.ln2.b, c.ln_eps, 0, w.xn, st));
    {
      GemmProblem g = linear(w.xn, M, D, H.w1, c.head_ffn, H.b1);
      g.act_lo = ACT_GELU; g.act_hi = ACT_GELU;
      g.out = w.mid; g.ld_out = c.head_ffn;
      if (calib) W2V_TRY(colmean_launch(w.xn, D, D, 1, (int)M, 0, H.x.ln2, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.head");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
    {
      GemmProblem g = linear(w.mid, M, c.head_ffn, H.w2, D, H.b2);
      g.resid = y; g.ld_resid = D; g.out = y; g.ld_out = D; g.out_f32 = 1;
      if (calib) W2V_TRY(colmean_launch(w.mid, c.head_ffn, c.head_ffn, 1, (int)M, 0, H.x.mid, h->calib_scratch, kCalibScratch, st));
      prof_tag("gemm.head");
      W2V_TRY(gemm_tc2_launch(g, st));
    }
  }
  W2V_TRY(head_final_launch(y, B, R, D, h->head.lnf.g, h->head.lnf.b, c.ln_eps, h->head.wout,
                            h->head.bout, out_len, logits, probs, prob_stride > 0 ? prob_stride : R,
                            row_cols, flag_col, w.included, st));
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int32_t w2vseg_create(const w2vseg_config* cfg, w2vseg_handle** out) {
  W2V_REQUIRE(cfg != nullptr && out != nullptr, "create: null argument");
  W2V_REQUIRE(cfg->hidden == 1024 && cfg->conv_dim == 512,
              "create: hidden=%d conv_dim=%d unsupported (XLS-R-300m geometry 1024/512 only)",
              cfg->hidden, cfg->conv_dim);
  W2V_REQUIRE(cfg->heads > 0 && cfg->hidden / cfg->heads == 64, "create: encoder head_dim must be 64");
  W2V_REQUIRE(cfg->n_layers >= 0 && cfg->n_adapter_layers >= 0 && cfg->n_adapter_layers <= cfg->n_layers,
              "create: bad layer counts (%d layers, %d adapters)", cfg->n_layers, cfg->n_adapter_layers);
  W2V_REQUIRE(!(cfg->post_layer_norm && cfg->n_adapter_layers > 0),
              "create: FFN adapters exist for the stable-LayerNorm encoder layer only (lib/models.py:390-428)");
  W2V_REQUIRE(cfg->ffn % 256 == 0 && cfg->adapter_dim % 256 == 0 && cfg->head_ffn % 256 == 0,
              "create: ffn/adapter/head_ffn sizes must be multiples of 256");
  W2V_REQUIRE(cfg->pos_kernel == 128 && cfg->pos_groups == 16, "create: positional conv must be k=128, g=16");
  W2V_REQUIRE(cfg->head_layers == 0 || cfg->head_layers == 1, "create: head_layers must be 0 or 1");
  if (cfg->head_layers == 1) {
    const int hd = cfg->hidden / (cfg->head_heads > 0 ? cfg->head_heads : 1);
    W2V_REQUIRE(cfg->head_heads > 0 && (hd == 64 || hd == 128), "create: head head_dim must be 64 or 128");
  }
  W2V_TRY(w2vseg_device_ok());

  w2vseg_handle* h = new w2vseg_handle();
  h->cfg = *cfg;
  h->D = cfg->hidden;
  h->DH = cfg->hidden / cfg->heads;
  h->F1max = cfg->ffn + (cfg->n_adapter_layers > 0 ? cfg->adapter_dim : 0);
  {
    const char* e = getenv("W2VSEG_CONV0");
    h->conv0_cuda_cores = e != nullptr && strcmp(e, "cuda") == 0;
  }
  if (cfg->head_layers > 0 && cfg->head_ffn > h->F1max) h->F1max = cfg->head_ffn;
  build_layout(h);  // sizing pass
  h->arena_bytes = h->arena_used + 4096;
  if (cudaMalloc(&h->arena, h->arena_bytes) != cudaSuccess ||
      cudaMemset(h->arena, 0, h->arena_bytes) != cudaSuccess) {
    set_error("create: cudaMalloc of %zu weight bytes failed", h->arena_bytes);
    delete h;
    return W2VSEG_ERR_CUDA;
  }
  build_layout(h);
  *out = h;
  return 0;
}

void w2vseg_destroy(w2vseg_handle* h) {
  if (h == nullptr) return;
  if (h->arena != nullptr) cudaFree(h->arena);
  delete h;
}

int32_t w2vseg_set_weight(w2vseg_handle* h, const char* name, const float* src, int64_t numel,
                          void* stream) {
  W2V_REQUIRE(h != nullptr && name != nullptr && src != nullptr, "set_weight: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  auto it = h->slots.find(name);
  W2V_REQUIRE(it != h->slots.end(), "set_weight: unknown tensor name '%s'", name);
  const Slot& s = it->second;
  W2V_REQUIRE(numel == s.numel, "set_weight: '%s' has %lld elements, expected %lld", name,
              (long long)numel, (long long)s.numel);
  switch (s.kind) {
    case SLOT_VEC:
      W2V_TRY(axpby_launch(src, s.scale, nullptr, 0.f, s.fdst, (int)s.numel, st));
      break;
    case SLOT_RAW:
      W2V_CHECK_CUDA(cudaMemcpyAsync(s.fdst, src, s.numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
      break;
    case SLOT_MAT:
      W2V_TRY(pack_matrix_launch(src, s.rows, s.cols, s.scale, s.bdst, s.ld, st));
      break;
    case SLOT_CONV:
      W2V_TRY(pack_conv_launch(src, s.O, s.I, s.J, nullptr, s.bdst, st));
      break;
    case SLOT_CONV0:
      W2V_TRY(transpose_f32_launch(src, s.O, s.J, s.fdst, st));
      break;
  }
  h->was_set[name] = true;
  if (h->finalized && strncmp(name, "head.", 5) == 0) {
    // training loop (w2vseg_head_train_step): head parameters are re-uploaded every optimiser step. The handle
    // stays finalised; a re-uploaded head bias becomes the effective bias again (any bias correction of that
    // vector is dropped, as after w2vseg_finalize_weights).
    if (s.kind == SLOT_VEC)
      for (const BiasPair& bp : h->bias_pairs)
        if (bp.raw == s.fdst) W2V_TRY(axpby_launch(bp.raw, 1.f, nullptr, 0.f, bp.eff, bp.n, st));
    h->corrected.erase(name);
    return 0;
  }
  h->finalized = false;
  h->calibrated = false;
  return 0;
}

int32_t w2vseg_finalize_weights(w2vseg_handle* h, void* stream) {
  W2V_REQUIRE(h != nullptr, "finalize: null handle");
  cudaStream_t st = (cudaStream_t)stream;
  const bool has_gv = h->was_set.count("pos.weight_g") && h->was_set.count("pos.weight_v");
  const bool has_w = h->was_set.count("pos.weight") > 0;
  for (const auto& kv : h->slots) {
    if (kv.second.alt_group != 0) continue;
    if (!h->was_set.count(kv.first)) {
      set_error("finalize: tensor '%s' was never set", kv.first.c_str());
      return W2VSEG_ERR_STATE;
    }
  }
  if (!has_gv && !has_w) {
    set_error("finalize: positional conv needs pos.weight_g + pos.weight_v (weight-norm) or pos.weight");
    return W2VSEG_ERR_STATE;
  }
  const w2vseg_config& c = h->cfg;
  if (has_gv) {
    // weight_norm(dim=2): W[o,i,j] = g[j] * v[o,i,j] / ||v[:,:,j]||  (HF:343-355)
    const int gc = h->D / c.pos_groups;
    W2V_TRY(weightnorm_scale_launch(h->pos_v, h->pos_g, h->D * gc, c.pos_kernel, h->pos_scale, st));
    W2V_TRY(pack_conv_launch(h->pos_v, h->D, gc, c.pos_kernel, h->pos_scale, h->pos_w, st));
  }
  // conv layer 0 (LayerNorm variant): centred, gamma-scaled fp16 taps + Cholesky factor of the channel Gram matrix
  if (!c.feat_group_norm)
    W2V_TRY(conv0_tc_pack_launch(h->conv0_wt, 1, c.conv_dim, h->conv_b[0], h->conv_ln[0].g,
                                 h->conv_ln[0].b, h->conv0_pack, st));
  for (auto& L : h->enc) {
    if (L.adapter) W2V_TRY(axpby_launch(L.b2_raw, 1.f, L.bu_raw, c.adapter_scale, L.b2, h->D, st));
    else W2V_TRY(axpby_launch(L.b2_raw, 1.f, nullptr, 0.f, L.b2, h->D, st));
  }
  // effective GEMM biases = checkpoint values (any earlier bias correction is dropped)
  for (const BiasPair& bp : h->bias_pairs) W2V_TRY(axpby_launch(bp.raw, 1.f, nullptr, 0.f, bp.eff, bp.n, st));
  h->corrected.clear();
  h->calibrated = false;
  h->finalized = true;
  return 0;
}

size_t w2vseg_workspace_bytes(const w2vseg_handle* h, int32_t B, int64_t l_max) {
  if (h == nullptr || B <= 0 || l_max <= 0) return 0;
  const int R = w2vseg_frame_stride(l_max);
  return carve(h, nullptr, B, R).bytes + 1024;
}

static int check_ws(const w2vseg_handle* h, void* workspace, size_t workspace_bytes, int B, int R,
                    Workspace* w) {
  W2V_REQUIRE(workspace != nullptr, "null workspace");
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  *w = carve(h, base, B, R);
  const size_t need = w->bytes + (size_t)(base - reinterpret_cast<uint8_t*>(workspace));
  W2V_REQUIRE(workspace_bytes >= need, "workspace too small: %zu bytes given, %zu needed (B=%d, R=%d)",
              workspace_bytes, need, B, R);
  return 0;
}

int32_t w2vseg_calibrate(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                         const int32_t* sample_len, const int32_t* norm_len, const int32_t* out_len,
                         int32_t B, int64_t l_max, void* workspace, size_t workspace_bytes, void* stream) {
  W2V_TRY(check_ready(h));
  W2V_REQUIRE(audio && sample_len && norm_len && out_len, "calibrate: null argument");
  W2V_REQUIRE(B > 0 && l_max >= 400 && audio_stride >= l_max, "calibrate: bad shape (B=%d, l_max=%lld)", B,
              (long long)l_max);
  if (!h->corrected.empty()) {
    set_error("calibrate: biases already carry corrections; call w2vseg_finalize_weights first");
    return W2VSEG_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int R = w2vseg_frame_stride(l_max);
  Workspace w;
  W2V_TRY(check_ws(h, workspace, workspace_bytes, B, R, &w));
  W2V_TRY(run_encoder(h, w, audio, audio_stride, sample_len, norm_len, B, R, st, /*calib*/ true, l_max));
  W2V_TRY(run_head(h, w, w.h, B, R, out_len, nullptr, nullptr, st, 0, 0, -1, /*calib*/ true));
  h->calibrated = true;
  // the positional conv keeps its fp32 weight-norm factors in the handle: correct its bias right here
  if (h->was_set.count("pos.weight_g") && h->was_set.count("pos.weight_v")) {
    const int gc = h->D / h->cfg.pos_groups, J = h->cfg.pos_kernel;
    W2V_TRY(bias_correct_launch(h->pos_v, h->pos_w, (int64_t)gc * J, h->D, gc * J, 1.f, 1, gc, J, h->pos_scale,
                                h->pos_x, h->D, gc, h->pos_b, st));
    h->corrected["pos.weight_v"] = true;
  }
  return 0;
}

int32_t w2vseg_correct_bias(w2vseg_handle* h, const char* name, const float* src, int64_t numel, void* stream) {
  W2V_REQUIRE(h != nullptr && name != nullptr && src != nullptr, "correct_bias: null argument");
  if (!h->finalized || !h->calibrated) {
    set_error("correct_bias: needs finalised weights and a w2vseg_calibrate pass");
    return W2VSEG_ERR_STATE;
  }
  auto it = h->slots.find(name);
  W2V_REQUIRE(it != h->slots.end(), "correct_bias: unknown tensor name '%s'", name);
  const Slot& s = it->second;
  W2V_REQUIRE(numel == s.numel, "correct_bias: '%s' has %lld elements, expected %lld", name, (long long)numel,
              (long long)s.numel);
  if (s.c_bias == nullptr || (s.kind != SLOT_MAT && s.kind != SLOT_CONV)) return 0;   // nothing rounded to bf16
  if (h->corrected.count(name)) {
    set_error("correct_bias: '%s' was already corrected for this set of weights", name);
    return W2VSEG_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (s.kind == SLOT_MAT)
    W2V_TRY(bias_correct_launch(src, s.bdst, s.ld, s.rows, s.cols, s.scale, 0, 0, 0, nullptr, s.c_x, 0, 1, s.c_bias, st));
  else
    W2V_TRY(bias_correct_launch(src, s.bdst, (int64_t)s.I * s.J, s.O, s.I * s.J, 1.f, 1, s.I, s.J, nullptr, s.c_x,
                                s.c_x_tap_stride, s.I, s.c_bias, st));
  h->corrected[name] = true;
  return 0;
}

int32_t w2vseg_encode(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                      const int32_t* sample_len, const int32_t* norm_len, int32_t B, int64_t l_max,
                      float* hidden_out, int32_t* enc_len_out, int32_t* included_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  W2V_TRY(check_ready(h));
  W2V_REQUIRE(audio && sample_len && norm_len && hidden_out, "encode: null argument");
  W2V_REQUIRE(B > 0 && l_max >= 400 && audio_stride >= l_max, "encode: bad shape (B=%d, l_max=%lld, stride=%lld)",
              B, (long long)l_max, (long long)audio_stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int R = w2vseg_frame_stride(l_max);
  Workspace w;
  W2V_TRY(check_ws(h, workspace, workspace_bytes, B, R, &w));
  W2V_TRY(run_encoder(h, w, audio, audio_stride, sample_len, norm_len, B, R, st, false, l_max));
  W2V_CHECK_CUDA(cudaMemcpyAsync(hidden_out, w.h, (size_t)B * R * h->D * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
  if (enc_len_out != nullptr)
    W2V_CHECK_CUDA(cudaMemcpyAsync(enc_len_out, w.enc_len, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
  if (included_out != nullptr)
    W2V_CHECK_CUDA(cudaMemcpyAsync(included_out, w.included, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int32_t w2vseg_head(w2vseg_handle* h, const float* hidden, int64_t batch_stride, int32_t T,
                    const int32_t* out_len, int32_t B, float* logits_out, float* probs_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  W2V_TRY(check_ready(h));
  W2V_REQUIRE(hidden && out_len, "head: null argument");
  W2V_REQUIRE(B > 0 && T > 0 && batch_stride >= (int64_t)T * h->D && batch_stride % 4 == 0,
              "head: bad shape (B=%d, T=%d, batch_stride=%lld)", B, T, (long long)batch_stride);
  cudaStream_t st = (cudaStream_t)stream;
  // the workspace was sized for R >= T rows per window; use T as the row stride here
  Workspace w;
  W2V_TRY(check_ws(h, workspace, workspace_bytes, B, T, &w));
  W2V_TRY(gather_rows_launch(hidden, batch_stride, B, T, h->D, w.h, st));
  W2V_TRY(run_head(h, w, w.h, B, T, out_len, logits_out, probs_out, st));
  return 0;
}

int32_t w2vseg_sfc_forward(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                           const int32_t* sample_len, const int32_t* norm_len,
                           const int32_t* out_len, int32_t B, int64_t l_max, float* logits_out,
                           float* probs_out, int32_t* included_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  W2V_TRY(check_ready(h));
  W2V_REQUIRE(audio && sample_len && norm_len && out_len, "sfc_forward: null argument");
  W2V_REQUIRE(B > 0 && l_max >= 400 && audio_stride >= l_max,
              "sfc_forward: bad shape (B=%d, l_max=%lld, stride=%lld)", B, (long long)l_max,
              (long long)audio_stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int R = w2vseg_frame_stride(l_max);
  Workspace w;
  W2V_TRY(check_ws(h, workspace, workspace_bytes, B, R, &w));
  W2V_TRY(run_encoder(h, w, audio, audio_stride, sample_len, norm_len, B, R, st, false, l_max));
  W2V_TRY(run_head(h, w, w.h, B, R, out_len, logits_out, probs_out, st));
  if (included_out != nullptr)
    W2V_CHECK_CUDA(cudaMemcpyAsync(included_out, w.included, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int32_t w2vseg_sfc_forward_rows(w2vseg_handle* h, const float* audio, int64_t audio_stride,
                                const int32_t* sample_len, const int32_t* norm_len,
                                const int32_t* out_len, int32_t B, int64_t l_max, float* rows_out,
                                int64_t row_stride, int32_t row_cols, int32_t flag_col, void* workspace,
                                size_t workspace_bytes, void* stream) {
  W2V_TRY(check_ready(h));
  W2V_REQUIRE(audio && sample_len && norm_len && out_len && rows_out, "sfc_forward_rows: null argument");
  W2V_REQUIRE(B > 0 && l_max >= 400 && audio_stride >= l_max,
              "sfc_forward_rows: bad shape (B=%d, l_max=%lld, stride=%lld)", B, (long long)l_max,
              (long long)audio_stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int R = w2vseg_frame_stride(l_max);
  W2V_REQUIRE(row_cols >= R && row_stride >= row_cols && flag_col < row_cols && (flag_col < 0 || flag_col >= R),
              "sfc_forward_rows: row layout (stride %lld, cols %d, flag col %d) does not hold R=%d frames + flag",
              (long long)row_stride, row_cols, flag_col, R);
  Workspace w;
  W2V_TRY(check_ws(h, workspace, workspace_bytes, B, R, &w));
  W2V_TRY(run_encoder(h, w, audio, audio_stride, sample_len, norm_len, B, R, st, false, l_max));
  W2V_TRY(run_head(h, w, w.h, B, R, out_len, nullptr, rows_out, st, row_stride, row_cols, flag_col));
  return 0;
}

// ---- talk-level reductions -------------------------------------------------------------------
int32_t w2vseg_scatter_rows(const float* rows, int64_t row_stride, const int32_t* start,
                            const int32_t* count, int32_t n_rows, double* talk, int64_t n_frames,
                            int32_t flag_col, void* stream) {
  W2V_REQUIRE(talk != nullptr && n_frames >= 0 && n_rows >= 0, "scatter_rows: bad argument");
  W2V_REQUIRE(n_rows == 0 || (rows && start && count), "scatter_rows: null argument");
  W2V_REQUIRE(flag_col < row_stride, "scatter_rows: flag column outside the row");
  return scatter_rows_launch(rows, row_stride, start, count, n_rows, talk, n_frames, flag_col, (cudaStream_t)stream);
}
int32_t w2vseg_nanfill(double* talk, int64_t n_frames, const int32_t* idx, int32_t n_idx, void* stream) {
  W2V_REQUIRE(talk != nullptr && (n_idx == 0 || idx != nullptr), "nanfill: null argument");
  return nanfill_launch(talk, n_frames, idx, n_idx, (cudaStream_t)stream);
}
int32_t w2vseg_overlap_average(const double* tilings, int32_t n_tilings, int64_t n_frames,
                               double* out, void* stream) {
  W2V_REQUIRE(tilings && out, "overlap_average: null argument");
  return overlap_average_launch(tilings, n_tilings, n_frames, out, (cudaStream_t)stream);
}
int32_t w2vseg_moving_average(const double* arr, int64_t n, int32_t window, double* out, void* stream) {
  W2V_REQUIRE(arr && out && arr != out, "moving_average: null or aliased argument");
  return moving_average_launch(arr, n, window, out, (cudaStream_t)stream);
}

// ---- single kernels -----------------------------------------------------------------------------
int32_t w2vseg_gemm(const void* A, const void* W, int32_t M, int32_t N, int32_t K, const float* bias,
                    int32_t act, const float* resid, void* out, int32_t out_f32, int32_t block_n,
                    void* stream) {
  W2V_REQUIRE(A && W && out && M > 0, "gemm: bad argument");
  W2V_TRY(w2vseg_device_ok());
  GemmProblem g = linear((const bf16*)A, M, K, (const bf16*)W, N, bias);
  g.act_lo = act; g.act_hi = act;
  g.resid = resid; g.ld_resid = N; g.out = out; g.ld_out = N; g.out_f32 = out_f32;
  if (block_n == 512) return gemm_tc2_launch(g, (cudaStream_t)stream);  // CTA-pair 256x256 tiles
  return gemm_tc_launch(g, block_n, (cudaStream_t)stream);
}

int32_t w2vseg_conv_gemm(const void* x, int64_t rows_out, int32_t C, int32_t kw, int32_t stride,
                         const void* W, int32_t N, const float* bias, void* out, void* stream) {
  W2V_REQUIRE(x && W && out && rows_out > 0, "conv_gemm: bad argument");
  W2V_TRY(w2vseg_device_ok());
  GemmProblem g = linear((const bf16*)x, rows_out, kw * C, (const bf16*)W, N, bias);
  g.a_row_stride = (int64_t)stride * C;
  g.out = out; g.ld_out = N;
  // N % 256 == 0 (the feature extractor: N = 512): the CTA-pair kernel the forward pass itself runs on this
  // overlapping-row view; other widths fall back to the single-CTA kernel
  if (N % 256 == 0) return gemm_tc2_launch(g, (cudaStream_t)stream);
  return gemm_tc_launch(g, N % 128 == 0 ? 128 : 64, (cudaStream_t)stream);
}

int32_t w2vseg_posconv(const void* zpad, const void* W, const float* bias, int32_t B, int32_t R, int32_t D,
                       int32_t taps, float* h, int32_t impl, void* stream) {
  W2V_REQUIRE(zpad && W && h && B > 0 && R > 0, "posconv: bad argument");
  W2V_REQUIRE(D % 64 == 0 && taps > 0 && taps % 2 == 0 && taps <= 128, "posconv: D=%d taps=%d unsupported", D, taps);
  W2V_TRY(w2vseg_device_ok());
  GemmProblem g = posconv_problem((const bf16*)zpad, (const bf16*)W, bias, B, R, D, taps, h);
  return impl == 0 ? posconv_tc_launch(g, (cudaStream_t)stream) : gemm_tc_launch(g, 64, (cudaStream_t)stream);
}

int32_t w2vseg_conv0(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                     const float* stats, const float* w, const float* bias, const float* gamma,
                     const float* beta, float eps, void* out, int32_t B, int32_t R0, int32_t impl,
                     void* scratch, size_t scratch_bytes, void* stream) {
  W2V_REQUIRE(audio && sample_len && stats && w && bias && gamma && beta && out && scratch,
              "conv0: null argument");
  W2V_REQUIRE(B > 0 && R0 > 0 && scratch_bytes >= 65536, "conv0: bad shape or scratch < 64 KiB");
  W2V_TRY(w2vseg_device_ok());
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 0) {
    W2V_TRY(conv0_tc_pack_launch(w, 10, 1, bias, gamma, beta, scratch, st));
    return conv0_tc_launch(audio, audio_stride, sample_len, (const float2*)stats, scratch, eps, (bf16*)out, B, R0, st);
  }
  float* wt = reinterpret_cast<float*>(scratch);
  W2V_TRY(transpose_f32_launch(w, 512, 10, wt, st));
  return conv0_ln_gelu_launch(audio, audio_stride, sample_len, (const float2*)stats, wt, bias, gamma, beta, eps,
                              (bf16*)out, B, R0, st);
}

int32_t w2vseg_layernorm(const void* in, int32_t in_f32, int64_t rows, int32_t C, const float* gamma,
                         const float* beta, float eps, int32_t act, void* out, void* stream) {
  W2V_REQUIRE(in && gamma && beta && out, "layernorm: null argument");
  return layernorm_launch(in, in_f32 != 0, rows, C, gamma, beta, eps, act, (bf16*)out, (cudaStream_t)stream);
}

int32_t w2vseg_clock_probe(float* mhz_out, int32_t n_blocks, int32_t spin_us, void* stream) {
  W2V_REQUIRE(mhz_out != nullptr && n_blocks > 0 && spin_us > 0 && spin_us <= 10000, "clock_probe: bad argument");
  return clock_probe_launch(mhz_out, n_blocks, spin_us, (cudaStream_t)stream);
}

int32_t w2vseg_attention(const void* qkv, int32_t B, int32_t R, int32_t heads, int32_t head_dim,
                         const int32_t* kv_len, float scale, void* ctx, void* stream) {
  W2V_REQUIRE(qkv && kv_len && ctx, "attention: null argument");
  W2V_TRY(w2vseg_device_ok());
  return attention_tc_launch((const bf16*)qkv, B, R, heads, head_dim, kv_len, scale, (bf16*)ctx,
                             (cudaStream_t)stream);
}

int32_t w2vseg_attention_mma(const void* qkv, int32_t B, int32_t R, int32_t heads, int32_t head_dim,
                             const int32_t* kv_len, float scale, void* ctx, void* stream) {
  W2V_REQUIRE(qkv && kv_len && ctx, "attention: null argument");
  return attention_launch((const bf16*)qkv, B, R, heads, head_dim, kv_len, scale, (bf16*)ctx,
                          (cudaStream_t)stream);
}

int32_t w2vseg_attention_train(const void* qkv, int32_t B, int32_t R, int32_t heads, int32_t head_dim,
                               const int32_t* kv_len, float scale, void* ctx, float* lse, float dropout,
                               uint32_t seed, void* stream) {
  W2V_REQUIRE(qkv && kv_len && ctx && lse, "attention_train: null argument");
  W2V_REQUIRE(dropout >= 0.f && dropout < 1.f, "attention_train: dropout %g outside [0, 1)", (double)dropout);
  const DropSite d = make_drop_site(dropout, seed, 1);
  return attention_launch((const bf16*)qkv, B, R, heads, head_dim, kv_len, scale, (bf16*)ctx,
                          (cudaStream_t)stream, lse, &d);
}

int32_t w2vseg_attention_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse,
                             float* delta_scratch, int32_t B, int32_t R, int32_t heads, int32_t head_dim,
                             const int32_t* kv_len, float scale, void* dqkv, float dropout, uint32_t seed,
                             void* stream) {
  W2V_REQUIRE(qkv && ctx && dctx && lse && delta_scratch && kv_len && dqkv, "attention_bwd: null argument");
  W2V_REQUIRE(dropout >= 0.f && dropout < 1.f, "attention_bwd: dropout %g outside [0, 1)", (double)dropout);
  return attention_bwd_launch((const bf16*)qkv, (const bf16*)ctx, (const bf16*)dctx, lse, delta_scratch, B, R,
                              heads, head_dim, kv_len, scale, (bf16*)dqkv, make_drop_site(dropout, seed, 1),
                              (cudaStream_t)stream);
}

}  // extern "C"

// ---- head-only training step (frozen encoder) --------------------------------------------------------------
namespace {

struct GradSlot { const char* name; int64_t off; int64_t numel; };

// gradient buffer layout: every seg_model parameter in its PyTorch shape, fp32, in this order
std::vector<GradSlot> head_grad_layout(const w2vseg_handle* h) {
  const int64_t D = h->D, F = h->cfg.head_ffn;
  std::vector<GradSlot> v;
  int64_t off = 0;
  auto add = [&](const char* n, int64_t numel) { v.push_back({n, off, numel}); off += (numel + 63) / 64 * 64; };
  if (h->cfg.head_layers > 0) {
    add("head.in_proj.weight", 3 * D * D); add("head.in_proj.bias", 3 * D);
    add("head.o.weight", D * D);           add("head.o.bias", D);
    add("head.ff1.weight", F * D);         add("head.ff1.bias", F);
    add("head.ff2.weight", D * F);         add("head.ff2.bias", D);
    add("head.ln1.weight", D); add("head.ln1.bias", D);
    add("head.ln2.weight", D); add("head.ln2.bias", D);
  }
  add("head.ln_f.weight", D); add("head.ln_f.bias", D);
  add("head.out.weight", D);  add("head.out.bias", 1);
  v.push_back({nullptr, off, 0});
  return v;
}

struct TrainWs {
  float *x0, *x1, *x2, *dx2, *dx1, *dlogit, *loss_rows, *lse, *delta, *red, *tmpA, *tmpS;
  float2 *st0, *st1, *st2;
  bf16 *u1, *u2, *qkv, *ctx, *z1, *m, *dx2b, *dm, *dz1, *du2, *dx1b, *dctx, *dqkv, *du1, *ta, *tb, *wt;
  size_t bytes;
};
constexpr size_t kRedFloats = 64 * 2 * 4096;

TrainWs carve_train(const w2vseg_handle* h, uint8_t* base, int B, int T) {
  TrainWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    off = align_up(off, 1024);
    uint8_t* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const size_t M = (size_t)B * T, D = h->D, F = h->cfg.head_ffn, Mp = (M + 63) / 64 * 64;
  const size_t heads = h->cfg.head_heads > 0 ? h->cfg.head_heads : 1;
  auto f32 = [&](size_t n) { return reinterpret_cast<float*>(take(n * sizeof(float))); };
  auto b16 = [&](size_t n) { return reinterpret_cast<bf16*>(take(n * sizeof(bf16))); };
  w.x0 = f32(M * D); w.x1 = f32(M * D); w.x2 = f32(M * D); w.dx2 = f32(M * D); w.dx1 = f32(M * D);
  w.dlogit = f32(M); w.loss_rows = f32(M); w.lse = f32(heads * M); w.delta = f32(heads * M);
  w.red = f32(kRedFloats); w.tmpA = f32(4096); w.tmpS = f32(4096);
  w.st0 = reinterpret_cast<float2*>(take(M * sizeof(float2)));
  w.st1 = reinterpret_cast<float2*>(take(M * sizeof(float2)));
  w.st2 = reinterpret_cast<float2*>(take(M * sizeof(float2)));
  w.u1 = b16(M * D); w.u2 = b16(M * D); w.qkv = b16(M * 3 * D); w.ctx = b16(M * D);
  w.z1 = b16(M * F); w.m = b16(M * F); w.dx2b = b16(M * D); w.dm = b16(M * F); w.dz1 = b16(M * F);
  w.du2 = b16(M * D); w.dx1b = b16(M * D); w.dctx = b16(M * D); w.dqkv = b16(M * 3 * D); w.du1 = b16(M * D);
  const size_t wide = std::max<size_t>(3 * D, F);
  w.ta = b16(wide * Mp); w.tb = b16(wide * Mp);       // transposed activations [cols, Mp] for the wgrad GEMMs
  w.wt = b16(std::max<size_t>(3 * D * D, D * F));    // transposed weight for the dgrad GEMMs
  w.bytes = align_up(off, 1024);
  return w;
}

// out[M, N] = A[M, K] * W[N, K]^T (+ bias); plain problem on the CTA-pair kernel
int gemm_plain(const bf16* A, int64_t M, int K, const bf16* W, int N, const float* bias, void* out, int64_t ld_out,
               bool out_f32, float* resid, const char* tag, cudaStream_t st) {
  GemmProblem g = linear(A, M, K, W, N, bias);
  g.out = out; g.ld_out = ld_out; g.out_f32 = out_f32 ? 1 : 0;
  if (resid != nullptr) { g.resid = resid; g.ld_resid = ld_out; }
  prof_tag(tag);
  return gemm_tc2_launch(g, st);
}
// dW[N_w, K_w] (fp32) = dY[M, N_w]^T X[M, K_w]  as  A' = dY^T [N_w, Mp], W' = X^T [K_w, Mp]
int wgrad(const bf16* dY, int64_t ld_dy, int N_w, const bf16* X, int64_t ld_x, int K_w, int64_t M, const TrainWs& w,
          float* dW, cudaStream_t st) {
  const int64_t Mp = (M + 63) / 64 * 64;
  W2V_TRY(transpose_bf16_launch(dY, ld_dy, M, N_w, w.ta, Mp, st));
  W2V_TRY(transpose_bf16_launch(X, ld_x, M, K_w, w.tb, Mp, st));
  return gemm_plain(w.ta, N_w, (int)Mp, w.tb, K_w, nullptr, dW, K_w, true, nullptr, "train.wgrad", st);
}
// dX[M, K_w] (bf16) = dY[M, N_w] W[N_w, K_w]  as  W' = W^T [K_w, N_w]
int dgrad(const bf16* dY, int64_t M, int N_w, const bf16* W, int64_t ld_w, int K_w, const TrainWs& w, bf16* dX,
          cudaStream_t st) {
  W2V_TRY(transpose_bf16_launch(W, ld_w, N_w, K_w, w.wt, N_w, st));
  return gemm_plain(dY, M, N_w, w.wt, K_w, nullptr, dX, K_w, false, nullptr, "train.dgrad", st);
}

}  // namespace

extern "C" {

int64_t w2vseg_head_grad_floats(const w2vseg_handle* h) {
  if (h == nullptr) return 0;
  return head_grad_layout(h).back().off;
}
int64_t w2vseg_head_grad_offset(const w2vseg_handle* h, const char* name, int64_t* numel_out) {
  if (h == nullptr || name == nullptr) return -1;
  for (const GradSlot& g : head_grad_layout(h))
    if (g.name != nullptr && strcmp(g.name, name) == 0) {
      if (numel_out != nullptr) *numel_out = g.numel;
      return g.off;
    }
  return -1;
}
size_t w2vseg_head_train_workspace_bytes(const w2vseg_handle* h, int32_t B, int32_t T) {
  if (h == nullptr || B <= 0 || T <= 0) return 0;
  return carve_train(h, nullptr, B, T).bytes + 1024;
}

int32_t w2vseg_head_train_step(w2vseg_handle* h, const float* hidden, int64_t batch_stride, int32_t T,
                               const int32_t* out_len, const float* target, float pos_weight, int32_t B,
                               float* loss_out, float* logits_out, float* grads, size_t grads_floats,
                               float init_dropout, float layer_dropout, uint32_t seed,
                               void* workspace, size_t workspace_bytes, void* stream) {
  W2V_TRY(check_ready(h));
  W2V_REQUIRE(init_dropout >= 0.f && init_dropout < 1.f && layer_dropout >= 0.f && layer_dropout < 1.f,
              "head_train_step: dropout (%g, %g) outside [0, 1)", (double)init_dropout, (double)layer_dropout);
  W2V_REQUIRE(hidden && out_len && target && loss_out && grads && workspace, "head_train_step: null argument");
  W2V_REQUIRE(B > 0 && T > 0 && batch_stride >= (int64_t)T * h->D && batch_stride % 4 == 0,
              "head_train_step: bad shape (B=%d, T=%d, batch_stride=%lld)", B, T, (long long)batch_stride);
  W2V_REQUIRE(h->cfg.head_layers == 1, "head_train_step: needs the 1-layer transformer head");
  const std::vector<GradSlot> lay = head_grad_layout(h);
  W2V_REQUIRE(grads_floats >= (size_t)lay.back().off, "head_train_step: gradient buffer too small (%zu floats, %lld needed)",
              grads_floats, (long long)lay.back().off);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  const TrainWs w = carve_train(h, base, B, T);
  W2V_REQUIRE(workspace_bytes >= w.bytes + (size_t)(base - reinterpret_cast<uint8_t*>(workspace)),
              "head_train_step: workspace too small (%zu bytes given, %zu needed)", workspace_bytes, w.bytes + 1024);
  auto G = [&](const char* name) -> float* {
    for (const GradSlot& g : lay)
      if (g.name != nullptr && strcmp(g.name, name) == 0) return grads + g.off;
    return nullptr;
  };
  const w2vseg_config& c = h->cfg;
  const HeadW& H = h->head;
  const int D = h->D, F = c.head_ffn, hd = D / c.head_heads;
  const int64_t M = (int64_t)B * T;
  const float scale = 1.0f / sqrtf((float)hd);

  // Dropout sites of the head in train() mode (lib/models.py:291-319): 0 init_dropout on the encoder output,
  // 1 attention weights, 2 after the attention block (dropout1), 3 inside the FFN, 4 after the FFN (dropout2).
  // Masks are regenerated from (seed, site, element index) wherever they are needed (dropout.cuh).
  const DropSite d0 = make_drop_site(init_dropout, seed, 0), d1 = make_drop_site(layer_dropout, seed, 1),
                 d2 = make_drop_site(layer_dropout, seed, 2), d3 = make_drop_site(layer_dropout, seed, 3),
                 d4 = make_drop_site(layer_dropout, seed, 4);
  const bool drop_layer = layer_dropout > 0.f;

  // ---------------- forward, keeping what the backward needs
  if (init_dropout > 0.f) W2V_TRY(gather_dropout_launch(hidden, batch_stride, B, T, D, w.x0, d0, st));
  else W2V_TRY(gather_rows_launch(hidden, batch_stride, B, T, D, w.x0, st));
  W2V_TRY(layernorm_launch(w.x0, true, M, D, H.ln1.g, H.ln1.b, c.ln_eps, 0, w.u1, st));
  W2V_TRY(gemm_plain(w.u1, M, D, H.win, 3 * D, H.bin, w.qkv, 3 * D, false, nullptr, "train.fwd", st));
  W2V_TRY(attention_launch(w.qkv, B, T, c.head_heads, hd, out_len, scale, w.ctx, st, w.lse, &d1));
  if (drop_layer) {     // branch output to a scratch (dx1 is free until the backward), then x1 = x0 + dropout(branch)
    W2V_TRY(gemm_plain(w.ctx, M, D, H.wo, D, H.bo, w.dx1, D, true, nullptr, "train.fwd", st));
    W2V_TRY(resid_dropout_launch(w.x0, w.dx1, w.x1, M * D, d2, st));
  } else {
    W2V_TRY(copy_f32_launch(w.x0, w.x1, M * D, st));
    W2V_TRY(gemm_plain(w.ctx, M, D, H.wo, D, H.bo, w.x1, D, true, w.x1, "train.fwd", st));
  }
  W2V_TRY(layernorm_launch(w.x1, true, M, D, H.ln2.g, H.ln2.b, c.ln_eps, 0, w.u2, st));
  W2V_TRY(gemm_plain(w.u2, M, D, H.w1, F, H.b1, w.z1, F, false, nullptr, "train.fwd", st));
  W2V_TRY(gelu_fwd_launch(w.z1, w.m, M * F, d3, st));
  if (drop_layer) {
    W2V_TRY(gemm_plain(w.m, M, F, H.w2, D, H.b2, w.dx2, D, true, nullptr, "train.fwd", st));
    W2V_TRY(resid_dropout_launch(w.x1, w.dx2, w.x2, M * D, d4, st));
  } else {
    W2V_TRY(copy_f32_launch(w.x1, w.x2, M * D, st));
    W2V_TRY(gemm_plain(w.m, M, F, H.w2, D, H.b2, w.x2, D, true, w.x2, "train.fwd", st));
  }

  // ---------------- loss + backward
  W2V_TRY(head_loss_backward_launch(w.x2, B, T, H.lnf.g, H.lnf.b, c.ln_eps, H.wout, H.bout, out_len, target,
                                    pos_weight, w.dx2, w.dx2b, w.dlogit, w.st2, logits_out, w.loss_rows, loss_out, st));
  W2V_TRY(final_param_grads_launch(w.dlogit, w.x2, w.st2, M, D, H.lnf.g, H.lnf.b, H.wout, w.red, kRedFloats, w.tmpA,
                                   w.tmpS, G("head.out.weight"), G("head.ln_f.weight"), G("head.ln_f.bias"),
                                   G("head.out.bias"), st));
  // FFN: x2 = x1 + drop4(drop3(gelu(u2 W1^T + b1)) W2^T + b2); the branch sees the masked gradient (dx2b)
  if (drop_layer) {
    W2V_TRY(mask_cast_launch(w.dx2, w.dx2b, M * D, d4, st));
    W2V_TRY(colsum_launch(w.dx2b, true, D, M, D, w.red, kRedFloats, G("head.ff2.bias"), st));
  } else {
    W2V_TRY(colsum_launch(w.dx2, false, D, M, D, w.red, kRedFloats, G("head.ff2.bias"), st));
  }
  W2V_TRY(wgrad(w.dx2b, D, D, w.m, F, F, M, w, G("head.ff2.weight"), st));
  W2V_TRY(dgrad(w.dx2b, M, D, H.w2, F, F, w, w.dm, st));
  W2V_TRY(gelu_bwd_launch(w.z1, w.dm, w.dz1, M * F, d3, st));
  W2V_TRY(colsum_launch(w.dz1, true, F, M, F, w.red, kRedFloats, G("head.ff1.bias"), st));
  W2V_TRY(wgrad(w.dz1, F, F, w.u2, D, D, M, w, G("head.ff1.weight"), st));
  W2V_TRY(dgrad(w.dz1, M, F, H.w1, D, D, w, w.du2, st));
  // LN2: dx1 = dx2 + LN2'(du2)
  W2V_TRY(layernorm_bwd_launch(w.x1, w.du2, M, H.ln2.g, c.ln_eps, w.dx2, w.dx1, w.dx1b, w.st1, st));
  W2V_TRY(ln_param_grads_launch(w.du2, w.x1, w.st1, M, D, w.red, kRedFloats, G("head.ln2.weight"), G("head.ln2.bias"), st));
  // attention output projection: x1 = x0 + drop2(ctx Wo^T + bo)
  if (drop_layer) {
    W2V_TRY(mask_cast_launch(w.dx1, w.dx1b, M * D, d2, st));
    W2V_TRY(colsum_launch(w.dx1b, true, D, M, D, w.red, kRedFloats, G("head.o.bias"), st));
  } else {
    W2V_TRY(colsum_launch(w.dx1, false, D, M, D, w.red, kRedFloats, G("head.o.bias"), st));
  }
  W2V_TRY(wgrad(w.dx1b, D, D, w.ctx, D, D, M, w, G("head.o.weight"), st));
  W2V_TRY(dgrad(w.dx1b, M, D, H.wo, D, D, w, w.dctx, st));
  // attention
  W2V_TRY(attention_bwd_launch(w.qkv, w.ctx, w.dctx, w.lse, w.delta, B, T, c.head_heads, hd, out_len, scale, w.dqkv, d1, st));
  // input projection: qkv = u1 Win^T + bin
  W2V_TRY(colsum_launch(w.dqkv, true, 3 * D, M, 3 * D, w.red, kRedFloats, G("head.in_proj.bias"), st));
  W2V_TRY(wgrad(w.dqkv, 3 * D, 3 * D, w.u1, D, D, M, w, G("head.in_proj.weight"), st));
  W2V_TRY(dgrad(w.dqkv, M, 3 * D, H.win, D, D, w, w.du1, st));
  // LN1 parameters (the encoder is frozen: no gradient flows into x0)
  W2V_TRY(layernorm_bwd_launch(w.x0, w.du1, M, H.ln1.g, c.ln_eps, nullptr, nullptr, nullptr, w.st0, st));
  W2V_TRY(ln_param_grads_launch(w.du1, w.x0, w.st0, M, D, w.red, kRedFloats, G("head.ln1.weight"), G("head.ln1.bias"), st));
  return 0;
}

}  // extern "C"
