// CUDA-core kernels of the head-only training step (frozen encoder; reference train.py:381-480 around
// SegmentationFrameClassifier, lib/models.py:279-319): the loss + final-layer backward, LayerNorm backward,
// GELU forward / backward on stored pre-activations, deterministic column reductions for the bias and
// LayerNorm parameter gradients, and the transposes that turn dgrad / wgrad into A[M,K] * W[N,K]^T problems
// for the tcgen05 GEMM (gemm_tc2.cu). All reductions use fixed summation orders (no atomics).
#include <algorithm>

#include "kernels.cuh"
#include "ptx.cuh"
#include "train_kernels.cuh"
#include "dropout.cuh"

namespace w2v {

namespace {

inline unsigned blocks_for_t(long long n, int per) { return (unsigned)((n + per - 1) / per); }

// ---- row statistics shared by the LayerNorm backward kernels: one warp per row of C = 1024 ---------------
struct RowLN {
  float v[32];       // this lane's 32 elements: columns (i*32 + lane)*4 + e
  float mean, rstd;
};
__device__ __forceinline__ void row_ln_load(const float* __restrict__ x, long long row, int lane, float eps, RowLN& r) {
  const float4* p = reinterpret_cast<const float4*>(x) + row * 256;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 f = p[i * 32 + lane];
    r.v[4 * i] = f.x; r.v[4 * i + 1] = f.y; r.v[4 * i + 2] = f.z; r.v[4 * i + 3] = f.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += r.v[i];
  r.mean = warp_sum(s) * (1.f / 1024.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { const float d = r.v[i] - r.mean; q = fmaf(d, d, q); }
  r.rstd = rsqrtf(warp_sum(q) * (1.f / 1024.f) + eps);
}

// Final LayerNorm + Linear(1024 -> 1) + BCE-with-logits loss and their backward in one pass over x2
// (lib/models.py:317-319, train.py:416-459 with ma_window unset):
//   logit = LN(x2) . w + b;  loss_row = pos_weight * t * softplus(-logit) + (1 - t) * softplus(logit)   (masked rows: 0)
//   loss = sum over frames / B  (".sum(dim=1).mean()")   ->  dlogit = (sigma * (1 - t + pw t) - pw t) / B
//   dx2 = LayerNorm backward of dy = dlogit * w
// Writes dx2 (fp32 + bf16 copy), dlogit[row], stats[row] = (mean, rstd), logits (optional), loss_rows[row].
__global__ void __launch_bounds__(256)
head_loss_backward_kernel(const float* __restrict__ x2, long long rows, int R, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, const float* __restrict__ w_out,
                          const float* __restrict__ b_out, const int* __restrict__ out_len,
                          const float* __restrict__ target, float pos_weight, float inv_batch,
                          float* __restrict__ dx2, __nv_bfloat16* __restrict__ dx2_bf, float* __restrict__ dlogit,
                          float2* __restrict__ stats, float* __restrict__ logits, float* __restrict__ loss_rows) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  RowLN r;
  row_ln_load(x2, row, lane, eps, r);
  float g[32];       // gamma * w_out per column
  float dot = 0.f, sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 ga = *reinterpret_cast<const float4*>(gamma + col);
    const float4 be = *reinterpret_cast<const float4*>(beta + col);
    const float4 w = *reinterpret_cast<const float4*>(w_out + col);
    const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (r.v[4 * i + e] - r.mean) * r.rstd;
      r.v[4 * i + e] = xh;                                   // keep x-hat
      dot = fmaf(fmaf(xh, gam[e], bet[e]), ww[e], dot);
      g[4 * i + e] = gam[e] * ww[e];
      sg += g[4 * i + e];
      sgx = fmaf(g[4 * i + e], xh, sgx);
    }
  }
  dot = warp_sum(dot);
  sg = warp_sum(sg) * (1.f / 1024.f);
  sgx = warp_sum(sgx) * (1.f / 1024.f);
  const long long b = row / R;
  const int t = (int)(row - b * R);
  const bool keep = t < out_len[b];
  const float logit = dot + b_out[0];
  float dl = 0.f, loss = 0.f;
  if (keep) {
    const float tg = target[row];
    const float sp_pos = logit > 0.f ? logit + log1pf(expf(-logit)) : log1pf(expf(logit));   // softplus(logit)
    const float sp_neg = sp_pos - logit;                                                     // softplus(-logit)
    loss = (pos_weight * tg * sp_neg + (1.f - tg) * sp_pos) * inv_batch;
    const float sig = 1.f / (1.f + expf(-logit));
    dl = (sig * (1.f - tg + pos_weight * tg) - pos_weight * tg) * inv_batch;
  }
  if (lane == 0) {
    dlogit[row] = dl;
    stats[row] = make_float2(r.mean, r.rstd);
    loss_rows[row] = loss;
    if (logits != nullptr) logits[row] = keep ? logit : 0.f;
  }
  const float a = dl * r.rstd;
  float4* o = reinterpret_cast<float4*>(dx2) + row * 256;
  uint2* ob = reinterpret_cast<uint2*>(dx2_bf) + row * 256;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float d[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) d[e] = a * (g[4 * i + e] - sg - r.v[4 * i + e] * sgx);
    o[i * 32 + lane] = make_float4(d[0], d[1], d[2], d[3]);
    uint2 u;
    u.x = pack_bf16x2(d[0], d[1]);
    u.y = pack_bf16x2(d[2], d[3]);
    ob[i * 32 + lane] = u;
  }
}

// LayerNorm backward wrt the input (C = 1024): dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g * xhat)),
// g = dy * gamma, dy bf16. Writes fp32 (+ optional bf16 copy) and the row statistics for the parameter grads.
// dx_in / dx_out may be null (LN1 of the frozen-encoder step only needs the statistics).
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, long long rows,
                     const float* __restrict__ gamma, float eps, const float* __restrict__ dx_in,
                     float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  RowLN r;
  row_ln_load(x, row, lane, eps, r);
  if (lane == 0) stats[row] = make_float2(r.mean, r.rstd);
  if (dx_out == nullptr) return;
  float g[32];
  float sg = 0.f, sgx = 0.f;
  const uint2* pdy = reinterpret_cast<const uint2*>(dy) + row * 256;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 ga = *reinterpret_cast<const float4*>(gamma + col);
    const uint2 u = pdy[i * 32 + lane];
    const float2 d01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 d23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, dd[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (r.v[4 * i + e] - r.mean) * r.rstd;
      r.v[4 * i + e] = xh;
      g[4 * i + e] = dd[e] * gam[e];
      sg += g[4 * i + e];
      sgx = fmaf(g[4 * i + e], xh, sgx);
    }
  }
  sg = warp_sum(sg) * (1.f / 1024.f);
  sgx = warp_sum(sgx) * (1.f / 1024.f);
  const float4* pin = dx_in != nullptr ? reinterpret_cast<const float4*>(dx_in) + row * 256 : nullptr;
  float4* o = reinterpret_cast<float4*>(dx_out) + row * 256;
  uint2* ob = dx_bf != nullptr ? reinterpret_cast<uint2*>(dx_bf) + row * 256 : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 base = pin != nullptr ? pin[i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    float d[4] = {base.x, base.y, base.z, base.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) d[e] += r.rstd * (g[4 * i + e] - sg - r.v[4 * i + e] * sgx);
    o[i * 32 + lane] = make_float4(d[0], d[1], d[2], d[3]);
    if (ob != nullptr) {
      uint2 u;
      u.x = pack_bf16x2(d[0], d[1]);
      u.y = pack_bf16x2(d[2], d[3]);
      ob[i * 32 + lane] = u;
    }
  }
}

// ---- deterministic column reductions: partial[slab][...] over row slabs, then a fixed-order final sum ------
// mode 0: out0[c] = sum_r a[r, c]                                  (bias gradients)
// mode 1: out0[c] = sum_r a[r, c] * xhat[r, c], out1[c] = sum_r a[r, c]     (LayerNorm gamma / beta gradients;
//         xhat from x and the row statistics)
// mode 2: a[r, c] = rowscale[r] (x-hat weighted by the per-row dlogit): out0[c] = sum_r rowscale[r] * xhat[r, c],
//         out1[c] = sum_r rowscale[r]                              (final layer: d w_out, d gamma_f, d beta_f, d b_out)
template <typename TA>
__global__ void __launch_bounds__(256)
colreduce_partial_kernel(const TA* __restrict__ a, long long lda, const float* __restrict__ x,
                         const float2* __restrict__ stats, const float* __restrict__ rowscale, long long rows,
                         int C, int mode, float* __restrict__ partial) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = per * blockIdx.y, r1 = min(rows, r0 + per);
  // four rows in flight per thread (independent loads), four accumulators combined in a fixed order
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {
    float av[4], xv[4];
    float2 st[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      av[u] = mode == 2 ? rowscale[r + u] : (float)a[(r + u) * lda + c];
      if (mode != 0) { xv[u] = x[(r + u) * C + c]; st[u] = stats[r + u]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (mode == 0) { s0[u] += av[u]; continue; }
      s0[u] = fmaf(av[u], (xv[u] - st[u].x) * st[u].y, s0[u]);
      s1[u] += av[u];
    }
  }
  for (; r < r1; ++r) {
    const float av = mode == 2 ? rowscale[r] : (float)a[r * lda + c];
    if (mode == 0) { s0[0] += av; continue; }
    const float2 st = stats[r];
    s0[0] = fmaf(av, (x[r * C + c] - st.x) * st.y, s0[0]);
    s1[0] += av;
  }
  partial[((long long)blockIdx.y * 2 + 0) * C + c] = (s0[0] + s0[1]) + (s0[2] + s0[3]);
  partial[((long long)blockIdx.y * 2 + 1) * C + c] = (s1[0] + s1[1]) + (s1[2] + s1[3]);
}
__global__ void __launch_bounds__(256)
colreduce_final_kernel(const float* __restrict__ partial, int C, int slabs, float* __restrict__ out0,
                       float* __restrict__ out1) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f;
  for (int s = 0; s < slabs; ++s) {
    s0 += partial[((long long)s * 2 + 0) * C + c];
    s1 += partial[((long long)s * 2 + 1) * C + c];
  }
  if (out0 != nullptr) out0[c] = s0;
  if (out1 != nullptr) out1[c] = s1;
}
// final-layer parameter gradients from A[c] = sum_r dlogit_r xhat_rc and S = sum_r dlogit_r:
//   d w_out = gamma * A + beta * S,  d gamma_f = w_out * A,  d beta_f = w_out * S,  d b_out = S
__global__ void __launch_bounds__(256)
final_param_grads_kernel(const float* __restrict__ A, const float* __restrict__ S, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ w_out, int C,
                         float* __restrict__ d_w, float* __restrict__ d_gamma, float* __restrict__ d_beta,
                         float* __restrict__ d_b) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const float s = S[0];        // every column of the "sum of rowscale" vector holds the same total
  d_w[c] = gamma[c] * A[c] + beta[c] * s;
  d_gamma[c] = w_out[c] * A[c];
  d_beta[c] = w_out[c] * s;
  if (c == 0) d_b[0] = s;
}
// loss = sum of the per-row losses, fixed order (one block)
__global__ void __launch_bounds__(1024)
sum_rows_kernel(const float* __restrict__ v, long long n, float* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc += (double)v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sh[0];
}

// ---- GELU on stored pre-activations (erf form, lib/models.py:296 activation="gelu") -----------------------
// element index of the dropout mask = flat index into the [rows, cols] matrix
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ out, long long n2, DropSite d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n2) return;
  const float2 v = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(z)[i]);
  float a = 0.5f * v.x * (1.f + erff(v.x * 0.70710678f)), b = 0.5f * v.y * (1.f + erff(v.y * 0.70710678f));
  if (d.thresh != 0) { a *= drop_factor(d, (uint32_t)(2 * i)); b *= drop_factor(d, (uint32_t)(2 * i + 1)); }
  reinterpret_cast<uint32_t*>(out)[i] = pack_bf16x2(a, b);
}
__device__ __forceinline__ float gelu_grad(float x) {   // Phi(x) + x phi(x)
  return 0.5f * (1.f + erff(x * 0.70710678f)) + x * 0.3989422804f * expf(-0.5f * x * x);
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dm,
                __nv_bfloat16* __restrict__ dz, long long n2, DropSite d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n2) return;
  const float2 v = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(z)[i]);
  float2 g = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(dm)[i]);
  if (d.thresh != 0) { g.x *= drop_factor(d, (uint32_t)(2 * i)); g.y *= drop_factor(d, (uint32_t)(2 * i + 1)); }
  reinterpret_cast<uint32_t*>(dz)[i] = pack_bf16x2(g.x * gelu_grad(v.x), g.y * gelu_grad(v.y));
}

// x_out = x_in + dropout(y)  (residual branch with the TransformerEncoderLayer's dropout1 / dropout2), fp32
__global__ void __launch_bounds__(256)
resid_dropout_kernel(const float* __restrict__ x_in, const float* __restrict__ y, float* __restrict__ x_out,
                     long long n4, DropSite d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 a = reinterpret_cast<const float4*>(x_in)[i];
  const float4 b = reinterpret_cast<const float4*>(y)[i];
  const uint32_t e = (uint32_t)(4 * i);
  reinterpret_cast<float4*>(x_out)[i] =
      make_float4(fmaf(b.x, drop_factor(d, e), a.x), fmaf(b.y, drop_factor(d, e + 1), a.y),
                  fmaf(b.z, drop_factor(d, e + 2), a.z), fmaf(b.w, drop_factor(d, e + 3), a.w));
}
// dst_bf16 = dropout-mask(src) (the gradient entering a dropped branch), or a plain gather + dropout in fp32
__global__ void __launch_bounds__(256)
mask_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4, DropSite d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 a = reinterpret_cast<const float4*>(src)[i];
  const uint32_t e = (uint32_t)(4 * i);
  uint2 u;
  u.x = pack_bf16x2(a.x * drop_factor(d, e), a.y * drop_factor(d, e + 1));
  u.y = pack_bf16x2(a.z * drop_factor(d, e + 2), a.w * drop_factor(d, e + 3));
  reinterpret_cast<uint2*>(dst)[i] = u;
}
__global__ void __launch_bounds__(256)
gather_dropout_kernel(const float* __restrict__ src, long long batch_stride, int T, int C, float* __restrict__ dst,
                      long long n4, DropSite d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const long long per_b = (long long)T * C / 4;
  const long long b = i / per_b, r = i - b * per_b;
  const float4 a = *reinterpret_cast<const float4*>(src + b * batch_stride + r * 4);
  const uint32_t e = (uint32_t)(4 * i);
  reinterpret_cast<float4*>(dst)[i] = make_float4(a.x * drop_factor(d, e), a.y * drop_factor(d, e + 1),
                                                   a.z * drop_factor(d, e + 2), a.w * drop_factor(d, e + 3));
}

// ---- bf16 transpose with zero padding: dst[c, r] = src[r, c] for r < rows, 0 for rows <= r < rows_pad ------
// 64 x 64 tiles through shared memory, 16-byte global loads and stores (cols, ld_src multiples of 8; rows_pad a
// multiple of 64; 16-byte aligned bases — the launcher checks)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, long long rows, int cols,
                      __nv_bfloat16* __restrict__ dst, long long rows_pad) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][72];      // row stride 144 B: 16-byte aligned, conflict-light
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
#pragma unroll
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int r = i >> 3, k = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < rows && c0 + 8 * k < cols) v = *reinterpret_cast<const uint4*>(src + (r0 + r) * ld_src + c0 + 8 * k);
    *reinterpret_cast<uint4*>(&tile[r][8 * k]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int c = i & 63, k = i >> 6;                       // consecutive threads: consecutive columns (smem rows
    if (c0 + c >= cols) continue;                           // read along a column, two threads per bank word)
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = tile[8 * k + e][c];
    *reinterpret_cast<uint4*>(dst + (long long)(c0 + c) * rows_pad + r0 + 8 * k) = *reinterpret_cast<const uint4*>(o);
  }
}
__global__ void __launch_bounds__(256)
copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n4) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
}

}  // namespace

int head_loss_backward_launch(const float* x2, int B, int R, const float* gamma, const float* beta, float eps,
                              const float* w_out, const float* b_out, const int32_t* out_len, const float* target,
                              float pos_weight, float* dx2, __nv_bfloat16* dx2_bf, float* dlogit, float2* stats,
                              float* logits, float* loss_rows, float* loss, cudaStream_t s) {
  const long long rows = (long long)B * R;
  if (rows <= 0) return 0;
  ProfScope ps(s, "train.loss_bwd");
  head_loss_backward_kernel<<<blocks_for_t(rows, 8), 256, 0, s>>>(x2, rows, R, gamma, beta, eps, w_out, b_out, out_len,
                                                                  target, pos_weight, 1.f / (float)B, dx2, dx2_bf,
                                                                  dlogit, stats, logits, loss_rows);
  W2V_CHECK_LAUNCH();
  sum_rows_kernel<<<1, 1024, 0, s>>>(loss_rows, rows, loss);
  W2V_CHECK_LAUNCH();
  return 0;
}

int layernorm_bwd_launch(const float* x, const __nv_bfloat16* dy, int64_t rows, const float* gamma, float eps,
                         const float* dx_in, float* dx_out, __nv_bfloat16* dx_bf, float2* stats, cudaStream_t s) {
  if (rows <= 0) return 0;
  ProfScope ps(s, "train.ln_bwd");
  layernorm_bwd_kernel<<<blocks_for_t(rows, 8), 256, 0, s>>>(x, dy, rows, gamma, eps, dx_in, dx_out, dx_bf, stats);
  W2V_CHECK_LAUNCH();
  return 0;
}

static int colreduce(const void* a, bool a_bf16, int64_t lda, const float* x, const float2* stats,
                     const float* rowscale, int64_t rows, int C, int mode, float* scratch, size_t scratch_floats,
                     float* out0, float* out1, cudaStream_t s) {
  if (rows <= 0 || C <= 0) return 0;
  int slabs = (int)std::min<long long>(256, (rows + 63) / 64);
  slabs = (int)std::min<long long>(slabs, (long long)(scratch_floats / ((size_t)2 * C)));
  W2V_REQUIRE(slabs >= 1, "colreduce: scratch of %zu floats too small for C=%d", scratch_floats, C);
  dim3 grid(blocks_for_t(C, 256), slabs);
  ProfScope ps(s, "train.colreduce");
  if (a_bf16)
    colreduce_partial_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)a, lda, x, stats, rowscale,
                                                                 rows, C, mode, scratch);
  else
    colreduce_partial_kernel<float><<<grid, 256, 0, s>>>((const float*)a, lda, x, stats, rowscale, rows, C, mode,
                                                         scratch);
  W2V_CHECK_LAUNCH();
  colreduce_final_kernel<<<blocks_for_t(C, 256), 256, 0, s>>>(scratch, C, slabs, out0, out1);
  W2V_CHECK_LAUNCH();
  return 0;
}

int colsum_launch(const void* a, bool a_bf16, int64_t lda, int64_t rows, int C, float* scratch,
                  size_t scratch_floats, float* out, cudaStream_t s) {
  return colreduce(a, a_bf16, lda, nullptr, nullptr, nullptr, rows, C, 0, scratch, scratch_floats, out, nullptr, s);
}
int ln_param_grads_launch(const __nv_bfloat16* dy, const float* x, const float2* stats, int64_t rows, int C,
                          float* scratch, size_t scratch_floats, float* d_gamma, float* d_beta, cudaStream_t s) {
  return colreduce(dy, true, C, x, stats, nullptr, rows, C, 1, scratch, scratch_floats, d_gamma, d_beta, s);
}
int final_param_grads_launch(const float* dlogit, const float* x2, const float2* stats, int64_t rows, int C,
                             const float* gamma, const float* beta, const float* w_out, float* scratch,
                             size_t scratch_floats, float* tmpA, float* tmpS, float* d_w, float* d_gamma,
                             float* d_beta, float* d_b, cudaStream_t s) {
  W2V_TRY(colreduce(nullptr, false, 0, x2, stats, dlogit, rows, C, 2, scratch, scratch_floats, tmpA, tmpS, s));
  final_param_grads_kernel<<<blocks_for_t(C, 256), 256, 0, s>>>(tmpA, tmpS, gamma, beta, w_out, C, d_w, d_gamma,
                                                                d_beta, d_b);
  W2V_CHECK_LAUNCH();
  return 0;
}

int gelu_fwd_launch(const __nv_bfloat16* z, __nv_bfloat16* out, int64_t n, const DropSite& d, cudaStream_t s) {
  if (n <= 0) return 0;
  ProfScope ps(s, "train.gelu");
  gelu_fwd_kernel<<<blocks_for_t(n / 2, 256), 256, 0, s>>>(z, out, n / 2, d);
  W2V_CHECK_LAUNCH();
  return 0;
}
int gelu_bwd_launch(const __nv_bfloat16* z, const __nv_bfloat16* dm, __nv_bfloat16* dz, int64_t n, const DropSite& d,
                    cudaStream_t s) {
  if (n <= 0) return 0;
  ProfScope ps(s, "train.gelu_bwd");
  gelu_bwd_kernel<<<blocks_for_t(n / 2, 256), 256, 0, s>>>(z, dm, dz, n / 2, d);
  W2V_CHECK_LAUNCH();
  return 0;
}
int resid_dropout_launch(const float* x_in, const float* y, float* x_out, int64_t n, const DropSite& d, cudaStream_t s) {
  if (n <= 0) return 0;
  ProfScope ps(s, "train.dropout");
  resid_dropout_kernel<<<blocks_for_t(n / 4, 256), 256, 0, s>>>(x_in, y, x_out, n / 4, d);
  W2V_CHECK_LAUNCH();
  return 0;
}
int mask_cast_launch(const float* src, __nv_bfloat16* dst, int64_t n, const DropSite& d, cudaStream_t s) {
  if (n <= 0) return 0;
  ProfScope ps(s, "train.dropout");
  mask_cast_kernel<<<blocks_for_t(n / 4, 256), 256, 0, s>>>(src, dst, n / 4, d);
  W2V_CHECK_LAUNCH();
  return 0;
}
int gather_dropout_launch(const float* src, int64_t batch_stride, int B, int T, int C, float* dst, const DropSite& d,
                          cudaStream_t s) {
  const long long n4 = (long long)B * T * C / 4;
  if (n4 <= 0) return 0;
  ProfScope ps(s, "train.dropout");
  gather_dropout_kernel<<<blocks_for_t(n4, 256), 256, 0, s>>>(src, batch_stride, T, C, dst, n4, d);
  W2V_CHECK_LAUNCH();
  return 0;
}
int transpose_bf16_launch(const __nv_bfloat16* src, int64_t ld_src, int64_t rows, int cols, __nv_bfloat16* dst,
                          int64_t rows_pad, cudaStream_t s) {
  if (rows_pad <= 0 || cols <= 0) return 0;
  W2V_REQUIRE(cols % 8 == 0 && ld_src % 8 == 0 && rows_pad % 64 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "transpose: cols %d / ld %lld / rows_pad %lld must be multiples of 8 / 8 / 64, 16-byte aligned", cols,
              (long long)ld_src, (long long)rows_pad);
  dim3 grid(blocks_for_t(rows_pad, 64), blocks_for_t(cols, 64));
  ProfScope ps(s, "train.transpose");
  transpose_bf16_kernel<<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, rows_pad);
  W2V_CHECK_LAUNCH();
  return 0;
}
int copy_f32_launch(const float* src, float* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  copy_f32_kernel<<<blocks_for_t(n / 4, 256), 256, 0, s>>>(src, dst, n / 4);
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
