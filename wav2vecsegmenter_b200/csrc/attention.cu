// Fused non-causal multi-head attention with a per-window key-length mask (flash-style: S and P
// never leave the SM; online softmax in fp32).
//
// Replaces Wav2Vec2Attention's softmax(QK^T*scale + mask)V (HF:500-549, 16 heads x 64) and the
// nn.MultiheadAttention inside the head's TransformerEncoderLayer (lib/models.py:291-300, 8 x 128).
//
// Round-1 implementation: warp-level mma.sync (m16n8k16 bf16, fp32 accumulate) with ldmatrix
// operand fetch from XOR-swizzled shared memory and cp.async double buffering of K/V tiles.
// A tcgen05/TMEM version (S and O in TMEM) is the planned replacement; this one defines the
// numerics and is the parity baseline for it.
//
// Layout: qkv bf16 [B*R, 3*D], D = heads*DH; Q | K | V column blocks. Window b owns rows
// [b*R, (b+1)*R); keys t >= kv_len[b] are masked. All R rows are produced as queries.
#include "kernels.cuh"
#include "ptx.cuh"
#include "attention_mma.cuh"

namespace w2v {

namespace {

// DROP (training forward only): the attention weights that multiply V are masked and rescaled, the row sum
// (and the log-sum-exp the backward uses) stays that of the full softmax — torch's dropout(softmax(S)) V.
template <int DH, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, int R, int heads,
                 const int* __restrict__ kv_len, float scale_log2,
                 __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse, DropSite drop) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  constexpr int TILE_BYTES = ATT_BKV * DH * 2;
  const uint32_t sQ = smem_u32(att_smem);
  const uint32_t sK = sQ + TILE_BYTES;           // 2 buffers
  const uint32_t sV = sK + 2 * TILE_BYTES;       // 2 buffers

  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int D = heads * DH;
  const long long ld = 3LL * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klen = min(kv_len[b], R);
  const int n_tiles = (klen + ATT_BKV - 1) / ATT_BKV;

  const __nv_bfloat16* gQ = qkv + (long long)b * R * ld + head * DH;
  const __nv_bfloat16* gK = gQ + D;
  const __nv_bfloat16* gV = gQ + 2 * D;

  load_tile<DH>(sQ, gQ, ld, q0, R);
  if (n_tiles > 0) {
    load_tile<DH>(sK, gK, ld, 0, klen);
    load_tile<DH>(sV, gV, ld, 0, klen);
  }
  cp_async_commit();

  constexpr int KSTEPS = DH / 16;   // k-steps of Q K^T
  constexpr int ONT = DH / 8;       // n-tiles of O
  float o[ONT][4];
#pragma unroll
  for (int i = 0; i < ONT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[KSTEPS][4];
  bool q_loaded = false;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_tiles) {
      load_tile<DH>(sK + (buf ^ 1) * TILE_BYTES, gK, ld, (j + 1) * ATT_BKV, klen);
      load_tile<DH>(sV + (buf ^ 1) * TILE_BYTES, gV, ld, (j + 1) * ATT_BKV, klen);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (!q_loaded) {
      // A fragments of this warp's 16 query rows, all k-steps (kept in registers for the CTA life)
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk)
        ldsm_x4(sQ + tile_off<DH>(r, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2],
                qf[kk][3]);
      q_loaded = true;
    }

    // ---- S = Q K^T for 64 keys: 8 n-tiles
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    const uint32_t kb = sK + buf * TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of n-tiles (16 keys)
        const int mi = lane >> 3;
        const int krow = np * 16 + (lane & 7) + (mi >> 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + tile_off<DH>(krow, kk * 2 + (mi & 1)), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], qf[kk], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[kk], b2, b3);
      }
    }

    // ---- mask keys beyond klen, online softmax (rows lane/4 and lane/4+8)
    const int key0 = j * ATT_BKV + 2 * (lane & 3);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = key0 + i * 8;
      if (k >= klen) { s[i][0] = -INFINITY; s[i][2] = -INFINITY; }
      if (k + 1 >= klen) { s[i][1] = -INFINITY; s[i][3] = -INFINITY; }
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
    float alpha[2], moff[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f((m_run[r] - m_new) * scale_log2);
      moff[r] = (m_new == -INFINITY) ? 0.f : m_new * scale_log2;
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float p0 = exp2f(fmaf(s[i][0], scale_log2, -moff[0]));
      float p1 = exp2f(fmaf(s[i][1], scale_log2, -moff[0]));
      float p2 = exp2f(fmaf(s[i][2], scale_log2, -moff[1]));
      float p3 = exp2f(fmaf(s[i][3], scale_log2, -moff[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      if (DROP) {
        const uint32_t q_lo = (uint32_t)(q0 + warp * 16 + (lane >> 2));
        const uint32_t e0 = ((uint32_t)(b * heads + head) * (uint32_t)R + q_lo) * (uint32_t)R + (uint32_t)(key0 + i * 8);
        const uint32_t e1 = e0 + 8u * (uint32_t)R;
        p0 *= drop_factor(drop, e0); p1 *= drop_factor(drop, e0 + 1);
        p2 *= drop_factor(drop, e1); p3 *= drop_factor(drop, e1 + 1);
      }
      pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int i = 0; i < ONT; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }

    // ---- O += P V
    const uint32_t vb = sV + buf * TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < ONT / 2; ++dp) {  // pairs of d n-tiles
        const int mi = lane >> 3;
        const int vrow = kk * 16 + (lane & 7) + (mi & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(vb + tile_off<DH>(vrow, dp * 2 + (mi >> 1)), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * dp], pf[kk], b0, b1);
        mma_bf16_16816(o[2 * dp + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  if (n_tiles == 0) {
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- finalise: O / l, stage through this warp's rows of sQ, coalesced 16-byte stores
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    inv[r] = l > 0.f ? 1.f / l : 0.f;
    // training forward: row log-sum-exp in the log2 domain the backward re-exponentiates with,
    // p = exp2(s * scale_log2 - lse); +inf for a row without any valid key (p = 0 everywhere)
    if (lse != nullptr && (lane & 3) == 0) {
      const int row = q0 + warp * 16 + (lane >> 2) + 8 * r;
      if (row < R)
        lse[((long long)b * heads + head) * R + row] = l > 0.f ? m_run[r] * scale_log2 + log2f(l) : INFINITY;
    }
  }
  __syncwarp();
  const int r_lo = warp * 16 + (lane >> 2);
#pragma unroll
  for (int i = 0; i < ONT; ++i) {
    const int col = i * 8 + 2 * (lane & 3);
    const uint32_t a0 = sQ + tile_off<DH>(r_lo, col >> 3) + (col & 7) * 2;
    const uint32_t a1 = sQ + tile_off<DH>(r_lo + 8, col >> 3) + (col & 7) * 2;
    const uint32_t v0 = pack_bf16x2(o[i][0] * inv[0], o[i][1] * inv[0]);
    const uint32_t v1 = pack_bf16x2(o[i][2] * inv[1], o[i][3] * inv[1]);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(v0) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(v1) : "memory");
  }
  __syncwarp();
  constexpr int CHUNKS = DH / 8;
  __nv_bfloat16* gO = ctx + (long long)b * R * D + head * DH;
  for (int i = lane; i < 16 * CHUNKS; i += 32) {
    const int r = warp * 16 + i / CHUNKS, c = i % CHUNKS;
    if (q0 + r < R) {
      uint4 v;
      const uint32_t a = sQ + tile_off<DH>(r, c);
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(a));
      *reinterpret_cast<uint4*>(gO + (long long)(q0 + r) * D + c * 8) = v;
    }
  }
}

}  // namespace

int attention_launch(const __nv_bfloat16* qkv, int B, int R, int heads, int head_dim,
                     const int32_t* kv_len, float scale, __nv_bfloat16* ctx, cudaStream_t s, float* lse,
                     const DropSite* drop) {
  if (B <= 0 || R <= 0) return 0;
  W2V_REQUIRE(head_dim == 64 || head_dim == 128, "attention: head_dim %d unsupported (64 / 128)",
              head_dim);
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((R + ATT_BQ - 1) / ATT_BQ, heads, B);
  const int smem = 5 * ATT_BKV * head_dim * 2;
  ProfScope ps(s, head_dim == 64 ? "attention_d64" : "attention_d128");
  const bool dropping = drop != nullptr && drop->thresh != 0;
  const DropSite d = dropping ? *drop : DropSite{0, 0, 1.f};
  if (head_dim == 64) {
    if (dropping) attention_kernel<64, true><<<grid, ATT_THREADS, smem, s>>>(qkv, R, heads, kv_len, scale_log2, ctx, lse, d);
    else attention_kernel<64, false><<<grid, ATT_THREADS, smem, s>>>(qkv, R, heads, kv_len, scale_log2, ctx, lse, d);
  } else {
    W2V_ONCE_BEGIN
      W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<128, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<128, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    W2V_ONCE_END
    if (dropping) attention_kernel<128, true><<<grid, ATT_THREADS, smem, s>>>(qkv, R, heads, kv_len, scale_log2, ctx, lse, d);
    else attention_kernel<128, false><<<grid, ATT_THREADS, smem, s>>>(qkv, R, heads, kv_len, scale_log2, ctx, lse, d);
  }
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
