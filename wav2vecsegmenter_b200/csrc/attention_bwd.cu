// Backward of the head's multi-head self-attention (torch.nn.MultiheadAttention inside the
// TransformerEncoderLayer of SegmentationFrameClassifier, reference lib/models.py:291-300), used by the
// head-only training step (frozen encoder, reference train.py:381-480). Flash-style: S / P are recomputed
// tile by tile from Q, K and the forward's row log-sum-exp; nothing of size T x T is ever stored.
//
//   P   = exp2(S * scale_log2 - lse)                    S = Q K^T (raw scores), lse from the forward
//   dV  = P^T dO
//   dP  = dO V^T
//   dS  = P * (dP - delta) * scale                      delta_i = sum_d dO_id O_id
//   dQ  = dS K          dK = dS^T Q
// With dropout on the attention weights (mask M in {0, 1/(1-p)}, regenerated from dropout.cuh's hash):
//   O = (P . M) V   ->   dV = (P . M)^T dO,   dP = (dO V^T) . M,   delta unchanged (sum_j P_ij dP_ij = dO_i . O_i).
//
// Three launches of one kernel template, each accumulating ONE gradient in registers (dQ per query tile with
// the key tiles streaming; dV and dK per key tile with the query tiles streaming) — no atomics, deterministic,
// at the price of recomputing S (the head is 2.4 % of the model's FLOPs). Warp-level mma.sync (m16n8k16) with
// the operand patterns of attention.cu; all five matrix products map onto its two ldmatrix patterns:
// "row-major rows as the n dimension" (K in Q K^T) and ".trans: rows as the k dimension" (V in P V).
#include "kernels.cuh"
#include "ptx.cuh"
#include "attention_mma.cuh"

namespace w2v {

namespace {

enum BwdMode : int { BWD_DQ = 0, BWD_DV = 1, BWD_DK = 2 };

// delta[b, h, r] = sum_d dctx[b*R + r, h*DH + d] * ctx[b*R + r, h*DH + d]; one warp per (row, head)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx, int B, int R,
                  int heads, int DH, float* __restrict__ delta) {
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * R * heads) return;
  const int head = (int)(w % heads);
  const long long row = w / heads;
  const long long off = row * (long long)(heads * DH) + head * DH;
  float acc = 0.f;
  for (int d = lane * 2; d < DH; d += 64) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ctx + off + d));
    const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dctx + off + d));
    acc = fmaf(a.x, g.x, fmaf(a.y, g.y, acc));
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const long long b = row / R, r = row - b * R;
    delta[(b * heads + head) * R + r] = acc;
  }
}

// A fragments (16 rows x DH) of this warp's rows of a [64][DH] swizzled tile
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t tile, int warp, int lane, uint32_t (&f)[DH / 16][4]) {
  const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk)
    ldsm_x4(tile + tile_off<DH>(r, kk * 2 + (lane >> 4)), f[kk][0], f[kk][1], f[kk][2], f[kk][3]);
}
// acc[16 x 64] = A[16 x DH] * T[64 x DH]^T  (the rows of the tile T are the n dimension)
template <int DH>
__device__ __forceinline__ void mma_rows_as_n(float (&acc)[8][4], const uint32_t (&a)[DH / 16][4], uint32_t tile,
                                              int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int mi = lane >> 3;
      const int row = np * 16 + (lane & 7) + (mi >> 1) * 8;
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile + tile_off<DH>(row, kk * 2 + (mi & 1)), b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * np], a[kk], b0, b1);
      mma_bf16_16816(acc[2 * np + 1], a[kk], b2, b3);
    }
  }
}
// out[16 x DH] += P[16 x 64] * T[64 x DH]  (the rows of the tile T are the k dimension)
template <int DH>
__device__ __forceinline__ void mma_rows_as_k(float (&out)[DH / 8][4], const uint32_t (&p)[4][4], uint32_t tile,
                                              int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < DH / 16; ++dp) {
      const int mi = lane >> 3;
      const int row = kk * 16 + (lane & 7) + (mi & 1) * 8;
      uint32_t b0, b1, b2, b3;
      ldsm_x4_trans(tile + tile_off<DH>(row, dp * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma_bf16_16816(out[2 * dp], p[kk], b0, b1);
      mma_bf16_16816(out[2 * dp + 1], p[kk], b2, b3);
    }
  }
}
__device__ __forceinline__ void pack_frags(const float (&s)[8][4], uint32_t (&p)[4][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i >> 1][(i & 1) * 2 + 0] = pack_bf16x2(s[i][0], s[i][1]);
    p[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(s[i][2], s[i][3]);
  }
}

// MODE BWD_DQ: the CTA owns 64 QUERY rows (fixed tiles: Q, dO), streams the key tiles (K, V).
// MODE BWD_DV / BWD_DK: the CTA owns 64 KEY rows (fixed tiles: K, V), streams the query tiles (Q, dO).
// Shared memory: fixed A | fixed B | 2 x (stream A | stream B), each 64 x DH bf16; then lse / delta of the
// streamed query tile (key-owner modes).
template <int DH, int MODE>
__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dctx,
                     const float* __restrict__ lse, const float* __restrict__ delta, int R, int heads,
                     const int* __restrict__ kv_len, float scale, float scale_log2,
                     __nv_bfloat16* __restrict__ dqkv, DropSite drop) {
  extern __shared__ __align__(128) uint8_t bwd_smem[];
  constexpr int TILE_BYTES = 64 * DH * 2;
  constexpr int ONT = DH / 8;
  const uint32_t sFixA = smem_u32(bwd_smem);                 // DQ: Q      | DV/DK: K
  const uint32_t sFixB = sFixA + TILE_BYTES;                 // DQ: dO     | DV/DK: V
  const uint32_t sStrA = sFixB + TILE_BYTES;                 // DQ: K x 2  | DV/DK: Q x 2
  const uint32_t sStrB = sStrA + 2 * TILE_BYTES;             // DQ: V x 2  | DV/DK: dO x 2
  float* sL = reinterpret_cast<float*>(bwd_smem + 6 * TILE_BYTES);   // [2][64] lse of the streamed queries
  float* sD = sL + 128;                                              // [2][64] delta

  const int tile0 = blockIdx.x * 64;       // first owned row (query rows for DQ, key rows otherwise)
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int D = heads * DH;
  const long long ld = 3LL * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klen = min(kv_len[b], R);
  const __nv_bfloat16* gQ = qkv + (long long)b * R * ld + head * DH;
  const __nv_bfloat16* gK = gQ + D;
  const __nv_bfloat16* gV = gQ + 2 * D;
  const __nv_bfloat16* gDO = dctx + (long long)b * R * D + head * DH;
  const float* gL = lse + ((long long)b * heads + head) * R;
  const float* gD = delta + ((long long)b * heads + head) * R;

  // streamed dimension: key tiles for DQ (only the valid keys), query tiles otherwise (all R rows are queries)
  const int n_stream = MODE == BWD_DQ ? (klen + 63) / 64 : (R + 63) / 64;
  auto load_stream = [&](int j, int buf) {
    if (MODE == BWD_DQ) {
      load_tile<DH>(sStrA + buf * TILE_BYTES, gK, ld, j * 64, klen);
      load_tile<DH>(sStrB + buf * TILE_BYTES, gV, ld, j * 64, klen);
    } else {
      load_tile<DH>(sStrA + buf * TILE_BYTES, gQ, ld, j * 64, R);
      load_tile<DH>(sStrB + buf * TILE_BYTES, gDO, (long long)D, j * 64, R);
      if (threadIdx.x < 64) {
        const int q = j * 64 + threadIdx.x;
        sL[buf * 64 + threadIdx.x] = q < R ? gL[q] : INFINITY;      // +inf: p = 0 for rows past the window
        sD[buf * 64 + threadIdx.x] = q < R ? gD[q] : 0.f;
      }
    }
  };
  if (MODE == BWD_DQ) {
    load_tile<DH>(sFixA, gQ, ld, tile0, R);
    load_tile<DH>(sFixB, gDO, (long long)D, tile0, R);
  } else {
    load_tile<DH>(sFixA, gK, ld, tile0, klen);
    load_tile<DH>(sFixB, gV, ld, tile0, klen);
  }
  if (n_stream > 0) load_stream(0, 0);
  cp_async_commit();

  float acc[ONT][4];
#pragma unroll
  for (int i = 0; i < ONT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  uint32_t fa[DH / 16][4];                   // A fragments of the fixed tile A (Q or K rows of this warp)
  // this thread's rows of the 16-row slice: g and g + 8
  const int row_lo = tile0 + warp * 16 + (lane >> 2);
  float l_row[2] = {INFINITY, INFINITY}, d_row[2] = {0.f, 0.f};
  if (MODE == BWD_DQ) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (row_lo + 8 * r < R) { l_row[r] = gL[row_lo + 8 * r]; d_row[r] = gD[row_lo + 8 * r]; }
  }
  const bool own_valid[2] = {MODE == BWD_DQ ? true : row_lo < klen, MODE == BWD_DQ ? true : row_lo + 8 < klen};
  bool frags_loaded = false;

  for (int j = 0; j < n_stream; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_stream) {
      load_stream(j + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (!frags_loaded) {
      load_a_frags<DH>(sFixA, warp, lane, fa);
      frags_loaded = true;
    }
    const uint32_t tA = sStrA + buf * TILE_BYTES, tB = sStrB + buf * TILE_BYTES;
    float s[8][4];
    mma_rows_as_n<DH>(s, fa, tA, lane);      // DQ: Q K^T [queries x keys] | DV/DK: K Q^T [keys x queries]
    // P (or P^T): exp2(s * scale_log2 - lse[query]); masked keys and rows past the window give 0
    const int col0 = j * 64 + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = col0 + 8 * i + (e & 1);
        float lq;
        bool ok;
        if (MODE == BWD_DQ) { lq = l_row[e >> 1]; ok = col < klen; }
        else { lq = sL[buf * 64 + 8 * i + 2 * (lane & 3) + (e & 1)]; ok = own_valid[e >> 1]; }
        s[i][e] = ok ? exp2f(fmaf(s[i][e], scale_log2, -lq)) : 0.f;
      }
    }
    // dropout factor of element (i, e) of this thread's fragment: rows = owned rows, columns = streamed rows
    const uint32_t bh = (uint32_t)(b * heads + head) * (uint32_t)R;
    auto mask_of = [&](int i, int e) -> float {
      const uint32_t own = (uint32_t)(row_lo + 8 * (e >> 1)), str = (uint32_t)(col0 + 8 * i + (e & 1));
      const uint32_t q = MODE == BWD_DQ ? own : str, k = MODE == BWD_DQ ? str : own;
      return drop_factor(drop, (bh + q) * (uint32_t)R + k);
    };
    uint32_t pf[4][4];
    if (MODE == BWD_DV) {
      if (drop.thresh != 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[i][e] *= mask_of(i, e);
      }
      pack_frags(s, pf);
      mma_rows_as_k<DH>(acc, pf, tB, lane);  // dV += P^T dO
    } else {
      // dP (or dP^T) = fixed-B rows x streamed-B rows: DQ: dO V^T | DK: V dO^T
      uint32_t fb[DH / 16][4];
      load_a_frags<DH>(sFixB, warp, lane, fb);
      float dp[8][4];
      mma_rows_as_n<DH>(dp, fb, tB, lane);
      if (drop.thresh != 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) dp[i][e] *= mask_of(i, e);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dl = MODE == BWD_DQ ? d_row[e >> 1] : sD[buf * 64 + 8 * i + 2 * (lane & 3) + (e & 1)];
          s[i][e] = s[i][e] * (dp[i][e] - dl) * scale;           // dS
        }
      }
      pack_frags(s, pf);
      mma_rows_as_k<DH>(acc, pf, tA, lane);  // DQ: dQ += dS K | DK: dK += dS^T Q
    }
    __syncthreads();   // everyone done with buf before it is refilled two iterations later
  }
  if (n_stream == 0) {
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- gradient tile -> bf16 -> this warp's rows of the fixed tile A (retired) -> coalesced 16-byte stores
  __syncwarp();
  const int r_lo = warp * 16 + (lane >> 2);
#pragma unroll
  for (int i = 0; i < ONT; ++i) {
    const int col = i * 8 + 2 * (lane & 3);
    const uint32_t a0 = sFixA + tile_off<DH>(r_lo, col >> 3) + (col & 7) * 2;
    const uint32_t a1 = sFixA + tile_off<DH>(r_lo + 8, col >> 3) + (col & 7) * 2;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(pack_bf16x2(acc[i][0], acc[i][1])) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(pack_bf16x2(acc[i][2], acc[i][3])) : "memory");
  }
  __syncwarp();
  constexpr int CHUNKS = DH / 8;
  __nv_bfloat16* gOut = dqkv + (long long)b * R * ld + head * DH + (MODE == BWD_DQ ? 0 : MODE == BWD_DK ? D : 2 * D);
  for (int i = lane; i < 16 * CHUNKS; i += 32) {
    const int r = warp * 16 + i / CHUNKS, c = i % CHUNKS;
    if (tile0 + r < R) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(sFixA + tile_off<DH>(r, c)));
      *reinterpret_cast<uint4*>(gOut + (long long)(tile0 + r) * ld + c * 8) = v;
    }
  }
}

template <int DH, int MODE>
int launch_mode(const __nv_bfloat16* qkv, const __nv_bfloat16* dctx, const float* lse, const float* delta, int B,
                int R, int heads, const int32_t* kv_len, float scale, __nv_bfloat16* dqkv, const DropSite& drop,
                cudaStream_t s) {
  const int smem = 6 * 64 * DH * 2 + 4 * 64 * (int)sizeof(float);
  W2V_ONCE_BEGIN
  W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<DH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  W2V_ONCE_END
  dim3 grid((R + 63) / 64, heads, B);
  ProfScope ps(s, MODE == BWD_DQ ? "attention_bwd_dq" : MODE == BWD_DV ? "attention_bwd_dv" : "attention_bwd_dk");
  attention_bwd_kernel<DH, MODE><<<grid, ATT_THREADS, smem, s>>>(qkv, dctx, lse, delta, R, heads, kv_len, scale,
                                                                  scale * 1.4426950408889634f, dqkv, drop);
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int attention_bwd_launch(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                         const float* lse, float* delta, int B, int R, int heads, int head_dim,
                         const int32_t* kv_len, float scale, __nv_bfloat16* dqkv, const DropSite& drop,
                         cudaStream_t s) {
  if (B <= 0 || R <= 0) return 0;
  W2V_REQUIRE(head_dim == 64 || head_dim == 128, "attention_bwd: head_dim %d unsupported (64 / 128)", head_dim);
  const long long warps = (long long)B * R * heads;
  attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(ctx, dctx, B, R, heads, head_dim, delta);
  W2V_CHECK_LAUNCH();
  if (head_dim == 64) {
    W2V_TRY((launch_mode<64, BWD_DQ>(qkv, dctx, lse, delta, B, R, heads, kv_len, scale, dqkv, drop, s)));
    W2V_TRY((launch_mode<64, BWD_DV>(qkv, dctx, lse, delta, B, R, heads, kv_len, scale, dqkv, drop, s)));
    W2V_TRY((launch_mode<64, BWD_DK>(qkv, dctx, lse, delta, B, R, heads, kv_len, scale, dqkv, drop, s)));
  } else {
    W2V_TRY((launch_mode<128, BWD_DQ>(qkv, dctx, lse, delta, B, R, heads, kv_len, scale, dqkv, drop, s)));
    W2V_TRY((launch_mode<128, BWD_DV>(qkv, dctx, lse, delta, B, R, heads, kv_len, scale, dqkv, drop, s)));
    W2V_TRY((launch_mode<128, BWD_DK>(qkv, dctx, lse, delta, B, R, heads, kv_len, scale, dqkv, drop, s)));
  }
  return 0;
}

}  // namespace w2v
