// Fused non-causal attention on the 5th-gen tensor cores: S = Q K^T and O += P V are tcgen05.mma
// with both accumulators in TMEM; Q/K/V tiles arrive by TMA straight out of the fused QKV
// activation [B*R, 3*D] (no head split / transpose kernels); softmax runs one thread per query
// row on the TMEM lanes (no shuffles), P goes back through 128B-swizzled shared memory as the A
// operand of the second MMA, V is consumed in place as an MN-major B operand.
//
// Replaces Wav2Vec2Attention's softmax(QK^T*scale + key mask) V (HF:500-549, 16 x 64) and the
// nn.MultiheadAttention of the head's TransformerEncoderLayer (lib/models.py:291-300, 8 x 128).
//
// CTA = (128 query rows, one head, one window). 5 warps:
//   warps 0..3  softmax: thread r owns query row r == TMEM lane r. Per 128-key tile: tcgen05.ld the
//               S row into registers, release S, row max, (rare) rescale of O in TMEM, exp2,
//               row sum, bf16 P row -> swizzled smem, fence.proxy.async, arrive.
//   warp 4      one lane: TMA producer (Q once, K/V double-buffered) AND MMA issuer. S(j+1) is
//               issued as soon as the softmax threads hold S(j) in registers, so QK^T of the next
//               tile and PV of the current one overlap the exponentials (the MUFU-bound part).
// Rescaling is lazy (as in FlashAttention-4): the running max used in the exponent is only
// advanced when the tile max exceeds it by more than 2^8, so O is rarely touched after tile 0.
#include <cstdlib>
#include <cstring>
#include "kernels.cuh"
#include "ptx.cuh"

namespace w2v {

namespace {

constexpr int AT_BM = 128;        // query rows per CTA
constexpr int AT_BN = 128;        // keys per tile
constexpr int AT_THREADS = 160;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int DH>
struct AttCfg {
  static constexpr int SUB = DH / 64;                    // 64-column (128-byte) sub-tiles per row
  static constexpr int TILE_BYTES = AT_BN * DH * 2;      // one Q / K / V tile
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;      // 32 KB
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + TILE_BYTES;       // 2 buffers
  static constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;   // 2 buffers
  static constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  // d=64: no alignment slack, so that two CTAs (2 x (112 KB + 128 B + 1 KB reserved)) fit in one
  // SM's 228 KB; the kernel traps if the dynamic smem base is not 1024-byte aligned.
  static constexpr int SLACK = (DH == 64) ? 0 : 1024;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + SLACK;
  static constexpr int TMEM_COLS = 256;                  // S: 128 columns, O: DH columns
  static constexpr int O_COL = 128;
};

// MN-major, 128B-swizzled B operand (V tile: rows = keys = K dimension, 64 contiguous d per row):
// SBO = 1024 B between 8-key groups, LBO = bytes between consecutive 64-wide d chunks.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, (DH == 64) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, int R, int heads,
                    const int* __restrict__ kv_len, float scale_log2,
                    __nv_bfloat16* __restrict__ ctx) {
  using Cfg = AttCfg<DH>;
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = att_raw;
  if constexpr (Cfg::SLACK > 0) {
    smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_raw) + 1023) &
                                      ~static_cast<uintptr_t>(1023));
  } else if ((smem_u32(att_raw) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("w2vseg: attention smem base not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int D = heads * DH;
  const int klen = min(__ldg(kv_len + b), R);
  const int n_tiles = (klen + AT_BN - 1) / AT_BN;
  const int row_base = b * R;   // first row of this window in the flat [B*R] row space

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      mbar_init(q_full, 1);
      mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
      mbar_init(&v_full[0], 1); mbar_init(&v_full[1], 1);
      mbar_init(s_full, 1);
      mbar_init(s_free, 128);
      mbar_init(p_full, 128);
      mbar_init(pv_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + Cfg::O_COL;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer + MMA issuer
    // every lane runs the (warp-uniform) control flow so that descriptors and addresses live in
    // uniform registers; one elected lane issues the TMA / MMA / commit instructions. (With the
    // whole role under `lane == 0` ptxas moved every operand vector->uniform before each
    // tcgen05.mma: ~80 cycles per MMA instead of ~30, measured with clock64.)
    if (n_tiles > 0) {
      const int qcol = head * DH, kcol = D + head * DH, vcol = 2 * D + head * DH;
      auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col0, int row) {
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SUB; ++s)
            tma_load_2d(dst + s * (AT_BN * 128), &tmap_qkv, bar, col0 + s * 64, row);
        }
        __syncwarp();
      };
      load_tile(sQ, q_full, qcol, row_base + q0);
      load_tile(sK, &k_full[0], kcol, row_base);
      load_tile(sV, &v_full[0], vcol, row_base);
      if (n_tiles > 1) {
        load_tile(sK + Cfg::TILE_BYTES, &k_full[1], kcol, row_base + AT_BN);
        load_tile(sV + Cfg::TILE_BYTES, &v_full[1], vcol, row_base + AT_BN);
      }
      constexpr uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN);
      constexpr uint32_t idesc_o = make_idesc_bf16(AT_BM, DH) | (1u << 16);  // B (=V) is MN-major
      const uint64_t q_desc = make_desc_k_sw128(smem_u32(sQ));
      const uint64_t k_desc0 = make_desc_k_sw128(smem_u32(sK));
      const uint64_t p_desc = make_desc_k_sw128(smem_u32(sP));
      const uint64_t v_desc0 = make_desc_mn_sw128(smem_u32(sV), AT_BN * 128);

      auto issue_s = [&](int j) {
        const uint64_t kd = k_desc0 + (uint64_t)(((j & 1) * Cfg::TILE_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const uint64_t off = (uint64_t)(((kk >> 2) * (AT_BN * 128) + (kk & 3) * 32) >> 4);
            tc_mma_ss(tS, q_desc + off, kd + off, idesc_s, (uint32_t)(kk != 0));
          }
          tc_commit(s_full);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);

      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          mbar_wait(&k_full[(j + 1) & 1], (uint32_t)(((j + 1) >> 1) & 1));
          mbar_wait(s_free, (uint32_t)(j & 1));      // softmax holds S(j) in registers
          tc_fence_after();
          issue_s(j + 1);
          // S(j) has retired (the softmax threads read it), so K buffer j&1 can be refilled
          if (j + 2 < n_tiles)
            load_tile(sK + (j & 1) * Cfg::TILE_BYTES, &k_full[j & 1], kcol, row_base + (j + 2) * AT_BN);
        }
        mbar_wait(&v_full[j & 1], (uint32_t)((j >> 1) & 1));
        mbar_wait(p_full, (uint32_t)(j & 1));        // P(j) in smem, O rescaled if it had to be
        tc_fence_after();
        {
          const uint64_t vd = v_desc0 + (uint64_t)(((j & 1) * Cfg::TILE_BYTES) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < AT_BN / 16; ++kk)
              tc_mma_ss(tO, p_desc + (uint64_t)(((kk >> 2) * (AT_BM * 128) + (kk & 3) * 32) >> 4),
                        vd + (uint64_t)((kk * 16 * 128) >> 4), idesc_o, (uint32_t)((j | kk) != 0));
            tc_commit(pv_done);
          }
          __syncwarp();
        }
        if (j + 2 < n_tiles) {
          mbar_wait(pv_done, (uint32_t)(j & 1));     // V buffer j&1 is free once PV(j) retired
          load_tile(sV + (j & 1) * Cfg::TILE_BYTES, &v_full[j & 1], vcol, row_base + (j + 2) * AT_BN);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax (thread = query row)
    const int r = threadIdx.x;                       // 0..127 == TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_used = 0.f;                              // running max (log2 units) used in exponents
    float l_sum = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, (uint32_t)(j & 1));
      tc_fence_after();
      // ---- pass 1: row max. All four 32-column loads are issued before the single wait; the
      // values are dropped again (S stays in TMEM), which keeps the kernel inside the 168-register
      // budget of two CTAs per SM without spilling.
      float mx;
      {
        uint32_t raw[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + lane_off + (uint32_t)(c * 32), raw[c]);
        tc_wait_ld();
        const int kbase = j * AT_BN;
        float m8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m8[i] = -INFINITY;
        if (kbase + AT_BN > klen) {                  // only the last tile has masked keys
#pragma unroll
          for (int i = 0; i < AT_BN; ++i) {
            const float v = (kbase + i < klen) ? __uint_as_float(raw[i >> 5][i & 31]) : -INFINITY;
            m8[i & 7] = fmaxf(m8[i & 7], v);
          }
        } else {
#pragma unroll
          for (int i = 0; i < AT_BN; ++i) m8[i & 7] = fmaxf(m8[i & 7], __uint_as_float(raw[i >> 5][i & 31]));
        }
        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                   fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7]))) * scale_log2;
      }

      float factor = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = mx;
      } else {
        need = mx > m_used + AT_RESCALE_THRESHOLD;
        if (need) {
          factor = ex2_approx(m_used - mx);
          m_used = mx;
          l_sum *= factor;
        }
        // PV(j-1) must have retired before P is overwritten or O is rescaled
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll
          for (int c = 0; c < DH / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + lane_off + (uint32_t)(c * 32), o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32b_x32(tO + lane_off + (uint32_t)(c * 32), o);
          }
          tc_wait_st();
        }
      }

      // ---- pass 2: p = 2^(s*scale - m) (one FFMA + one MUFU per element), 4 partial sums, bf16,
      // 16-byte chunks into the K-major SW128 layout; the load of chunk c+1 is in flight while
      // chunk c is processed. Masked keys get p = 0 exactly.
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_used;
      const int n_valid = klen - j * AT_BN;          // >= 1; >= 128 except for the last tile
      uint32_t rb[2][32];
      tmem_ld_32x32b_x32(tS + lane_off, rb[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc_wait_ld();
        if (c + 1 < 4) {
          tmem_ld_32x32b_x32(tS + lane_off + (uint32_t)((c + 1) * 32), rb[(c + 1) & 1]);
        } else {
          tc_fence_before();
          mbar_arrive(s_free);                       // S may be overwritten by QK^T of tile j+1
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = c * 32 + g * 8 + i;
            const float e = ex2_approx(fmaf(__uint_as_float(rb[c & 1][g * 8 + i]), scale_log2, neg_m));
            p[i] = (col < n_valid) ? e : 0.f;
            sum4[i & 3] += p[i];
          }
          uint4 u;
          u.x = pack_bf16x2(p[0], p[1]);
          u.y = pack_bf16x2(p[2], p[3]);
          u.z = pack_bf16x2(p[4], p[5]);
          u.w = pack_bf16x2(p[6], p[7]);
          const int ch = c * 4 + g;
          const uint32_t addr = smem_u32(sP) + (uint32_t)((ch >> 3) * (AT_BM * 128) + r * 128 +
                                                          (((ch & 7) ^ (r & 7)) << 4));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y),
                       "r"(u.z), "r"(u.w)
                       : "memory");
        }
      }
      const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      l_sum += sum;
      fence_proxy_async_smem();                      // generic-proxy stores -> visible to the MMA
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> ctx
    const int row = q0 + r;
    __nv_bfloat16* out = ctx + ((long long)(row_base + row)) * D + head * DH;
    if (n_tiles > 0) {
      mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
      tc_fence_after();
      const float inv = 1.f / l_sum;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + lane_off + (uint32_t)(c * 32), o);
        tc_wait_ld();
        if (row < R) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
            *reinterpret_cast<uint4*>(out + c * 32 + i * 8) = u;
          }
        }
      }
    } else if (row < R) {                            // no valid key at all: zeros
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(0, 0, 0, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace

int attention_tc_launch(const __nv_bfloat16* qkv, int B, int R, int heads, int head_dim,
                        const int32_t* kv_len, float scale, __nv_bfloat16* ctx, cudaStream_t s) {
  if (B <= 0 || R <= 0) return 0;
  W2V_REQUIRE(head_dim == 64 || head_dim == 128, "attention: head_dim %d unsupported (64 / 128)",
              head_dim);
  if (head_dim == 64) {
    static const bool use_v15 = [] {
      const char* e = getenv("W2VSEG_ATT64");
      return e != nullptr && strcmp(e, "v15") == 0;
    }();
    if (!use_v15) return attention_tc64_launch(qkv, B, R, heads, kv_len, scale, ctx, s);
  }
  const int D = heads * head_dim;
  CUtensorMap tm;
  W2V_TRY(make_tmap_2d_bf16(&tm, qkv, (uint64_t)3 * D, (uint64_t)B * R, (uint64_t)3 * D, 64, AT_BN));
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((R + AT_BM - 1) / AT_BM, heads, B);
  ProfScope ps(s, head_dim == 64 ? "attention_d64" : "attention_d128");
  if (head_dim == 64) {
    W2V_ONCE_BEGIN
      W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<64>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttCfg<64>::SMEM_BYTES));
    W2V_ONCE_END
    attention_tc_kernel<64><<<grid, AT_THREADS, AttCfg<64>::SMEM_BYTES, s>>>(tm, R, heads, kv_len,
                                                                             scale_log2, ctx);
  } else {
    W2V_ONCE_BEGIN
      W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<128>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttCfg<128>::SMEM_BYTES));
    W2V_ONCE_END
    attention_tc_kernel<128><<<grid, AT_THREADS, AttCfg<128>::SMEM_BYTES, s>>>(tm, R, heads, kv_len,
                                                                               scale_log2, ctx);
  }
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
