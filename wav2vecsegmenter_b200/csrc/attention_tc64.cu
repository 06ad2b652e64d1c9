// Fused non-causal attention for head_dim 64 (the 16 x 64 encoder layers): tcgen05 / TMEM / TMA,
// EIGHT softmax warps per CTA with the S tile split by COLUMNS, each warp keeping its 64 scores
// in registers (one TMEM read per score, exact row max, no second pass).
//
// Why (profiles/attention_r01b.md): ncu on the 4-warp kernel (attention_tc.cu, thread = whole
// 128-key row, two TMEM passes) shows issue slots 49 % and MUFU 50 % busy with the warps stalled on
// the long scoreboard (tcgen05.ld + mbarrier polls): with 2 softmax warps per scheduler nothing
// hides those latencies. A single warp only sustains ~44 B/clk of tcgen05.ld (16 warps reach
// 477 B/clk per SM, experiments/tmem_ld_bw.cu), so the per-warp chain, not the TMEM, is the limit.
// Here two warps share each TMEM lane quarter: warp w (0..3) owns key columns 0..63 of the tile,
// warp w+4 columns 64..127 of the SAME 32 query rows. That halves every per-warp latency chain,
// doubles the warps per scheduler (2 CTAs per SM -> 16 softmax warps) and lets the 64 scores stay
// in registers between the max and the exponentials. The two warps agree on the row max through a
// small fp32 exchange in shared memory and one
// 64-thread named barrier per tile; row sums stay per-warp until the end. P goes to TMEM as packed
// bf16x2 (64 columns) and is the A operand of a TS-form MMA: shared memory only carries Q, K and V
// (with P in shared memory the tile needs ~1150 cycles of shared-memory bandwidth per SM, more
// than the 1024 cycles of MUFU work: measured 121 -> 113 us).
//
//   warps 0..7  softmax (thread = one query row x 64 keys), P -> TMEM (tcgen05.st 32x32b)
//   warp 8      TMA producer (Q once, K/V double-buffered) + MMA issuer (warp-uniform, elect_one)
//   TMEM        S 128 columns fp32 | O 64 columns fp32 | P 64 columns bf16x2   (two CTAs per SM)
// Rescaling is lazy (FlashAttention-4): the exponent's running max only advances when a tile max
// exceeds it by more than 2^8. Replaces Wav2Vec2Attention's softmax(QK^T*scale + key mask) V
// (HF:500-549).
#include "kernels.cuh"
#include "ptx.cuh"

namespace w2v {

namespace {

constexpr int A6_BM = 128;        // query rows per CTA
constexpr int A6_BN = 128;        // keys per tile
constexpr int A6_DH = 64;
constexpr int A6_THREADS = 288;   // 8 softmax warps + 1 TMA/MMA warp
constexpr int A6_CTRL_WARP = 8;
constexpr int A6_TILE_BYTES = A6_BN * A6_DH * 2;       // 16 KB: one Q / K / V tile
constexpr int A6_OFF_Q = 0;
constexpr int A6_OFF_K = A6_OFF_Q + A6_TILE_BYTES;     // 2 buffers
constexpr int A6_OFF_V = A6_OFF_K + 2 * A6_TILE_BYTES; // 2 buffers
constexpr int A6_OFF_BAR = A6_OFF_V + 2 * A6_TILE_BYTES;   // 9 mbarriers + TMEM slot (80 B)
constexpr int A6_OFF_X = A6_OFF_BAR + 128;             // fp32 [2 tile parities][2 halves][128] row max
constexpr int A6_OFF_L = A6_OFF_X + 2048;              // fp32 [2 halves][128] row sums (epilogue)
// the kernel traps if the dynamic smem base is not 1024-byte aligned (no alignment slack)
constexpr int A6_SMEM_BYTES = A6_OFF_L + 1024;
constexpr int A6_TMEM_COLS = 256;
constexpr int A6_O_COL = 128;
constexpr int A6_P_COL = 192;     // P as packed bf16x2: 64 columns = 128 keys (A operand of P V)
constexpr float A6_RESCALE_THRESHOLD = 8.0f;           // log2 units

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void pair_barrier(int id) {   // the two warps of one lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// 32 lanes x 64 consecutive fp32 columns in one instruction (thread i = TMEM lane base + i)
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, float (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]),
        "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]),
        "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]),
        "=f"(r[29]), "=f"(r[30]), "=f"(r[31]), "=f"(r[32]), "=f"(r[33]), "=f"(r[34]), "=f"(r[35]),
        "=f"(r[36]), "=f"(r[37]), "=f"(r[38]), "=f"(r[39]), "=f"(r[40]), "=f"(r[41]), "=f"(r[42]),
        "=f"(r[43]), "=f"(r[44]), "=f"(r[45]), "=f"(r[46]), "=f"(r[47]), "=f"(r[48]), "=f"(r[49]),
        "=f"(r[50]), "=f"(r[51]), "=f"(r[52]), "=f"(r[53]), "=f"(r[54]), "=f"(r[55]), "=f"(r[56]),
        "=f"(r[57]), "=f"(r[58]), "=f"(r[59]), "=f"(r[60]), "=f"(r[61]), "=f"(r[62]), "=f"(r[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
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
// 8-column TMEM load / store: the (rare) rescale of O runs in small chunks so that it does not
// push the 64 live scores out of the register file
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// -DA6_TRACE: clock64 timeline of one CTA (softmax warp 0 and the control warp), printed at exit
#ifdef A6_TRACE
__device__ long long g_trace[2][8][8];   // [role][tile][event]
#define A6_T(role, ev)                                                                     \
  do {                                                                                     \
    if (traced && lane == 0 && j < 8) g_trace[role][j][ev] = clock64() - t_start;          \
  } while (0)
#else
#define A6_T(role, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(A6_THREADS, 2)
attention_tc64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, int R, int heads,
                      const int* __restrict__ kv_len, float scale_log2,
                      __nv_bfloat16* __restrict__ ctx) {
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = att_raw;
  if ((smem_u32(att_raw) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("w2vseg: attention smem base not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + A6_OFF_Q;
  uint8_t* sK = smem + A6_OFF_K;
  uint8_t* sV = smem + A6_OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A6_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xch = reinterpret_cast<float*>(smem + A6_OFF_X);   // [2][2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * A6_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int D = heads * A6_DH;
  const int klen = min(__ldg(kv_len + b), R);
  const int n_tiles = (klen + A6_BN - 1) / A6_BN;
  const int row_base = b * R;   // first row of this window in the flat [B*R] row space
#ifdef A6_TRACE
  const bool traced = blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 7 && (warp == 0 || warp == A6_CTRL_WARP);
  const long long t_start = clock64();
#endif

  if (warp == A6_CTRL_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      mbar_init(q_full, 1);
      mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
      mbar_init(&v_full[0], 1); mbar_init(&v_full[1], 1);
      mbar_init(s_full, 1);
      mbar_init(s_free, 8);      // one elected arrive per softmax warp
      mbar_init(p_full, 8);
      mbar_init(pv_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, A6_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + A6_O_COL;

  if (warp == A6_CTRL_WARP) {
    // ------------------------------------------------------------ TMA producer + MMA issuer
    // every lane runs the warp-uniform control flow (descriptors in uniform registers); one elected
    // lane issues the TMA / MMA / commit instructions.
    if (n_tiles > 0) {
      const int qcol = head * A6_DH, kcol = D + head * A6_DH, vcol = 2 * D + head * A6_DH;
      auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col0, int row) {
        if (elect_one()) {
          mbar_arrive_expect_tx(bar, A6_TILE_BYTES);
          tma_load_2d(dst, &tmap_qkv, bar, col0, row);
        }
        __syncwarp();
      };
      load_tile(sQ, q_full, qcol, row_base + q0);
      load_tile(sK, &k_full[0], kcol, row_base);
      load_tile(sV, &v_full[0], vcol, row_base);
      if (n_tiles > 1) {
        load_tile(sK + A6_TILE_BYTES, &k_full[1], kcol, row_base + A6_BN);
        load_tile(sV + A6_TILE_BYTES, &v_full[1], vcol, row_base + A6_BN);
      }
      constexpr uint32_t idesc_s = make_idesc_bf16(A6_BM, A6_BN);
      constexpr uint32_t idesc_o = make_idesc_bf16(A6_BM, A6_DH) | (1u << 16);  // B (=V) MN-major
      const uint64_t q_desc = make_desc_k_sw128(smem_u32(sQ));
      const uint64_t k_desc0 = make_desc_k_sw128(smem_u32(sK));
      const uint64_t v_desc0 = desc_mn_sw128(smem_u32(sV), A6_BN * 128);

      auto issue_s = [&](int j) {
        const uint64_t kd = k_desc0 + (uint64_t)(((j & 1) * A6_TILE_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < A6_DH / 16; ++kk)
            tc_mma_ss(tS, q_desc + (uint64_t)((kk * 32) >> 4), kd + (uint64_t)((kk * 32) >> 4), idesc_s,
                      (uint32_t)(kk != 0));
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
          mbar_wait(s_free, (uint32_t)(j & 1));      // every softmax warp holds S(j) in registers
          A6_T(0, 1);   // s_free seen
          tc_fence_after();
          issue_s(j + 1);
          A6_T(0, 2);   // S(j+1) issued
#ifdef A6_TRACE
          if (traced) {                                // MMA round trip: issue -> completion visible
            mbar_wait(s_full, (uint32_t)((j + 1) & 1));
            A6_T(0, 5);
          }
#endif
          // S(j) has retired (the softmax threads read it), so K buffer j&1 can be refilled
          if (j + 2 < n_tiles)
            load_tile(sK + (j & 1) * A6_TILE_BYTES, &k_full[j & 1], kcol, row_base + (j + 2) * A6_BN);
        }
        mbar_wait(&v_full[j & 1], (uint32_t)((j >> 1) & 1));
        mbar_wait(p_full, (uint32_t)(j & 1));        // P(j) in smem, O rescaled if it had to be
        A6_T(0, 3);   // p_full seen
        tc_fence_after();
        {
          const uint64_t vd = v_desc0 + (uint64_t)(((j & 1) * A6_TILE_BYTES) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < A6_BN / 16; ++kk) {
              tc_mma_ts(tO, tmem_base + A6_P_COL + (uint32_t)(kk * 8), vd + (uint64_t)((kk * 16 * 128) >> 4),
                        idesc_o, (uint32_t)((j | kk) != 0));
            }
            tc_commit(pv_done);
            A6_T(0, 4);   // PV issued
          }
          __syncwarp();
        }
        if (j + 2 < n_tiles) {
          mbar_wait(pv_done, (uint32_t)(j & 1));     // V buffer j&1 is free once PV(j) retired
          load_tile(sV + (j & 1) * A6_TILE_BYTES, &v_full[j & 1], vcol, row_base + (j + 2) * A6_BN);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax: thread = row x 64 keys
    const int quarter = warp & 3;                    // TMEM lane quarter (hardware: warp id % 4)
    const int hf = warp >> 2;                        // key-column half of the tile
    const int r = quarter * 32 + lane;               // query row in the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS_mine = tS + lane_off + (uint32_t)(hf * 64);
    const uint32_t tO_mine = tO + lane_off + (uint32_t)(hf * 32);
    const uint32_t tP_mine = tmem_base + A6_P_COL + lane_off + (uint32_t)(hf * 32);
    float* x_own = xch + hf * 128 + r;               // + 256 * (tile parity)
    const float* x_peer = xch + (hf ^ 1) * 128 + r;
    float m_used = 0.f;                              // running max (log2 units) used in exponents
    float l_sum = 0.f;                               // this warp's half of the row sum

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, (uint32_t)(j & 1));
      A6_T(1, 1);   // s_full seen
      tc_fence_after();
      float s[64];
      tmem_ld_x64(tS_mine, s);
      tc_wait_ld();
      tc_fence_before();                               // S(j) is in registers: release the S columns
      __syncwarp();                                    // for QK^T of tile j+1 right away
      if (lane == 0) mbar_arrive(s_free);
      A6_T(1, 2);   // S in registers
      const int n_valid = klen - j * A6_BN - hf * 64;    // valid keys among this warp's 64 columns
      if (n_valid < 64) {                                // only the last tile has masked keys
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= n_valid) s[i] = -INFINITY;            // exp2(-inf) = 0 exactly
      }
      float m4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) m4[i] = max3(s[i], s[4 + i], s[8 + i]);
#pragma unroll
      for (int i = 12; i < 60; i += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) m4[k] = max3(m4[k], s[i + k], s[i + 4 + k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) m4[k] = fmaxf(m4[k], s[60 + k]);
      const float mx_own = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
      // the exchange slots alternate with the tile parity: a warp can only reach its write for tile
      // j+2 after the pair barrier of tile j+1, which its peer enters after reading the slot of tile j
      x_own[(j & 1) * 256] = mx_own;
      pair_barrier(1 + quarter);
      A6_T(1, 3);   // max + exchange barrier passed
      const float mx = fmaxf(mx_own, x_peer[(j & 1) * 256]);

      float factor = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = mx;
      } else {
        need = mx > m_used + A6_RESCALE_THRESHOLD;
        if (need) {
          factor = ex2a(m_used - mx);
          m_used = mx;
          l_sum *= factor;
        }
        // PV(j-1) must have retired before P is overwritten or O is rescaled
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {           // the peer warp takes the same branch
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[8];
            tmem_ld_x8(tO_mine + (uint32_t)(c * 8), o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_x8(tO_mine + (uint32_t)(c * 8), o);
          }
          tc_wait_st();
        }
      }

      // p = 2^(s*scale - m): one FFMA + one MUFU per element, 4 partial sums, bf16, 16-byte chunks
      // into the K-major SW128 layout (this warp's 64 keys = one 128-byte row of P half `hf`).
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_used;
      {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e0 = ex2a(fmaf(s[2 * i], scale_log2, neg_m));
          const float e1 = ex2a(fmaf(s[2 * i + 1], scale_log2, neg_m));
          sum4[i & 3] += e0 + e1;
          pk[i] = pack_bf16x2(e0, e1);
        }
        tmem_st_x32(tP_mine, pk);
      }
      l_sum += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      tc_wait_st();
      tc_fence_before();                             // P (and a rescaled O) ordered before the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      A6_T(1, 5);   // P stored, arrived
    }

    // ---- epilogue: O / (l_own + l_peer) -> bf16 -> ctx; this warp writes O columns hf*32..+31
    const int row = q0 + r;
    __nv_bfloat16* out = ctx + ((long long)(row_base + row)) * D + head * A6_DH + hf * 32;
    if (n_tiles > 0) {
      mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
      tc_fence_after();
      float* lx = reinterpret_cast<float*>(smem + A6_OFF_L);   // fp32 [2][128] row-sum exchange
      lx[hf * 128 + r] = l_sum;
      pair_barrier(1 + quarter);
      const float inv = 1.f / (l_sum + lx[(hf ^ 1) * 128 + r]);
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO_mine, o);
      tc_wait_ld();
      if (row < R) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(out + i * 8) = u;
        }
      }
    } else if (row < R) {                            // no valid key at all: zeros
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(0, 0, 0, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A6_CTRL_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A6_TMEM_COLS);
  }
#ifdef A6_TRACE
  if (blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 7 && threadIdx.x == 0) {
    printf("end %lld\n", clock64() - t_start);
    for (int j = 0; j < 8; ++j)
      printf("tile %d ctrl: s_free %lld S_issued %lld S_done %lld p_full %lld PV_issued %lld | smax: s_full %lld loaded %lld xchg %lld pvdone %lld arrived %lld\n",
             j, g_trace[0][j][1], g_trace[0][j][2], g_trace[0][j][5], g_trace[0][j][3], g_trace[0][j][4], g_trace[1][j][1],
             g_trace[1][j][2], g_trace[1][j][3], g_trace[1][j][4], g_trace[1][j][5]);
  }
#endif
}

}  // namespace

int attention_tc64_launch(const __nv_bfloat16* qkv, int B, int R, int heads, const int32_t* kv_len,
                          float scale, __nv_bfloat16* ctx, cudaStream_t s) {
  if (B <= 0 || R <= 0) return 0;
  const int D = heads * A6_DH;
  CUtensorMap tm;
  W2V_TRY(make_tmap_2d_bf16(&tm, qkv, (uint64_t)3 * D, (uint64_t)B * R, (uint64_t)3 * D, 64, A6_BN));
  const float scale_log2 = scale * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        A6_SMEM_BYTES));
    attr = true;
  }
  dim3 grid((R + A6_BM - 1) / A6_BM, heads, B);
  ProfScope ps(s, "attention_d64");
  attention_tc64_kernel<<<grid, A6_THREADS, A6_SMEM_BYTES, s>>>(tm, R, heads, kv_len, scale_log2, ctx);
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
