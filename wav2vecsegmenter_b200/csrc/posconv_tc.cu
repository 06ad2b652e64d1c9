// 128-tap, 16-group positional convolution (HF:326-379: Conv1d(1024, 1024, k=128, padding=64,
// groups=16) + SamePad + GELU, and the residual add of HF:764-765) as a tcgen05 implicit GEMM whose A
// operand stays RESIDENT in shared memory.
//
// For one output tile (128 frames x the 64 channels of one group) tap j multiplies the 64 input
// channels of rows r0+j .. r0+j+127 of the zero-padded activations: the 128 taps read the same 255
// rows, each shifted by one row. The generic kernel (gemm_tc.cu, a_mode 1) re-fetched that 16 KB box
// for every tap — 24 KB of L2 traffic per 4 MMAs, 3 MB per tile, 5.4 GB per launch — and ran at
// 534 TFLOP/s, L2-fabric-bound. Here the 256-row block is loaded ONCE per tile (2 TMA boxes, 32 KB,
// double-buffered across tiles) and tap j addresses it through a shared-memory descriptor whose start
// is advanced by j rows (128 B each). The 128-byte swizzle of both TMA and tcgen05.mma is a function
// of the shared-memory ADDRESS bits (the XOR of bits 4-6 with bits 7-9), so a start that is not a
// multiple of 1024 B needs no correction: the descriptor's base-offset field stays 0 (measured —
// with (start >> 7) & 7 in that field the results are wrong). Only the tap weights (8 KB per tap)
// stream through the mbarrier ring: 1 MB per tile.
//
//   warp 0      TMA producer: A block per tile, W tile per tap
//   warp 1      MMA issuer (warp-uniform control flow, elect_one): 4 x tcgen05.mma M128 N64 K16 per tap
//   warps 2..9  epilogue (gemm_epilogue.cuh): bias + GELU, h += .. (red.global.add.v4.f32)
#include "gemm_epilogue.cuh"

namespace w2v {

namespace {

constexpr int PC_BM = 128;
constexpr int PC_BN = 64;                 // channels per group
constexpr int PC_BK = 64;                 // input channels per group = one 128-byte swizzle row
constexpr int PC_THREADS = 320;
constexpr int PC_A_BYTES = 2 * PC_BM * PC_BK * 2;   // 256 rows x 128 B = 32 KB per buffer
// Taps per pipeline stage: one mbarrier round trip per 32 MMAs. The N=64 MMAs are small (32 tensor
// cycles each), so with one tap per stage the issuing warp, not the tensor pipe, set the pace:
// 1 / 2 / 4 / 8 taps per stage -> ~470 (in step) / 308 / 247 / 222 us at B=14 (generic kernel: 380 us).
constexpr int PC_TPS = 8;
constexpr int PC_TAP_BYTES = PC_BN * PC_BK * 2;     // 8 KB of weights per tap
constexpr int PC_B_BYTES = PC_TPS * PC_TAP_BYTES;   // 64 KB per stage
constexpr int PC_STAGES = 2;
constexpr int PC_OFF_B = 2 * PC_A_BYTES;
constexpr int PC_OFF_BAR = PC_OFF_B + PC_STAGES * PC_B_BYTES;
constexpr int PC_OFF_STAGE = PC_OFF_BAR + 1024;
constexpr int PC_SMEM_BYTES = PC_OFF_STAGE + 8 * 4096 + 1024 /*align*/;
constexpr int PC_TMEM_COLS = 2 * PC_BN;

__global__ void __launch_bounds__(PC_THREADS, 1)
posconv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const KernelArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;                       // [2][32 KB]
  uint8_t* smem_b = smem + PC_OFF_B;            // [PC_STAGES][PC_TPS x 8 KB]
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + PC_OFF_BAR);
  uint64_t* b_empty = b_full + PC_STAGES;
  uint64_t* a_full = b_empty + PC_STAGES;       // [2]
  uint64_t* a_empty = a_full + 2;               // [2]
  uint64_t* tfull_bar = a_empty + 2;            // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / PC_BN;              // groups
  const int num_tiles = p.num_groups * p.tiles_m_per_group * n_tiles;
  const int taps = p.K / PC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < PC_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, PC_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int nb = tile % n_tiles;
      const int mt = tile / n_tiles;
      const int g = mt / p.tiles_m_per_group;
      const int i = mt - g * p.tiles_m_per_group;
      const int arow0 = (int)((long long)g * p.a_group_rows + (long long)i * PC_BM);
      const int buf = it & 1;
      mbar_wait(&a_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
      if (elect_one()) {
        mbar_arrive_expect_tx(&a_full[buf], PC_A_BYTES);
        tma_load_2d(smem_a + buf * PC_A_BYTES, &tmap_a, &a_full[buf], nb * PC_BN, arow0);
        tma_load_2d(smem_a + buf * PC_A_BYTES + PC_A_BYTES / 2, &tmap_a, &a_full[buf], nb * PC_BN, arow0 + PC_BM);
      }
      __syncwarp();
      for (int kb = 0; kb < taps; kb += PC_TPS) {
        const int nt = min(PC_TPS, taps - kb);
        mbar_wait(&b_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[stage], (uint32_t)(nt * PC_TAP_BYTES));
          for (int t = 0; t < nt; ++t)
            tma_load_2d(smem_b + stage * PC_B_BYTES + t * PC_TAP_BYTES, &tmap_b, &b_full[stage],
                        (kb + t) * PC_BK, nb * PC_BN);
        }
        __syncwarp();
        if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(PC_BM, PC_BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      mbar_wait(&a_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * PC_BN);
      const uint32_t a_base = smem_u32(smem_a + buf * PC_A_BYTES);
      for (int kb = 0; kb < taps; kb += PC_TPS) {
        const int nt = min(PC_TPS, taps - kb);
        mbar_wait(&b_full[stage], phase);
        tc_fence_after();
        // tap kb+t: the same block, start advanced by kb+t rows (base-offset field stays 0, see header)
        const uint64_t a_desc0 = make_desc_k_sw128(a_base + (uint32_t)(kb * 128));
        const uint64_t b_desc0 = make_desc_k_sw128(smem_u32(smem_b + stage * PC_B_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < PC_TPS; ++t) {
            if (t >= nt) break;
            const uint64_t a_desc = a_desc0 + (uint64_t)(t * (128 >> 4));
            const uint64_t b_desc = b_desc0 + (uint64_t)(t * (PC_TAP_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < PC_BK / 16; ++k)
              tc_mma_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                        (uint32_t)((kb | t | k) != 0));
          }
          tc_commit(&b_empty[stage]);
          if (kb + nt == taps) {
            tc_commit(&tfull_bar[acc]);   // accumulator complete
            tc_commit(&a_empty[buf]);     // every tap has read the resident block
          }
        }
        __syncwarp();
        if (++stage == PC_STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int COLS = PC_BN / 2;
    const uint32_t stage_buf = smem_u32(smem + PC_OFF_STAGE + (warp - 2) * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nb = tile % n_tiles;
      const int mt = tile / n_tiles;
      const int g = mt / p.tiles_m_per_group;
      const int i = mt - g * p.tiles_m_per_group;
      const int rg0 = i * PC_BM + q * 32;
      const long long orow0 = (long long)g * p.o_group_rows + rg0;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * PC_BN + half * COLS);
      gemm_epilogue_warp<COLS, true>(p, rg0, orow0, nb * PC_BN + half * COLS, t_base, stage_buf, lane, [&] {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, PC_TMEM_COLS);
  }
}

}  // namespace

// Same GemmProblem contract as gemm_tc_launch(p, 64, ..) with a_mode 1 (K = taps * 64, N = groups * 64,
// fp32 output), see engine.cu.
int posconv_tc_launch(const GemmProblem& g, cudaStream_t stream) {
  W2V_REQUIRE(g.a_mode == 1 && g.out_f32 == 1, "posconv: needs the shifted-row A view and fp32 output");
  W2V_REQUIRE(g.N % PC_BN == 0 && g.K % PC_BK == 0 && g.K / PC_BK <= 128,
              "posconv: N=%d must be groups x 64, K=%d must be taps x 64 with at most 128 taps", g.N, g.K);
  W2V_REQUIRE(g.a_cols == g.N, "posconv: A view must be [rows, groups x 64]");
  W2V_REQUIRE(g.ld_out % 8 == 0 && (g.resid == nullptr || g.ld_resid % 4 == 0),
              "posconv: output/residual leading dimensions must keep 16-byte alignment");
  CUtensorMap tm_a, tm_b;
  W2V_TRY(make_tmap_2d_bf16(&tm_a, g.A, (uint64_t)g.a_cols, (uint64_t)g.a_rows, (uint64_t)g.a_row_stride,
                            PC_BK, PC_BM));
  W2V_TRY(make_tmap_2d_bf16(&tm_b, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, PC_BK, PC_BN));
  KernelArgs a;
  a.tma_store = 0;
  a.N = g.N; a.K = g.K;
  a.num_groups = g.num_groups;
  a.rows_per_group = g.rows_per_group;
  a.tiles_m_per_group = (g.rows_per_group + PC_BM - 1) / PC_BM;
  a.a_group_rows = g.a_group_rows;
  a.o_group_rows = g.o_group_rows;
  a.a_mode = 1;
  a.bias = g.bias;
  a.act_split = g.act_split; a.act_lo = g.act_lo; a.act_hi = g.act_hi;
  a.resid = g.resid; a.ld_resid = g.ld_resid;
  a.out = g.out; a.ld_out = g.ld_out; a.out_f32 = 1;
  a.mask_len = g.mask_len; a.mask_period = g.mask_period > 0 ? g.mask_period : 1;
  W2V_ONCE_BEGIN
    W2V_CHECK_CUDA(cudaFuncSetAttribute(posconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        PC_SMEM_BYTES));
  W2V_ONCE_END
  const long long num_tiles = (long long)a.num_groups * a.tiles_m_per_group * (g.N / PC_BN);
  if (num_tiles == 0) return 0;
  const int grid = (int)(num_tiles < (long long)num_sms() ? num_tiles : (long long)num_sms());
  {
    ProfScope ps(stream, "posconv");
    posconv_tc_kernel<<<grid, PC_THREADS, PC_SMEM_BYTES, stream>>>(tm_a, tm_b, a);
  }
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
