// tcgen05 / TMEM / TMA persistent GEMM with fused epilogues — see gemm_tc.cuh for the contract.
//
// CTA = 6 warps, one CTA per SM (persistent, static round-robin over output tiles):
//   warp 0      TMA producer     (one lane): A tile 128x64 + W tile BLOCK_Nx64 per stage
//   warp 1      MMA issuer       (one lane): 4 x tcgen05.mma (K=16) per stage, commit -> empty[]
//   warps 2..9  epilogue         (8 warps: 4 TMEM lane quarters x 2 column halves; thread = one
//                                 accumulator row): tcgen05.ld -> bias/act/residual -> global.
//                                 TMEM holds TWO accumulators so the epilogue of tile i overlaps
//                                 the MMAs of tile i+1.
#include "gemm_epilogue.cuh"

namespace w2v {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;              // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int GEMM_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;  // 256:4, 128:6, 64:8
  static constexpr int TMEM_COLS = 2 * BLOCK_N;               // 512 / 256 / 128 (powers of two)
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
};


template <int BLOCK_N, bool OUT_F32>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const KernelArgs p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = p.num_groups * p.tiles_m_per_group * n_tiles;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nb = tile % n_tiles;
        const int mt = tile / n_tiles;
        const int g = mt / p.tiles_m_per_group;
        const int i = mt - g * p.tiles_m_per_group;
        const long long arow0 = (long long)g * p.a_group_rows + (long long)i * BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int a_c0 = p.a_mode ? nb * BLOCK_N : kb * BLOCK_K;
          const int a_c1 = (int)(p.a_mode ? arow0 + kb : arow0);
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage], a_c0, a_c1);
          tma_load_2d(smem_b + stage * Cfg::B_STAGE_BYTES, &tmap_b, &full_bar[stage], kb * BLOCK_K,
                      nb * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // descriptors are warp-uniform values (uniform registers); one elected lane issues
        const uint64_t a_desc = make_desc_k_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES));
        const uint64_t b_desc = make_desc_k_sw128(smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements = 32 B along K inside the 128B swizzle row: +2 in (addr>>4)
            tc_mma_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                      (uint32_t)((kb | k) != 0));
          }
          tc_commit(&empty_bar[stage]);                       // smem slot reusable when MMAs retire
          if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // 8 warps: warp w reads TMEM lane quarter (w & 3) — the only one it may touch — and the
    // column half (w - 2) >> 2 of the tile. TMEM hands each thread one accumulator ROW; written
    // to global memory that way every warp store would touch 32 different cache lines. So each
    // 32-row x 128-byte block is transposed through a per-warp, XOR-swizzled 4 KB staging buffer:
    // bias / activation / row mask are applied row-per-thread, then the block is re-read with
    // lane -> (row = k*4 + lane/8, 16-byte chunk = lane%8) so that every global load (residual)
    // and store covers 4 full 128-byte lines. TMEM loads run one step ahead of the math.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int COLS = BLOCK_N / 2;                 // columns per warp
    const uint32_t stage = smem_u32(smem + STAGES * Cfg::STAGE_BYTES + 256 + (warp - 2) * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nb = tile % n_tiles;
      const int mt = tile / n_tiles;
      const int g = mt / p.tiles_m_per_group;
      const int i = mt - g * p.tiles_m_per_group;
      const int rg0 = i * BLOCK_M + q * 32;           // first row (in group) of this warp's block
      const long long orow0 = (long long)g * p.o_group_rows + rg0;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) +
                              (uint32_t)(acc * BLOCK_N + half * COLS);
      gemm_epilogue_warp<COLS, OUT_F32>(p, rg0, orow0, nb * BLOCK_N + half * COLS, t_base, stage, lane,
                                        [&] {
                                          mbar_wait(&tfull_bar[acc], acc_phase);
                                          tc_fence_after();
                                        });
      // all TMEM reads of this accumulator are complete (wait::ld above): hand it back
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
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N>
int launch_impl(const GemmProblem& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap tm_a, tm_b;
  W2V_TRY(make_tmap_2d_bf16(&tm_a, g.A, (uint64_t)g.a_cols, (uint64_t)g.a_rows,
                            (uint64_t)g.a_row_stride, BLOCK_K, BLOCK_M));
  W2V_TRY(make_tmap_2d_bf16(&tm_b, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, BLOCK_K,
                            BLOCK_N));
  KernelArgs a;
  a.tma_store = 0;
  a.N = g.N; a.K = g.K;
  a.num_groups = g.num_groups;
  a.rows_per_group = g.rows_per_group;
  a.tiles_m_per_group = (g.rows_per_group + BLOCK_M - 1) / BLOCK_M;
  a.a_group_rows = g.a_group_rows;
  a.o_group_rows = g.o_group_rows;
  a.a_mode = g.a_mode;
  a.bias = g.bias;
  a.act_split = g.act_split; a.act_lo = g.act_lo; a.act_hi = g.act_hi;
  a.resid = g.resid; a.ld_resid = g.ld_resid;
  a.out = g.out; a.ld_out = g.ld_out; a.out_f32 = g.out_f32;
  a.mask_len = g.mask_len; a.mask_period = g.mask_period > 0 ? g.mask_period : 1;

  W2V_ONCE_BEGIN
    W2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    if constexpr (BLOCK_N >= 128) {
      W2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES));
    }
  W2V_ONCE_END
  const long long num_tiles =
      (long long)a.num_groups * a.tiles_m_per_group * (g.N / BLOCK_N);
  if (num_tiles == 0) return 0;
  const int grid = (int)(num_tiles < (long long)num_sms() ? num_tiles : (long long)num_sms());
  {
    ProfScope ps(stream, BLOCK_N == 256 ? "gemm_bn256" : (BLOCK_N == 128 ? "gemm_bn128" : "gemm_bn64"));
    if (g.out_f32) {
      gemm_tc_kernel<BLOCK_N, true><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_a, tm_b, a);
    } else {
      if constexpr (BLOCK_N >= 128)
        gemm_tc_kernel<BLOCK_N, false><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_a, tm_b, a);
    }
  }
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int gemm_tc_launch(const GemmProblem& g, int block_n, cudaStream_t stream) {
  W2V_REQUIRE(g.K > 0 && g.K % BLOCK_K == 0, "gemm: K=%d must be a positive multiple of 64", g.K);
  W2V_REQUIRE(g.N > 0 && g.N % block_n == 0, "gemm: N=%d must be a multiple of block_n=%d", g.N,
              block_n);
  W2V_REQUIRE(g.act_split % 32 == 0, "gemm: act_split=%d must be a multiple of 32", g.act_split);
  W2V_REQUIRE(g.ld_out % 8 == 0 && (g.resid == nullptr || g.ld_resid % 4 == 0),
              "gemm: output/residual leading dimensions must keep 16-byte alignment");
  W2V_REQUIRE(g.resid == nullptr || g.out_f32, "gemm: residual requires fp32 output");
  W2V_REQUIRE(g.a_row_stride % 8 == 0, "gemm: A row stride must be a multiple of 8 elements");
  W2V_REQUIRE(g.out_f32 || block_n >= 128, "gemm: bf16 output needs block_n >= 128");
  switch (block_n) {
    case 256: return launch_impl<256>(g, stream);
    case 128: return launch_impl<128>(g, stream);
    case 64:  return launch_impl<64>(g, stream);
    default:
      set_error("gemm: unsupported block_n=%d", block_n);
      return W2VSEG_ERR_ARG;
  }
}

}  // namespace w2v
