// CTA-pair variant of the tcgen05 GEMM (gemm_tc.cuh contract, a_mode 0, N % 256 == 0).
//
// ncu on the single-CTA kernel showed the tensor pipe only 50-60 % active: a 128x256 tile pulls
// (128+256)*K*2 bytes through L2 for 2*128*256*K flops, which is more than the L2->SM fabric
// sustains with all 148 SMs busy. Here two CTAs of a cluster (one TPC) cooperate on a 256x256 tile
// with tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and ONE HALF of the W
// tile (128 of the 256 columns); the tensor cores of both SMs read both halves. Bytes per flop
// through L2 drop by a third and each SM's shared memory sees 32 KB instead of 48 KB per k-block.
//
// Roles per CTA (10 warps): warp 0 TMA producer (own A rows + own W half, completion credited to
// the leader's mbarrier), warp 1 MMA issuer (leader CTA only; commits are multicast to both CTAs),
// warps 2..9 epilogue (each CTA drains its own 128 accumulator rows; TMEM double-buffered). The epilogue
// leaves through the TMA engine (bf16: cp.async.bulk.tensor stores of swizzled 32 x 32 boxes; in-place
// fp32 residual: cp.reduce.async.bulk.tensor .add of 32 x 32 fp32 boxes) and hands the accumulator back
// with a CTA-scope-release remote arrive — see gemm_epilogue.cuh / ptx.cuh for why (a cluster-scope
// release there cost ~3000 cycles per tile; profiles/gemm_epilogue_trace_r01_s3.md).
#include <stdlib.h>

#include "gemm_epilogue.cuh"

namespace w2v {

namespace {

constexpr int BLOCK_M = 128;     // rows per CTA (256 per pair)
constexpr int PAIR_M = 256;
constexpr int BLOCK_N = 256;     // columns per pair tile (128 W rows staged per CTA)
constexpr int HALF_N = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_BYTES = HALF_N * BLOCK_K * 2;    // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
// Operand ring depth and epilogue staging per warp: 6 stages + 4 KB (bf16 output: two 2 KB halves for the TMA
// stores; fp32 in-place residual: one 4 KB block per TMA reduce-add). -DW2V_PAIR_F32_STAGES=5 gives the fp32
// kernel TWO 4 KB staging blocks (the reduce-add of block u reads its block while block u+1 is staged) at the
// price of one ring stage: measured SLOWER (attn_out in place 28.2 -> 30.0 us, ffn_down 95.0 -> 95.4 us,
// profiles/experiments_r02.md), so the default stays 6.
#ifndef W2V_PAIR_F32_STAGES
#define W2V_PAIR_F32_STAGES 6
#endif
template <bool OUT_F32>
struct PairCfg {
  static constexpr int STAGES = OUT_F32 ? W2V_PAIR_F32_STAGES : 6;
  static constexpr int STAGE_BUF = STAGES == 6 ? 4096 : 8192;
  static constexpr int TMA_BUFS = 2;                       // bf16 path: 2 KB halves of the first 4 KB
  static constexpr int F32_BUFS = STAGE_BUF / 4096;        // fp32 path: 4 KB blocks
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 8 * STAGE_BUF;
};
constexpr int TMEM_COLS = 512;                   // two 256-column fp32 accumulators
constexpr int THREADS = 320;

template <bool OUT_F32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_o, const KernelArgs p) {
  constexpr int STAGES = PairCfg<OUT_F32>::STAGES;
  constexpr int STAGE_BUF = PairCfg<OUT_F32>::STAGE_BUF;
  constexpr int TMA_BUFS = PairCfg<OUT_F32>::TMA_BUFS;
  constexpr int F32_BUFS = PairCfg<OUT_F32>::F32_BUFS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  // epilogue staging (8 x 4 KB, 1024-byte aligned: a staging block doubles as a swizzled TMA box)
  uint8_t* smem_stage = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + 8 * STAGE_BUF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] (only the leader's copy is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = p.num_groups * p.tiles_m_per_group * n_tiles;  // tiles_m = 256-row blocks
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_store) tma_prefetch_desc(&tmap_o);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);  // 8 epilogue warps in each CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // the whole warp runs the loop (uniform registers for coordinates), one elected lane issues
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int nb = tile % n_tiles;
        const int mt = tile / n_tiles;
        const int g = mt / p.tiles_m_per_group;
        const int i = mt - g * p.tiles_m_per_group;
        const int arow0 = (int)((long long)g * p.a_group_rows + (long long)i * PAIR_M + rank * BLOCK_M);
        const int brow0 = nb * BLOCK_N + (int)rank * HALF_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            tma_load_2d_2sm(smem_a + stage * A_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, arow0);
            tma_load_2d_2sm(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * BLOCK_K, brow0);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        W2V_TR(3, it);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        W2V_TR(4, it);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // descriptors are warp-uniform values (uniform registers); one elected lane issues
          const uint64_t a_desc = make_desc_k_sw128(smem_u32(smem_a + stage * A_BYTES));
          const uint64_t b_desc = make_desc_k_sw128(smem_u32(smem_b + stage * B_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              tc_mma_ss_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                            (uint32_t)((kb | k) != 0));
            tc_commit_2sm_mc(&empty_bar[stage], 3);                      // both producers
            if (kb == num_kb - 1) tc_commit_2sm_mc(&tfull_bar[acc], 3);  // both epilogues
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        W2V_TR(5, it);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int COLS = BLOCK_N / 2;
    const uint32_t stage_buf = smem_u32(smem_stage + (warp - 2) * STAGE_BUF);
    const CUtensorMap* tm_o = p.tma_store ? &tmap_o : nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int nb = tile % n_tiles;
      const int mt = tile / n_tiles;
      const int g = mt / p.tiles_m_per_group;
      const int i = mt - g * p.tiles_m_per_group;
      const int rg0 = i * PAIR_M + (int)rank * BLOCK_M + q * 32;
      const long long orow0 = (long long)g * p.o_group_rows + rg0;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) +
                              (uint32_t)(acc * BLOCK_N + half * COLS);
      gemm_epilogue_warp<COLS, OUT_F32, TMA_BUFS, F32_BUFS>(p, rg0, orow0, nb * BLOCK_N + half * COLS, t_base, stage_buf,
                                        lane, [&] {
                                          if (warp == 2) W2V_TR(0, it);
                                          mbar_wait(&tfull_bar[acc], acc_phase);
                                          if (warp == 2) W2V_TR(1, it);
                                          tc_fence_after();
                                        }, warp == 2 ? it : -1, tm_o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);   // the leader's MMA thread waits here
      if (warp == 2) W2V_TR(2, it);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (tm_o != nullptr && lane == 0) bulk_wait_group_read0();   // staging must outlive the last store's read
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still signal or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int gemm_tc2_launch(const GemmProblem& g, cudaStream_t stream) {
  W2V_REQUIRE(g.a_mode == 0 && g.N % BLOCK_N == 0 && g.K % BLOCK_K == 0 && g.K > 0,
              "gemm2: needs a_mode 0, N %% 256 == 0, K %% 64 == 0 (N=%d, K=%d)", g.N, g.K);
  W2V_REQUIRE(g.act_split % 64 == 0, "gemm2: act_split=%d must be a multiple of 64", g.act_split);
  W2V_REQUIRE(g.ld_out % 8 == 0 && (g.resid == nullptr || g.ld_resid % 4 == 0),
              "gemm2: output/residual leading dimensions must keep 16-byte alignment");
  W2V_REQUIRE(g.resid == nullptr || g.out_f32, "gemm2: residual requires fp32 output");
  W2V_REQUIRE(g.a_row_stride % 8 == 0, "gemm2: A row stride must be a multiple of 8 elements");
  CUtensorMap tm_a, tm_b;
  W2V_TRY(make_tmap_2d_bf16(&tm_a, g.A, (uint64_t)g.a_cols, (uint64_t)g.a_rows,
                            (uint64_t)g.a_row_stride, BLOCK_K, BLOCK_M));
  W2V_TRY(make_tmap_2d_bf16(&tm_b, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, BLOCK_K, HALF_N));
  // bf16 output of a flat problem leaves through TMA stores (32-row x 32-column boxes, 64-byte swizzle)
  static const bool tma_allowed = !(getenv("W2VSEG_GEMM_TMA_STORE") != nullptr &&
                                    atoi(getenv("W2VSEG_GEMM_TMA_STORE")) == 0);   // =0: A/B measurements
  const bool tma_out = tma_allowed && !g.out_f32 && g.num_groups == 1 &&
                       (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && g.ld_out % 8 == 0;
  CUtensorMap tm_o = tm_a;
  if (tma_out)
    W2V_TRY(make_tmap_2d_bf16_sw64(&tm_o, g.out, (uint64_t)g.N, (uint64_t)g.rows_per_group, (uint64_t)g.ld_out, 32, 32));
  // ... and the in-place fp32 residual (h += acc + bias) through TMA reduce-adds (32 x 32 fp32 boxes)
  const bool tma_red = tma_allowed && g.out_f32 && g.num_groups == 1 && g.resid != nullptr &&
                       g.resid == g.out && g.ld_resid == g.ld_out && g.mask_len == nullptr &&
                       (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && g.ld_out % 4 == 0;
  if (tma_red)
    W2V_TRY(make_tmap_2d_f32(&tm_o, g.out, (uint64_t)g.N, (uint64_t)g.rows_per_group, (uint64_t)g.ld_out, 32, 32));
  KernelArgs a;
  a.tma_store = (tma_out || tma_red) ? 1 : 0;
  a.N = g.N; a.K = g.K;
  a.num_groups = g.num_groups;
  a.rows_per_group = g.rows_per_group;
  a.tiles_m_per_group = (g.rows_per_group + PAIR_M - 1) / PAIR_M;
  a.a_group_rows = g.a_group_rows;
  a.o_group_rows = g.o_group_rows;
  a.a_mode = 0;
  a.bias = g.bias;
  a.act_split = g.act_split; a.act_lo = g.act_lo; a.act_hi = g.act_hi;
  a.resid = g.resid; a.ld_resid = g.ld_resid;
  a.out = g.out; a.ld_out = g.ld_out; a.out_f32 = g.out_f32;
  a.mask_len = g.mask_len; a.mask_period = g.mask_period > 0 ? g.mask_period : 1;

  W2V_ONCE_BEGIN
  W2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<true>::SMEM_BYTES));
  W2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<false>::SMEM_BYTES));
  W2V_ONCE_END
  const long long num_tiles = (long long)a.num_groups * a.tiles_m_per_group * (g.N / BLOCK_N);
  if (num_tiles == 0) return 0;
  const long long max_clusters = num_sms() / 2;
  const int grid = 2 * (int)(num_tiles < max_clusters ? num_tiles : max_clusters);
  {
    ProfScope ps(stream, "gemm_pair256");
    if (g.out_f32)
      gemm_tc2_kernel<true><<<grid, THREADS, PairCfg<true>::SMEM_BYTES, stream>>>(tm_a, tm_b, tm_o, a);
    else
      gemm_tc2_kernel<false><<<grid, THREADS, PairCfg<false>::SMEM_BYTES, stream>>>(tm_a, tm_b, tm_o, a);
  }
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v

#ifdef W2VSEG_TRACE
extern "C" int32_t w2vseg_debug_trace(long long* host_dst, int32_t n) {
  return (int32_t)cudaMemcpyFromSymbol(host_dst, g_trace, sizeof(long long) * (size_t)(n < 768 ? n : 768));
}
#endif
