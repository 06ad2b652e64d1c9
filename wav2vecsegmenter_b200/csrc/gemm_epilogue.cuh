// Epilogue shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels (see gemm_tc.cu for the scheme).
#pragma once
#include "gemm_tc.cuh"
#include "ptx.cuh"

#ifdef W2VSEG_TRACE
// experiments only (scripts/prof_gemm_trace.py): clock64 stamps of CTA 0, one copy per translation unit
static __device__ long long g_trace[12 * 64];
#define W2V_TR(slot, idx)                                                                       \
  do {                                                                                          \
    if (blockIdx.x == 0 && lane == 0 && (idx) >= 0 && (idx) < 64) g_trace[(slot) * 64 + (idx)] = clock64(); \
  } while (0)
#else
#define W2V_TR(slot, idx) do { } while (0)
#endif

namespace w2v {

struct KernelArgs {
  int N, K;
  int num_groups, rows_per_group, tiles_m_per_group;
  long long a_group_rows, o_group_rows;
  int a_mode;
  const float* bias;
  int act_split, act_lo, act_hi;
  const float* resid;
  long long ld_resid;
  void* out;
  long long ld_out;
  int out_f32;
  const int* mask_len;
  int mask_period;
  int tma_store;   // bf16 output leaves through a TMA store of the staging block (tmap_o valid)
};

// One epilogue warp, one accumulator tile: rows rg0..rg0+31 (TMEM lane quarter of this warp),
// COLS columns starting at colbase. `wait_full` is invoked after the residual prefetch has been
// issued and must block until the accumulator is complete (mbarrier wait + tcgen05 fence).
template <int COLS, bool OUT_F32, int TMA_BUFS = 2, int F32_BUFS = 1, typename WaitFull>
__device__ __forceinline__ void gemm_epilogue_warp(const KernelArgs& p, int rg0, long long orow0,
                                                   int colbase, uint32_t t_base, uint32_t stage,
                                                   int lane, WaitFull&& wait_full, int trace_it = -1,
                                                   const CUtensorMap* tmap_o = nullptr) {
  constexpr int UNIT = OUT_F32 ? 32 : 64;           // columns per staging block (128 B per row)
  static_assert(COLS % UNIT == 0, "bf16 output needs BLOCK_N >= 128");
  constexpr int NUNIT = COLS / UNIT;
  constexpr int LPU = UNIT / 32;                    // tcgen05.ld x32 per unit
  const int c_row = lane >> 3;                      // coalesced phase: row within a group of 4
  const int c_chk = lane & 7;                       //                  16-byte chunk of the row
  bool zero_row = false;                            // row-per-thread view: row rg0 + lane
  if (p.mask_len != nullptr && rg0 + lane < p.rows_per_group) {
    const long long orow = orow0 + lane;
    const long long w = orow / p.mask_period;
    zero_row = (int)(orow - w * p.mask_period) >= __ldg(p.mask_len + w);
  }
  // in-place residual (out == resid, every residual GEMM of the engine): h += acc + bias as a
  // fire-and-forget vector reduction at L2 — the epilogue never waits for a residual load
  const bool use_red = OUT_F32 && p.resid != nullptr && p.resid == p.out && p.ld_resid == p.ld_out;
  const bool use_resid = OUT_F32 && p.resid != nullptr && !use_red;

  float4 res[2][8];
  if constexpr (OUT_F32) {
    if (use_resid) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = k * 4 + c_row;
        if (rg0 + rr < p.rows_per_group)
          res[0][k] = *reinterpret_cast<const float4*>(p.resid + (orow0 + rr) * p.ld_resid +
                                                       colbase + c_chk * 4);
      }
    }
  }

  wait_full();
  uint32_t raw[UNIT];
#pragma unroll
  for (int l = 0; l < LPU; ++l)
    tmem_ld_32x32b_x32(t_base + (uint32_t)(l * 32), *reinterpret_cast<uint32_t(*)[32]>(&raw[l * 32]));

#pragma unroll
  for (int u = 0; u < NUNIT; ++u) {
    tc_wait_ld();
    if (u == 0) W2V_TR(6, trace_it);
    if (u == 1) W2V_TR(8, trace_it);
    if (u + 1 < NUNIT) {
      if constexpr (OUT_F32) {
        if (use_resid) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = k * 4 + c_row;
            if (rg0 + rr < p.rows_per_group)
              res[(u + 1) & 1][k] = *reinterpret_cast<const float4*>(
                  p.resid + (orow0 + rr) * p.ld_resid + colbase + (u + 1) * UNIT + c_chk * 4);
          }
        }
      }
    }
    const int col0 = colbase + u * UNIT;
    const int act = (col0 < p.act_split) ? p.act_lo : p.act_hi;
    float v[UNIT];
#pragma unroll
    for (int j = 0; j < UNIT; ++j) v[j] = __uint_as_float(raw[j]);
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int j = 0; j < UNIT / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < UNIT; ++j) v[j] = gelu_erf(v[j]);
    } else if (act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < UNIT; ++j) v[j] = fmaxf(v[j], 0.f);
    }
#ifdef W2VSEG_EPI_EXPERIMENT
    else if (act == 3) {   // MUFU only
#pragma unroll
      for (int j = 0; j < UNIT; ++j) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v[j])); v[j] = t; }
    } else if (act == 4) {   // 8 FP ops, no MUFU
#pragma unroll
      for (int j = 0; j < UNIT; ++j) {
        const float x = v[j];
        const float x2 = fminf(x * x, 36.f);
        float p = fmaf(-0.00035854941503117723f, x2, 0.03704889695510829f);
        p = fmaf(p, x2, 0.7974606740658886f);
        const float t = p * x;
        const float hx = 0.5f * x;
        v[j] = fmaf(hx, t, hx);
      }
    } else if (act == 5) {   // 3 FP ops + MUFU
#pragma unroll
      for (int j = 0; j < UNIT; ++j) {
        const float x = v[j];
        float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * 0.79f));
        const float hx = 0.5f * x;
        v[j] = fmaf(hx, t, hx);
      }
    }
#endif
    if (zero_row) {
#pragma unroll
      for (int j = 0; j < UNIT; ++j) v[j] = 0.f;
    }
#ifdef W2VSEG_TRACE
    {   // keep the math above the stamp: consume one value through a volatile asm
      asm volatile("" ::"f"(v[UNIT - 1]), "f"(v[0]));
      if (u == 0) W2V_TR(9, trace_it);
      if (u == 1) W2V_TR(10, trace_it);
    }
#endif
    // bf16 output through TMA (tmap_o): the 64-column block leaves as TWO 32-row x 32-column boxes
    // (64-byte rows, 64-byte swizzle), each from its own 2 KB half of the staging buffer, so the warp
    // never waits for the store it has just issued: before a half is overwritten only the store issued
    // from it one block earlier must have been read (wait_group.read 1). The coalescing loop below
    // (8 x LDS.128 + STG.128 per block) held the warp ~2400 cycles per block under load — the SM's store
    // path drains at ~10 B/clk while the operand loads run — which made the epilogue, not the MMAs, set
    // the pace of the GEMMs with an activation.
    bool via_tma = false;
    if constexpr (!OUT_F32) via_tma = tmap_o != nullptr;
    if constexpr (!OUT_F32) {
      if (via_tma) {
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          // TMA_BUFS 2 KB buffers per warp, used round-robin (NUNIT * 2 blocks per tile is a multiple of
          // TMA_BUFS): before one is overwritten, all but the TMA_BUFS-1 newest stores must have been read
          const uint32_t buf = stage + (uint32_t)(((u * 2 + hb) % TMA_BUFS) * 2048);
          if (lane == 0) {
            if constexpr (TMA_BUFS == 4) bulk_wait_group_read3(); else bulk_wait_group_read1();
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = hb * 32 + c * 8;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                             buf + (uint32_t)(lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4))),
                         "r"(pack_bf16x2(v[j + 0], v[j + 1])), "r"(pack_bf16x2(v[j + 2], v[j + 3])),
                         "r"(pack_bf16x2(v[j + 4], v[j + 5])), "r"(pack_bf16x2(v[j + 6], v[j + 7]))
                         : "memory");
          }
          fence_proxy_async_smem();   // staging writes -> visible to the async proxy
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tmap_o, buf, col0 + hb * 32, (int)orow0);   // rows past the end are clipped
            bulk_commit_group();
          }
        }
        if (u + 1 < NUNIT) {   // v[] is dead: fetch the next block (TMEM load latency is ~60 cycles)
#pragma unroll
          for (int l = 0; l < LPU; ++l)
            tmem_ld_32x32b_x32(t_base + (uint32_t)((u + 1) * UNIT + l * 32),
                               *reinterpret_cast<uint32_t(*)[32]>(&raw[l * 32]));
        }
        if (u == 0) W2V_TR(7, trace_it);
        continue;
      }
    }
    // in-place fp32 residual through TMA (tmap_o): the 32 x 32 fp32 staging block (128-byte rows, the
    // XOR pattern below IS the 128-byte TMA swizzle) goes out as ONE bulk reduce-add (h += block at L2)
    // instead of 8 x (LDS.128 + RED.v4) per thread
    bool red_tma = false;
    if constexpr (OUT_F32) red_tma = use_red && tmap_o != nullptr;
    // F32_BUFS 4 KB staging blocks used round-robin (NUNIT per tile is a multiple of F32_BUFS): before one is
    // overwritten, all but the F32_BUFS-1 newest reductions must have finished READING their blocks
    const uint32_t sblk = stage + (uint32_t)((F32_BUFS > 1 && red_tma ? (u % F32_BUFS) : 0) * 4096);
    if (red_tma) {
      if (lane == 0) {
        if constexpr (F32_BUFS == 2) bulk_wait_group_read1(); else bulk_wait_group_read0();
      }
      __syncwarp();
    }
    // row-per-thread -> staging (row = lane, 8 chunks of 16 B, chunk index XOR row%8)
    const uint32_t srow = sblk + (uint32_t)(lane * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t w0, w1, w2, w3;
      if constexpr (OUT_F32) {
        w0 = __float_as_uint(v[4 * j + 0]); w1 = __float_as_uint(v[4 * j + 1]);
        w2 = __float_as_uint(v[4 * j + 2]); w3 = __float_as_uint(v[4 * j + 3]);
      } else {
        w0 = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); w1 = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        w2 = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); w3 = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)((j ^ (lane & 7)) << 4)),
                   "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                   : "memory");
    }
    if (u + 1 < NUNIT) {   // accumulator registers are free again: fetch the next block now,
#pragma unroll               // its latency hides behind the store phase below
      for (int l = 0; l < LPU; ++l)
        tmem_ld_32x32b_x32(t_base + (uint32_t)((u + 1) * UNIT + l * 32),
                           *reinterpret_cast<uint32_t(*)[32]>(&raw[l * 32]));
    }
    if (red_tma) {
      fence_proxy_async_smem();   // staging writes -> visible to the async proxy
      __syncwarp();
      if (u == 0) W2V_TR(7, trace_it);
      if (lane == 0) {
        tma_reduce_add_2d(tmap_o, sblk, col0, (int)orow0);   // rows past the end are clipped
        bulk_commit_group();
      }
      continue;
    }
    __syncwarp();
    if (u == 0) W2V_TR(7, trace_it);
    // staging -> global, coalesced: 4 rows x 128 B per warp instruction
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rr = k * 4 + c_row;
      uint4 x;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                   : "r"(stage + (uint32_t)(rr * 128 + ((c_chk ^ (rr & 7)) << 4))));
      if (rg0 + rr < p.rows_per_group) {
        if constexpr (OUT_F32) {
          float4 f = make_float4(__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z),
                                 __uint_as_float(x.w));
          if (use_resid) {
            const float4 rv = res[u & 1][k];
            f.x += rv.x; f.y += rv.y; f.z += rv.z; f.w += rv.w;
          }
          float* dst = reinterpret_cast<float*>(p.out) + (orow0 + rr) * p.ld_out + col0 + c_chk * 4;
          if (use_red)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(f.x), "f"(f.y),
                         "f"(f.z), "f"(f.w)
                         : "memory");
          else
            *reinterpret_cast<float4*>(dst) = f;
        } else {
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (orow0 + rr) * p.ld_out +
                                    col0 + c_chk * 8) = x;
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace w2v
