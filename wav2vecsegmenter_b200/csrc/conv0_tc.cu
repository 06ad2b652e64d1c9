// conv layer 0 (1 -> 512 channels, k = 10, stride 5) + LayerNorm(512) + GELU as ONE tcgen05 kernel
// (HF:281-299, Wav2Vec2LayerNormConvLayer with Cin = 1; input normalisation lib/datautils.py:122-125
// applied on the fly). Replaces conv0_ln_gelu_kernel (kernels.cu), which is kept as the second
// implementation for the parity tests.
//
// The layer is not GEMM-bound (K = 10: 14.7 GF per 14-window batch); what it costs is the ~50 CUDA
// core instructions per output element of the straightforward version (10 FMAs, two LayerNorm
// passes, affine, GELU) over 459 M elements, against a floor of 0.14 ms for writing its 917 MB
// output. Here everything except the GELU happens inside the MMA:
//
//   LayerNorm statistics need not be measured from the 512 outputs of a frame: with one input
//   channel they are a closed form of the frame's 10 samples. With z = [x_0..x_9, 1] and the
//   channel-CENTRED weights wt[c] = [w[c,:] - mean_c w, b_c - mean_c b]:
//       y_c - mean_c(y) = wt[c] . z                      (no mean to subtract afterwards)
//       var_c(y)        = z^T G z,  G = wt^T wt / 512    (11 x 11, positive semi-definite)
//   G is factored once at weight-pack time (G = U^T U, Cholesky in fp64), so the variance is a sum of
//   11 squares |U z|^2 — no cancellation — for 66 FMAs per FRAME on one producer warp.
//
//   Then LN(y)_c = gamma_c * rstd * (wt[c] . z) + beta_c is itself a K = 12 dot product:
//       A row (per frame, fp16)   = [rstd*x_0, .., rstd*x_9, rstd, 1, 0, 0, 0, 0]
//       W row (per channel, fp16) = [gamma_c*wt[c,0..9], gamma_c*wt[c,10], beta_c, 0, 0, 0, 0]
//   i.e. ONE tcgen05.mma (M=128, N=256, K=16, kind::f16 with fp16 operands: 11-bit mantissas for the
//   raw audio) per 128-frame x 256-channel tile leaves the LayerNorm output in TMEM and the epilogue
//   is GELU + bf16 pack + coalesced store: ~10 instructions per element.
//
// CTA = 18 warps, persistent, one per SM:
//   warp 0       producer: audio -> normalised samples -> rstd -> fp16 A tile (128 x 16) in the
//                128-byte-swizzled K-major layout tcgen05 expects (only the first 32 bytes of each
//                128-byte row are used: same descriptors as the GEMM kernels, K sub-block 0)
//   warp 1       MMA issuer: per frame block two MMAs (channel halves) into the two TMEM accumulators
//   warps 2..17  epilogue: 4 TMEM lane quarters x 4 column groups of 64; TMEM is released right after
//                the tcgen05.ld, so the MMA of the next tile overlaps the GELU/stores of this one
// The packed weights (512 x 32 B) stay resident in shared memory for the whole kernel.
#include <cuda_fp16.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace w2v {

namespace {

constexpr int C0T_ROWS = 128;                  // frames per block (MMA M)
constexpr int C0T_THREADS = 576;
constexpr int C0T_EPI_WARPS = 16;
constexpr int C0T_W_BYTES = 512 * 128;         // weights: 512 swizzle rows
constexpr int C0T_A_BYTES = C0T_ROWS * 128;    // one A tile
constexpr int C0T_STAGE_BYTES = C0T_EPI_WARPS * 4096;
constexpr int C0T_XS = C0T_ROWS * 5 + 5;       // samples one frame block reads (645)
constexpr int C0T_XPL = (C0T_XS + 31) / 32;    // per producer lane (21)
constexpr int C0T_MISC_BYTES = 4096;
constexpr int C0T_SMEM = 1024 + C0T_W_BYTES + 2 * C0T_A_BYTES + C0T_STAGE_BYTES + C0T_MISC_BYTES;
constexpr int C0T_NU = 66;                     // upper triangle of the 11 x 11 factor

// kind::f16 instruction descriptor with fp16 A/B (format 0), fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ double block_sum_512(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = (lane < 16) ? sh[lane] : 0.0;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return __shfl_sync(0xffffffffu, t, 0);
}

// One block of 512 threads (thread = channel). w element (c, k) at w[c*sc + k*sk].
__global__ void __launch_bounds__(512)
conv0_pack_kernel(const float* __restrict__ w, int sc, int sk, const float* __restrict__ bias,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  __half* __restrict__ wp, float* __restrict__ u_out) {
  __shared__ double sh[16];
  __shared__ double g_s[11][11];
  const int c = threadIdx.x;
  double v[11];
#pragma unroll
  for (int k = 0; k < 10; ++k) v[k] = (double)w[(long long)c * sc + (long long)k * sk];
  v[10] = (double)bias[c];
#pragma unroll
  for (int k = 0; k < 11; ++k) v[k] -= block_sum_512(v[k], sh) * (1.0 / 512.0);
#pragma unroll
  for (int i = 0; i < 11; ++i) {
#pragma unroll
    for (int j = i; j < 11; ++j) {
      const double s = block_sum_512(v[i] * v[j], sh) * (1.0 / 512.0);
      if (c == 0) g_s[i][j] = s;
    }
  }
  const float ga = gamma[c];
  uint4 c0, c1;
  c0.x = pack_half2(ga * (float)v[0], ga * (float)v[1]);
  c0.y = pack_half2(ga * (float)v[2], ga * (float)v[3]);
  c0.z = pack_half2(ga * (float)v[4], ga * (float)v[5]);
  c0.w = pack_half2(ga * (float)v[6], ga * (float)v[7]);
  c1.x = pack_half2(ga * (float)v[8], ga * (float)v[9]);
  c1.y = pack_half2(ga * (float)v[10], beta[c]);
  c1.z = 0u;
  c1.w = 0u;
  reinterpret_cast<uint4*>(wp)[c * 2 + 0] = c0;
  reinterpret_cast<uint4*>(wp)[c * 2 + 1] = c1;
  __syncthreads();
  if (c == 0) {
    // G = U^T U, U upper triangular; a vanishing pivot (G is only semi-definite, e.g. constant bias
    // and a tap no channel uses) zeroes its row, which is exact for a PSD matrix
    double U[11][11];
    for (int i = 0; i < 11; ++i)
      for (int j = 0; j < 11; ++j) U[i][j] = 0.0;
    double dmax = 0.0;
    for (int i = 0; i < 11; ++i) dmax = fmax(dmax, g_s[i][i]);
    for (int i = 0; i < 11; ++i) {
      double d = g_s[i][i];
      for (int k = 0; k < i; ++k) d -= U[k][i] * U[k][i];
      if (d > 1e-13 * dmax && d > 0.0) {
        const double r = sqrt(d);
        U[i][i] = r;
        for (int j = i + 1; j < 11; ++j) {
          double s = g_s[i][j];
          for (int k = 0; k < i; ++k) s -= U[k][i] * U[k][j];
          U[i][j] = s / r;
        }
      }
    }
    int idx = 0;
    for (int i = 0; i < 11; ++i)
      for (int j = i; j < 11; ++j) u_out[idx++] = (float)U[i][j];
  }
}

__global__ void __launch_bounds__(C0T_THREADS, 1)
conv0_tc_kernel(const float* __restrict__ audio, long long audio_stride,
                const int* __restrict__ sample_len, const float2* __restrict__ stats,
                const __half* __restrict__ wp, const float* __restrict__ u_g, float eps,
                __nv_bfloat16* __restrict__ out, int R0, int fb_per_window, int total_fb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + C0T_W_BYTES;
  uint8_t* smem_stage = smem_a + 2 * C0T_A_BYTES;
  uint8_t* misc = smem_stage + C0T_STAGE_BYTES;
  float* x_s = reinterpret_cast<float*>(misc);                 // 645 floats
  float* u_s = reinterpret_cast<float*>(misc + 2688);          // 66 floats
  uint64_t* a_full = reinterpret_cast<uint64_t*>(misc + 3072); // [2]
  uint64_t* a_free = a_full + 2;                               // [2]
  uint64_t* acc_full = a_free + 2;                             // [2]
  uint64_t* acc_free = acc_full + 2;                           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // packed weights -> swizzled rows (16-byte chunk j of row r lives at chunk j ^ (r & 7))
  for (int i = threadIdx.x; i < 1024; i += C0T_THREADS) {
    const int row = i >> 1, j = i & 1;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wp) + i);
    *reinterpret_cast<uint4*>(smem_w + row * 128 + ((j ^ (row & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  if (threadIdx.x < C0T_NU) u_s[threadIdx.x] = __ldg(u_g + threadIdx.x);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_free[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_free[s], C0T_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    // raw samples of the NEXT frame block are fetched while the current one is processed
    float xr[C0T_XPL];
    uint32_t valid = 0;
    float2 st = make_float2(0.f, 1.f);
    auto fetch = [&](int fb) {
      const int b = fb / fb_per_window;
      const long long s0 = (long long)(fb - b * fb_per_window) * (C0T_ROWS * 5);
      const int len = sample_len[b];
      const float* x = audio + (long long)b * audio_stride;
      st = stats[b];
      valid = 0;
#pragma unroll
      for (int m = 0; m < C0T_XPL; ++m) {
        const int idx = m * 32 + lane;
        const long long sidx = s0 + idx;
        const bool ok = idx < C0T_XS && sidx < len;
        xr[m] = ok ? __ldg(x + sidx) : 0.f;
        valid |= (uint32_t)ok << m;
      }
    };
    if ((int)blockIdx.x < total_fb) fetch(blockIdx.x);
    int i = 0;
    for (int fb = blockIdx.x; fb < total_fb; fb += gridDim.x, ++i) {
      const int abuf = i & 1;
#pragma unroll
      for (int m = 0; m < C0T_XPL; ++m) {
        const int idx = m * 32 + lane;
        if (idx < C0T_XS) x_s[idx] = ((valid >> m) & 1u) ? (xr[m] - st.x) * st.y : 0.f;
      }
      __syncwarp();
      if (fb + (int)gridDim.x < total_fb) fetch(fb + gridDim.x);
      mbar_wait(&a_free[abuf], (uint32_t)(((i >> 1) & 1) ^ 1));
      uint8_t* a_buf = smem_a + abuf * C0T_A_BYTES;
#pragma unroll 1
      for (int rr = 0; rr < C0T_ROWS / 32; ++rr) {
        const int row = rr * 32 + lane;
        float xv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) xv[k] = x_s[row * 5 + k];
        float var = 0.f;
        int idx = 0;
#pragma unroll
        for (int ii = 0; ii < 11; ++ii) {
          float s = u_s[idx + (10 - ii)];
#pragma unroll
          for (int j = ii; j < 10; ++j) s = fmaf(u_s[idx + (j - ii)], xv[j], s);
          idx += 11 - ii;
          var = fmaf(s, s, var);
        }
        const float rstd = rsqrtf(var + eps);
        // fp16 range guard: rstd * x can only leave it when x is huge along a direction the taps ignore
        // (a tap no channel uses); unclamped, inf * 0 would turn the frame into NaNs
#pragma unroll
        for (int k = 0; k < 10; ++k) xv[k] = fminf(fmaxf(xv[k] * rstd, -60000.f), 60000.f);
        uint4 c0, c1;
        c0.x = pack_half2(xv[0], xv[1]);
        c0.y = pack_half2(xv[2], xv[3]);
        c0.z = pack_half2(xv[4], xv[5]);
        c0.w = pack_half2(xv[6], xv[7]);
        c1.x = pack_half2(xv[8], xv[9]);
        c1.y = pack_half2(rstd, 1.f);
        c1.z = 0u;
        c1.w = 0u;
        *reinterpret_cast<uint4*>(a_buf + row * 128 + ((0 ^ (row & 7)) << 4)) = c0;
        *reinterpret_cast<uint4*>(a_buf + row * 128 + ((1 ^ (row & 7)) << 4)) = c1;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[abuf]);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(C0T_ROWS, 256);
    int i = 0;
    for (int fb = blockIdx.x; fb < total_fb; fb += gridDim.x, ++i) {
      const int abuf = i & 1;
      mbar_wait(&a_full[abuf], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint64_t a_desc = make_desc_k_sw128(smem_u32(smem_a + abuf * C0T_A_BYTES));
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        mbar_wait(&acc_free[hf], (uint32_t)((i & 1) ^ 1));
        tc_fence_after();
        const uint64_t b_desc = make_desc_k_sw128(smem_u32(smem_w + hf * 256 * 128));
        if (elect_one()) {
          tc_mma_ss(tmem_base + (uint32_t)(hf * 256), a_desc, b_desc, idesc, 0u);
          tc_commit(&acc_full[hf]);
          if (hf == 1) tc_commit(&a_free[abuf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..17)
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int cg = (warp - 2) >> 2;            // 64-column group inside the 256-column accumulator
    const uint32_t stage = smem_u32(smem_stage + (warp - 2) * 4096);
    const int c_row = lane >> 3, c_chk = lane & 7;
    int i = 0;
    for (int fb = blockIdx.x; fb < total_fb; fb += gridDim.x, ++i) {
      const int b = fb / fb_per_window;
      const int t0 = (fb - b * fb_per_window) * C0T_ROWS + q * 32;
      __nv_bfloat16* orow = out + ((long long)b * R0 + t0) * 512 + cg * 64 + c_chk * 8;
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        mbar_wait(&acc_full[hf], (uint32_t)(i & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 256 + cg * 64);
        uint32_t raw[64];
        tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
        tmem_ld_32x32b_x32(taddr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&raw[32]));
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free[hf]);   // accumulator is in registers: hand it back
        const uint32_t srow = stage + (uint32_t)(lane * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t w0 = pack_bf16x2(gelu_erf(__uint_as_float(raw[8 * j + 0])),
                                          gelu_erf(__uint_as_float(raw[8 * j + 1])));
          const uint32_t w1 = pack_bf16x2(gelu_erf(__uint_as_float(raw[8 * j + 2])),
                                          gelu_erf(__uint_as_float(raw[8 * j + 3])));
          const uint32_t w2 = pack_bf16x2(gelu_erf(__uint_as_float(raw[8 * j + 4])),
                                          gelu_erf(__uint_as_float(raw[8 * j + 5])));
          const uint32_t w3 = pack_bf16x2(gelu_erf(__uint_as_float(raw[8 * j + 6])),
                                          gelu_erf(__uint_as_float(raw[8 * j + 7])));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)((j ^ (lane & 7)) << 4)),
                       "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                       : "memory");
        }
        __syncwarp();
        // staging -> global: 4 rows x 128 B per warp instruction
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = k * 4 + c_row;
          uint4 x;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                       : "r"(stage + (uint32_t)(rr * 128 + ((c_chk ^ (rr & 7)) << 4))));
          if (t0 + rr < R0) *reinterpret_cast<uint4*>(orow + (long long)rr * 512 + hf * 256) = x;
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

size_t conv0_tc_pack_bytes() { return (size_t)512 * 16 * sizeof(__half) + 128 * sizeof(float); }

int conv0_tc_pack_launch(const float* w, int sc, int sk, const float* bias, const float* gamma,
                         const float* beta, void* pack, cudaStream_t s) {
  __half* wp = reinterpret_cast<__half*>(pack);
  float* u = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(pack) + (size_t)512 * 16 * sizeof(__half));
  conv0_pack_kernel<<<1, 512, 0, s>>>(w, sc, sk, bias, gamma, beta, wp, u);
  W2V_CHECK_LAUNCH();
  return 0;
}

int conv0_tc_launch(const float* audio, int64_t audio_stride, const int32_t* sample_len,
                    const float2* stats, const void* pack, float eps, __nv_bfloat16* out, int B,
                    int R0, cudaStream_t s) {
  if (B <= 0 || R0 <= 0) return 0;
  const __half* wp = reinterpret_cast<const __half*>(pack);
  const float* u = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(pack) +
                                                  (size_t)512 * 16 * sizeof(__half));
  W2V_ONCE_BEGIN
    W2V_CHECK_CUDA(cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C0T_SMEM));
  W2V_ONCE_END
  const int fb_per_window = (R0 + C0T_ROWS - 1) / C0T_ROWS;
  const long long total = (long long)B * fb_per_window;
  W2V_REQUIRE(total < (1ll << 30), "conv0: too many frame blocks (%lld)", total);
  const int grid = (int)(total < (long long)num_sms() ? total : (long long)num_sms());
  ProfScope ps(s, "conv0_ln_gelu");
  conv0_tc_kernel<<<grid, C0T_THREADS, C0T_SMEM, s>>>(audio, audio_stride, sample_len, stats, wp, u,
                                                       eps, out, R0, fb_per_window, (int)total);
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
