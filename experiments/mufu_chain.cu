// NOTE (round 2): this first attempt UNDER-COUNTS the MUFU work — s[] and nm_hi do not change between iterations,
// so the compiler hoists half of the exponentials out of the timed loop (it reports ~4 cycles per MUFU.EX2 instead of
// the real 8). experiments/pipe_rates.cu (loop-carried chains) is the measurement to trust. Kept for the record.
// Microbenchmark: how long does ONE warp need for the softmax exponential phase of a 16x128 S slice
// (64 x [FFMA, MUFU.EX2, FADD] + 32 bf16x2 packs), alone on its SM sub-partition and with 1..3 sibling
// warps on the same sub-partition? Answers whether the ~1100-cycle exp phase of attention_tc64 is a
// per-warp latency chain or XU sharing.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/mufu_chain experiments/mufu_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

template <int MODE>
__global__ void k(int iters, const float* in, float scale, long long* cyc, uint32_t* sink) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[i * 32 + (threadIdx.x & 31)];
  float nm_lo = in[3], nm_hi = in[5];
  uint32_t acc = 0; float l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float sl0 = 0, sl1 = 0, sh0 = 0, sh1 = 0;
    uint32_t pk[32];
    if (MODE == 0) {
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float e0 = ex2a(fmaf(s[4 * kk + 0], scale, nm_lo)), e1 = ex2a(fmaf(s[4 * kk + 1], scale, nm_lo));
        const float e2 = ex2a(fmaf(s[4 * kk + 2], scale, nm_hi)), e3 = ex2a(fmaf(s[4 * kk + 3], scale, nm_hi));
        sl0 += e0; sl1 += e1; sh0 += e2; sh1 += e3;
        pk[2 * kk] = pack(e0, e1); pk[2 * kk + 1] = pack(e2, e3);
      }
    } else if (MODE == 1) {   // MUFU only (no sums, no packs): XU-issue floor
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float e0 = ex2a(fmaf(s[4 * kk + 0], scale, nm_lo)), e1 = ex2a(fmaf(s[4 * kk + 1], scale, nm_lo));
        const float e2 = ex2a(fmaf(s[4 * kk + 2], scale, nm_hi)), e3 = ex2a(fmaf(s[4 * kk + 3], scale, nm_hi));
        pk[2 * kk] = __float_as_uint(e0) ^ __float_as_uint(e1); pk[2 * kk + 1] = __float_as_uint(e2) ^ __float_as_uint(e3);
      }
    } else {                  // all 64 MUFUs first (results in place), then sums + packs
      float e[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) e[i] = ex2a(fmaf(s[i], scale, (i & 2) ? nm_hi : nm_lo));
      float z = e[63] * 0.f;  // sums cannot start before the last MUFU has issued
      sl0 = z; sl1 = z; sh0 = z; sh1 = z;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        sl0 += e[4 * kk]; sl1 += e[4 * kk + 1]; sh0 += e[4 * kk + 2]; sh1 += e[4 * kk + 3];
        pk[2 * kk] = pack(e[4 * kk] + z, e[4 * kk + 1]); pk[2 * kk + 1] = pack(e[4 * kk + 2] + z, e[4 * kk + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    l += sl0 + sl1 + sh0 + sh1;
    nm_lo += __uint_as_float(acc & 1);   // loop-carried so iterations cannot merge
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(l);
}

// mode 0 exp phase in warps 0..7 (2 per sub-partition) while warps 8..15 spin on an mbarrier exactly like
// ptx.cuh's mbar_wait (try_wait + clock64 watchdog): do waiting warps steal issue slots from the exponentials?
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void kspin(int iters, const float* in, float scale, long long* cyc, uint32_t* sink, int spin) {
  __shared__ uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(8));
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp >= 8) {
    if (spin) {
      long long t0 = clock64();
      while (!try_wait(b, 0)) { if (clock64() - t0 > 8000000000LL) __trap(); }
    }
    return;
  }
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[i * 32 + (threadIdx.x & 31)];
  float nm_lo = in[3], nm_hi = in[5];
  uint32_t acc = 0; float l = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float sl0 = 0, sl1 = 0, sh0 = 0, sh1 = 0;
    uint32_t pk[32];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float e0 = ex2a(fmaf(s[4 * kk + 0], scale, nm_lo)), e1 = ex2a(fmaf(s[4 * kk + 1], scale, nm_lo));
      const float e2 = ex2a(fmaf(s[4 * kk + 2], scale, nm_hi)), e3 = ex2a(fmaf(s[4 * kk + 3], scale, nm_hi));
      sl0 += e0; sl1 += e1; sh0 += e2; sh1 += e3;
      pk[2 * kk] = pack(e0, e1); pk[2 * kk + 1] = pack(e2, e3);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    l += sl0 + sl1 + sh0 + sh1;
    nm_lo += __uint_as_float(acc & 1);
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { cyc[warp] = t1 - t0; asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
  sink[threadIdx.x] = acc + __float_as_uint(l);
}

int main() {
  float* in; long long* cyc; uint32_t* sink;
  cudaMalloc(&in, 64 * 32 * 4); cudaMemset(in, 0, 64 * 32 * 4);
  cudaMalloc(&cyc, 64 * 8); cudaMalloc(&sink, 4 * 1024 * 4);
  const int iters = 200;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 12, 16}) {   // warps per CTA, 1 CTA: warp w sits on sub-partition w % 4
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(iters, in, 0.125f, cyc, sink);
        if (mode == 1) k<1><<<1, warps * 32>>>(iters, in, 0.125f, cyc, sink);
        if (mode == 2) k<2><<<1, warps * 32>>>(iters, in, 0.125f, cyc, sink);
        cudaDeviceSynchronize();
      }
      long long h[16]; cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mode %d warps/CTA %2d (%d per sub-partition): %.0f cycles per 64-exp phase per warp  -> XU busy %.0f %%\n", mode, warps,
             (warps + 3) / 4, (double)mx / iters, 100.0 * ((warps + 3) / 4) * 512.0 / ((double)mx / iters));
    }
  for (int spin = 0; spin < 2; ++spin) {
    for (int rep = 0; rep < 2; ++rep) { kspin<<<1, 512>>>(iters, in, 0.125f, cyc, sink, spin); cudaDeviceSynchronize(); }
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 8; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("8 exp warps + 8 %s warps: %.0f cycles per 64-exp phase per warp\n", spin ? "mbarrier-spinning" : "exited", (double)mx / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
