// Microbenchmark: tcgen05.ld throughput per SM on sm_100a (is d=64 attention TMEM-read-bound?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/tmem_ld_bw experiments/tmem_ld_bw.cu
// Each CTA allocates 512 TMEM columns; W warps (warp w reads lane quarter w%4) loop over
// tcgen05.ld.32x32b.x32 on 128 columns per iteration. Reports bytes/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../wav2vecsegmenter_b200/csrc/ptx.cuh"
using namespace w2v;

template <int MODE>
__global__ void k(int iters, unsigned long long* cyc, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {           // 4 x (x32 load), wait once
      uint32_t r[4][32];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld_32x32b_x32(base + j * 32, r[j]);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[j][i];
    } else {                   // load, wait, load, wait (dependent)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(base + j * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(iters, cyc, sink);
        else k<1><<<148, warps * 32>>>(iters, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double c = (double)h[0];
      double bytes = (double)warps * 32 * 128 * 4 * iters;
      printf("mode %d warps %2d: %.0f cycles, %.1f B/clk/SM (%.1f cyc per 128x128 fp32 tile-equivalent 64KB)\n",
             mode, warps, c, bytes / c, 65536.0 / (bytes / c));
    }
  return 0;
}
