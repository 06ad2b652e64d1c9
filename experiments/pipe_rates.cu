// Microbenchmark: issue rates per SM sub-partition of the instructions the softmax inner loop is made of
// (sm_100a): MUFU.EX2, FFMA (3-register and with a uniform/constant operand), fma.rn.f32x2, FADD, add.f32x2,
// F2FP (cvt.rn.bf16x2.f32), FMNMX3. W warps per CTA, one CTA: warp w sits on sub-partition w % 4.
// Every instruction is part of a loop-carried chain (16 independent chains per thread), so nothing can be hoisted.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/pipe_rates experiments/pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 16
template <int OP>
__global__ void k(int iters, float a, float b, long long* cyc, float* sink) {
  float x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = a * (threadIdx.x + i) * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (OP == 0) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (OP == 1) {   // FFMA, 3 registers
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 1) % CHAINS]), "f"(x[(i + 2) % CHAINS]));
      } else if (OP == 2) {   // FFMA with kernel-parameter operands (uniform register / constant)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fmaf(x[i], a, b);
      } else if (OP == 3) {   // fma.rn.f32x2: two FMAs per instruction
#pragma unroll
        for (int i = 0; i < CHAINS; i += 2) {
          unsigned long long v, m, c;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(a), "f"(a));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(b), "f"(b));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(m), "l"(c));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
        }
      } else if (OP == 4) {   // FADD
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[(i + 1) % CHAINS]));
      } else if (OP == 5) {   // add.f32x2
#pragma unroll
        for (int i = 0; i < CHAINS; i += 2) {
          unsigned long long v, m;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(a), "f"(b));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(m));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
        }
      } else if (OP == 6) {   // F2FP pack
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          uint32_t p;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[i]), "f"(x[(i + 1) % CHAINS]));
          x[i] = __uint_as_float(p);
        }
      } else if (OP == 7) {   // FMNMX3
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 1) % CHAINS]), "f"(x[(i + 2) % CHAINS]));
      } else if (OP == 9) {   // MUFU.EX2.F16 on one half
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          unsigned short h = (unsigned short)__float_as_uint(x[i]);
          asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h));
          x[i] = __uint_as_float((uint32_t)h | 0x30000000u);
        }
      } else if (OP == 10) {  // MUFU.EX2.BF16 on one half
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          unsigned short h = (unsigned short)__float_as_uint(x[i]);
          asm volatile("ex2.approx.ftz.bf16 %0, %0;" : "+h"(h));
          x[i] = __uint_as_float((uint32_t)h | 0x30000000u);
        }
      } else if (OP == 11) {  // ex2.approx.f16x2: lowers to two MUFU.EX2.F16 + PRMT
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          uint32_t h = __float_as_uint(x[i]);
          asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
          x[i] = __uint_as_float(h & 0x3fff3fffu);
        }
      } else if (OP == 8) {   // the softmax element: FFMA(uniform) + MUFU + FADD + half an F2FP
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < CHAINS; i += 2) {
          float e0, e1;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(x[i], a, b)));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(x[i + 1], a, b)));
          sum += e0; sum += e1;
          uint32_t p;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(e0), "f"(e1));
          x[i] = __uint_as_float(p) * 1e-30f; x[i + 1] = sum * 1e-30f;
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
  sink[threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int per_iter, long long* cyc, float* sink) {
  const int iters = 500;
  printf("%-34s", name);
  for (int warps : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { k<OP><<<1, warps * 32>>>(iters, 0.999f, 1e-6f, cyc, sink); cudaDeviceSynchronize(); }
    long long h[16]; cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("  %dw/sp: %5.2f cyc/instr/sp", warps / 4, (double)mx / ((double)iters * per_iter * (warps / 4)));
  }
  printf("\n");
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 64 * 8); cudaMalloc(&sink, 4096);
  run<0>("MUFU.EX2", 4 * CHAINS, cyc, sink);
  run<1>("FFMA (3 registers)", 4 * CHAINS, cyc, sink);
  run<2>("FFMA (uniform operands)", 4 * CHAINS, cyc, sink);
  run<3>("fma.rn.f32x2 (per instr = 2 FMAs)", 4 * CHAINS / 2, cyc, sink);
  run<4>("FADD", 4 * CHAINS, cyc, sink);
  run<5>("add.rn.f32x2 (per instr = 2 adds)", 4 * CHAINS / 2, cyc, sink);
  run<6>("F2FP bf16x2 pack", 4 * CHAINS, cyc, sink);
  run<7>("FMNMX3", 4 * CHAINS, cyc, sink);
  run<8>("softmax element (per element)", 4 * CHAINS, cyc, sink);
  run<9>("MUFU.EX2.F16 (one half)", 4 * CHAINS, cyc, sink);
  run<10>("MUFU.EX2.BF16 (one half)", 4 * CHAINS, cyc, sink);
  run<11>("ex2.approx.f16x2 (per 2 exps)", 4 * CHAINS, cyc, sink);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
