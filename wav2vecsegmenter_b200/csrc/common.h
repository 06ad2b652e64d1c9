// Host-side plumbing shared by all translation units of libw2vseg: error reporting (no
// exceptions cross the C ABI), the TMA tensor-map encoder (resolved through the runtime so the
// library has no link-time dependency on libcuda), and kernel-launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/w2vseg.h"

namespace w2v {

void set_error(const char* fmt, ...);
// running count of kernels launched by this library in this process (for bench accounting)
void count_launch(int n = 1);

#define W2V_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      w2v::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,        \
                     __LINE__);                                                               \
      return W2VSEG_ERR_CUDA;                                                                 \
    }                                                                                         \
  } while (0)

#define W2V_CHECK_LAUNCH()                                                                    \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      w2v::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__,    \
                     __LINE__);                                                               \
      return W2VSEG_ERR_CUDA;                                                                 \
    }                                                                                         \
    w2v::count_launch();                                                                      \
  } while (0)

#define W2V_REQUIRE(cond, ...)                                                                \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      w2v::set_error(__VA_ARGS__);                                                            \
      return W2VSEG_ERR_ARG;                                                                  \
    }                                                                                         \
  } while (0)

// Runs the following block at least once per process, thread-safely: the guarded work (cudaFuncSetAttribute) is
// idempotent, so two threads racing through it is harmless — only the flag itself must not be a data race.
#define W2V_ONCE_BEGIN                                  \
  {                                                     \
    static std::atomic<int> _once_done{0};              \
    if (_once_done.load(std::memory_order_acquire) == 0) {
#define W2V_ONCE_END                                    \
      _once_done.store(1, std::memory_order_release);   \
    }                                                   \
  }

#define W2V_TRY(expr)                                                                         \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != 0) return _r;                                                                   \
  } while (0)

// 2-D bf16 tensor map, 128-byte swizzle. dim0 = contiguous extent (elements), dim1 = rows,
// row_stride_elems = distance between consecutive rows (may be SMALLER than dim0: overlapping
// rows are how the strided convolutions are presented to the GEMM as an im2col view).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                      uint64_t row_stride_elems, uint32_t box0, uint32_t box1);

// same with the 64-byte swizzle (boxes whose rows are 64 bytes: the GEMM epilogue's output blocks)
int make_tmap_2d_bf16_sw64(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                           uint64_t row_stride_elems, uint32_t box0, uint32_t box1);

// fp32 elements, 128-byte swizzle (the residual GEMMs' 32 x 32 reduce-add boxes)
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                     uint64_t row_stride_elems, uint32_t box0, uint32_t box1);

// 3-D variant (box = box0 x box1 x 1): the third dimension isolates planes (e.g. windows) so that a
// box hanging over the end of one plane is clipped instead of spilling into the next one.
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1);

int num_sms();

// ---- optional per-kernel timing (bench.py roofline numbers; off by default) -------------------
// When enabled, every launcher brackets its kernel with CUDA events on the launching stream.
// The engine names the next launch with prof_tag(); launchers fall back to their own name.
bool prof_enabled();
void prof_tag(const char* tag);  // applies to the next ProfScope only
struct ProfScope {
  ProfScope(cudaStream_t s, const char* default_name);
  ~ProfScope();
  cudaStream_t stream;
  int slot;
};

}  // namespace w2v
