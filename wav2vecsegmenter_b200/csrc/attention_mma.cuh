// Warp-level mma.sync (m16n8k16 bf16 -> fp32) building blocks shared by the mma.sync attention forward
// (attention.cu: second implementation for tests + the training forward that also returns the row
// log-sum-exp) and the attention backward of the head training step (attention_bwd.cu): cp.async tile loads
// into XOR-swizzled shared memory, ldmatrix operand fetch, the MMA itself.
#pragma once
#include "ptx.cuh"

namespace w2v {
namespace {

constexpr int ATT_BQ = 64;     // query rows per CTA (16 per warp)
constexpr int ATT_BKV = 64;    // keys per tile
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 zero bytes written
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                              uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][DH] bf16 tile, XOR-swizzled so
// that ldmatrix (8 rows x one chunk) and the row-contiguous cp.async fills are conflict-free
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return (uint32_t)(row * (DH * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int DH>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* gbase,
                                          long long ld, int row0, int rows_valid) {
  constexpr int CHUNKS = DH / 8;
  for (int i = threadIdx.x; i < ATT_BKV * CHUNKS; i += ATT_THREADS) {
    const int r = i / CHUNKS, c = i - r * CHUNKS;
    const bool ok = (row0 + r) < rows_valid;
    const __nv_bfloat16* src = gbase + (long long)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(smem_base + tile_off<DH>(r, c), src, ok);
  }
}


}  // namespace
}  // namespace w2v
