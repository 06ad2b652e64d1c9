// Fused non-causal attention for head_dim 64 (the 16 x 64 encoder layers): tcgen05 / TMEM / TMA.
// Persistent CTAs (two per SM) walk work items (window, head, 128-row query tile); the key tiles of
// all items of a CTA form ONE stream, so the K/V loads and the first QK^T of the next item overlap
// the tail of the current one.
//
//   warps 0..7  softmax. Warp w owns 16 query rows (TMEM lanes 32(w%4) + 16(w/4) ..) x all 128 keys
//               of the tile in the 16x256b TMEM shape: a row lives in one quad, so the row max / sum
//               are two shuffles (no cross-warp exchange) and the 64 scores of a thread stay in
//               registers between the max and the exponentials (ONE TMEM read per score, exact
//               max, no second pass). P goes back to TMEM as packed bf16x2 (tcgen05.st 16x128b).
//   warp 8      TMA producer (Q double-buffered per item, K/V per tile) + MMA issuer; the whole warp
//               runs the control flow (uniform registers), one elected lane issues.
//   TMEM        S 128 columns fp32 | O 64 columns fp32 | P 64 columns bf16x2 (A operand of the
//               TS-form P V MMA; V is consumed in place as an MN-major B operand)
//   output      O / l -> bf16 -> the item's own (retired) Q buffer -> one TMA store per item
//               through a [B][R][D] map that clips the rows past the end of a window.
//
// What the measurements said (profiles/attention_experiments_r01.md, section "session 2"):
//  * tcgen05.ld: one warp sustains only ~44 B/clk, 16 warps 477 B/clk per SM -> 8 softmax warps.
//  * P in shared memory costs ~1150 cycles of smem bandwidth per 128x128 tile per SM (more than the
//    1024 cycles of MUFU work) -> P in TMEM: 121 -> 113 us.
//  * S released right after the TMEM load, PV(g-1) awaited only right before the P store: 109 us.
//  * row-strided 16-byte output stores from registers cost ~2000 cycles per item -> TMA store;
//    the per-item global load of the key length ~2000 more -> key lengths in shared memory.
//  * persistent tile stream + the above: 106 us (B=14, R=1000; the 4-warp two-pass kernel: 124 us).
// Rescaling is lazy (FlashAttention-4): the exponent's running max only advances when a tile max
// exceeds it by more than 2^8. Replaces Wav2Vec2Attention's softmax(QK^T*scale + key mask) V
// (HF:500-549).
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace w2v {

namespace {

constexpr int A6_BM = 128;        // query rows per CTA
constexpr int A6_BN = 128;        // keys per tile
constexpr int A6_DH = 64;
constexpr int A6_THREADS = 288;   // 8 softmax warps + 1 TMA/MMA warp
constexpr int A6_CTRL_WARP = 8;
constexpr int A6_TILE_BYTES = A6_BN * A6_DH * 2;       // 16 KB: one Q / K / V tile
constexpr int A6_OFF_Q = 0;                            // 2 buffers (item parity)
constexpr int A6_OFF_K = A6_OFF_Q + 2 * A6_TILE_BYTES; // 2 buffers (tile parity)
constexpr int A6_OFF_V = A6_OFF_K + 2 * A6_TILE_BYTES; // 2 buffers
constexpr int A6_OFF_BAR = A6_OFF_V + 2 * A6_TILE_BYTES;   // 12 mbarriers + TMEM slot
// the kernel traps if the dynamic smem base is not 1024-byte aligned (no alignment slack)
constexpr int A6_OFF_KLEN = A6_OFF_BAR + 128;          // int32 [A6_MAX_B] clamped key lengths
constexpr int A6_MAX_B = 1024;
constexpr int A6_SMEM_BYTES = A6_OFF_KLEN + A6_MAX_B * 4;
constexpr int A6_TMEM_COLS = 256;
constexpr int A6_O_COL = 128;
constexpr int A6_P_COL = 192;     // P as packed bf16x2: 64 columns = 128 keys (A operand of P V)
constexpr float A6_RESCALE_THRESHOLD = 8.0f;           // log2 units
// Groups of 4 exponentials started OPTIMISTICALLY (with the running max) before the tile's own max is known.
// 0 = off, the shipped setting: at the 96-register cap of two CTAs per SM the extra live values spill
// (236-448 B) and the kernel gets SLOWER: 100 us -> 125 / 133 / 141 / 160 us for 2 / 4 / 6 / 8 groups
// (profiles/attention_experiments_r02.md). Kept as a build-time switch for a layout with more registers.
#ifndef A6_OPT_GROUPS
#define A6_OPT_GROUPS 0
#endif

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// TMEM access in the 16-lane shapes: a warp owns 16 query rows x all columns; thread t holds rows
// t/4 and t/4+8 and, of every 8-column group k, columns 8k + 2(t%4) and +1:
//   16x256b register 4k + e : row t/4 + 8(e>>1), column 8k + 2(t%4) + (e&1)
//   16x128b register 2k + e : row t/4 + 8e,      32-bit column 4k + t%4   (= the packed pair above)
// so a row lives in one quad (max / sum by two shuffles, no cross-warp exchange) and the bf16x2
// pairs a thread computes are exactly the P words it stores.
__device__ __forceinline__ void tmem_ld_16x256b_x16(uint32_t taddr, float (&r)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31]), "=f"(r[32]), "=f"(r[33]), "=f"(r[34]), "=f"(r[35]), "=f"(r[36]), "=f"(r[37]), "=f"(r[38]), "=f"(r[39]), "=f"(r[40]), "=f"(r[41]), "=f"(r[42]), "=f"(r[43]), "=f"(r[44]), "=f"(r[45]), "=f"(r[46]), "=f"(r[47]), "=f"(r[48]), "=f"(r[49]), "=f"(r[50]), "=f"(r[51]), "=f"(r[52]), "=f"(r[53]), "=f"(r[54]), "=f"(r[55]), "=f"(r[56]), "=f"(r[57]), "=f"(r[58]), "=f"(r[59]), "=f"(r[60]), "=f"(r[61]), "=f"(r[62]), "=f"(r[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x16(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
               : "memory");
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// -DA6_TRACE: clock64 timeline of softmax warp 0 of CTA 5 for stream tiles 16..23, printed at exit
#ifdef A6_TRACE
#ifndef A6_TRACE_CTA
#define A6_TRACE_CTA 5
#endif
__device__ long long g_trace[8][8];
__device__ long long g_tile_t[4][64];
__device__ long long g_ep[8];
#define A6_E(ev) do { if (blockIdx.x == 5 && threadIdx.x == 0 && seq == 2) g_ep[ev] = clock64() - t_cta0; } while (0)
   // per-tile s_full-seen stamps of 4 sample CTAs
#define A6_T(ev)                                                                                   \
  do {                                                                                             \
    if (blockIdx.x == A6_TRACE_CTA && threadIdx.x == 0 && g >= 20 && g < 28) g_trace[g - 20][ev] = clock64(); \
  } while (0)
#else
#define A6_T(ev) do {} while (0)
#define A6_E(ev) do {} while (0)
#endif

// One work item = (window b, head, 128-row query tile). A CTA walks items blockIdx.x, +gridDim.x, ...
// Item coordinates advance by a host-precomputed (qt, head, b) step with carries: an integer division
// per item change costs ~1500 cycles here, because I2F / MUFU.RCP / F2I queue behind the exponentials
// on the XU pipe (clock64 timeline, profiles/attention_experiments_r02.md).
struct Item {
  int q0, head, b, klen, n_tiles, row_base;
};
struct Step {
  int qt, head, b;      // idx -> (qt, head, b) decomposition of the grid stride
};
struct Pos {
  int idx, qt, head, b;
};
__device__ __forceinline__ Pos pos_first(int idx, int n_qt, int heads) {
  Pos p;
  p.idx = idx;
  p.qt = idx % n_qt;
  const int bh = idx / n_qt;
  p.head = bh % heads;
  p.b = bh / heads;
  return p;
}
__device__ __forceinline__ void pos_advance(Pos& p, const Step& st, int stride, int n_qt, int heads) {
  p.idx += stride;
  p.qt += st.qt;
  int c = 0;
  if (p.qt >= n_qt) { p.qt -= n_qt; c = 1; }
  p.head += st.head + c;
  c = 0;
  if (p.head >= heads) { p.head -= heads; c = 1; }
  p.b += st.b + c;
}
// kv_len points at the CTA's shared-memory copy of the (clamped) key lengths: a global load here
// sits on the critical path of every item change (~2000 cycles under the TMA traffic, measured)
__device__ __forceinline__ Item make_item(const Pos& p, int R, const int* kv_len) {
  Item it;
  it.head = p.head;
  it.b = p.b;
  it.q0 = p.qt * A6_BM;
  it.klen = kv_len[it.b];
  it.n_tiles = (it.klen + A6_BN - 1) / A6_BN;
  it.row_base = it.b * R;
  return it;
}
// Position in this CTA's stream of key tiles (items with no valid key contribute no tile).
struct Cursor {
  Pos pos;      // item (pos.idx >= n_items: end of stream)
  int seq;      // number of non-empty items before this one (Q / row-sum buffer parity)
  int j;        // key tile within the item
  Item it;
};
struct Walk {
  int n_items, stride, n_qt, heads, R;
  Step st;
  const int* kv_len;
};
__device__ __forceinline__ void cursor_skip_empty(Cursor& c, const Walk& w) {
  while (c.pos.idx < w.n_items) {
    c.it = make_item(c.pos, w.R, w.kv_len);
    if (c.it.n_tiles > 0) return;
    pos_advance(c.pos, w.st, w.stride, w.n_qt, w.heads);
  }
}
__device__ __forceinline__ void cursor_next(Cursor& c, const Walk& w) {
  if (++c.j < c.it.n_tiles) return;
  c.j = 0;
  c.seq += 1;
  pos_advance(c.pos, w.st, w.stride, w.n_qt, w.heads);
  cursor_skip_empty(c, w);
}

// 96 registers is the most two resident CTAs of 9 warps can have: the register file is allocated in units of two
// warps, and with 104 registers (__maxnreg__) only ONE CTA fits per SM (measured: 99 -> 138 us).
__global__ void __launch_bounds__(A6_THREADS, 2)
attention_tc64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                      int R, int heads, int n_qt, int n_items, int B, Step step, const int* __restrict__ kv_len_g,
                      float scale_log2, __nv_bfloat16* __restrict__ ctx) {
  extern __shared__ uint8_t att_raw[];
  uint8_t* smem = att_raw;
  if ((smem_u32(att_raw) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("w2vseg: attention smem base not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + A6_OFF_Q;
  uint8_t* sK = smem + A6_OFF_K;
  uint8_t* sV = smem + A6_OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A6_OFF_BAR);
  uint64_t* q_full = bars + 0;   // [2]
  uint64_t* k_full = bars + 2;   // [2]
  uint64_t* v_full = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* pv_done = bars + 9;
  uint64_t* o_free = bars + 10;
  uint64_t* stage_free = bars + 11;   // completion k: the output store of item k has read its staging tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * A6_DH;
  const int stride = gridDim.x;
  int* kv_len = reinterpret_cast<int*>(smem + A6_OFF_KLEN);
  const Walk walk = {n_items, stride, n_qt, heads, R, step, kv_len};
  for (int i = threadIdx.x; i < B; i += A6_THREADS) kv_len[i] = min(__ldg(kv_len_g + i), R);
#ifdef A6_TRACE
  const long long t_cta0 = clock64();
#endif

  if (warp == A6_CTRL_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_out);
      for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);   // q_full, k_full, v_full
      mbar_init(s_full, 1);
      mbar_init(s_free, 8);      // one elected arrive per softmax warp
      mbar_init(p_full, 8);
      mbar_init(pv_done, 1);
      mbar_init(o_free, 8);
      mbar_init(stage_free, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, A6_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + A6_O_COL;

  if (warp == A6_CTRL_WARP) {
    // ------------------------------------------------------------ TMA producer + MMA issuer
    // Every lane runs the warp-uniform control flow (descriptors in uniform registers); one elected
    // lane issues the TMA / MMA / commit instructions. The key tiles of all items of this CTA form
    // ONE stream g = 0, 1, ...: K/V buffers, S, P and their barriers are indexed by g, so the loads
    // and the first QK^T of the next item overlap the last tiles and the epilogue of the current one.
    auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col0, int row) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar, A6_TILE_BYTES);
        tma_load_2d(dst, &tmap_qkv, bar, col0, row);
      }
      __syncwarp();
    };
    Cursor Lc;   // next tile to LOAD (runs two tiles ahead of the PV cursor)
    Lc.pos = pos_first(blockIdx.x, n_qt, heads); Lc.seq = 0; Lc.j = 0;
    cursor_skip_empty(Lc, walk);
    Cursor Sc = Lc;   // next tile whose QK^T is to be issued
    Cursor Pc = Lc;   // next tile whose PV is to be issued
    // K(g) and V(g) are loaded separately; Q of an item goes with its first tile
    auto load_q = [&](const Cursor& c) {
      // The Q buffer doubles as the output staging tile of the item that used it two items ago: wait
      // until that item's output store has read it. The softmax warps signal this at the start of
      // the NEXT item, i.e. after an item epilogue that needs PV of that item's last tile — so this
      // wait must only be reached after that PV has been issued (see the main loop), or a one-tile
      // item in between deadlocks the CTA.
      if (c.seq >= 2) mbar_wait(stage_free, (uint32_t)((c.seq - 2) & 1));
      load_tile(sQ + (c.seq & 1) * A6_TILE_BYTES, &q_full[c.seq & 1], c.it.head * A6_DH, c.it.row_base + c.it.q0);
    };
    auto load_k = [&](const Cursor& c, int g) {
      load_tile(sK + (g & 1) * A6_TILE_BYTES, &k_full[g & 1], D + c.it.head * A6_DH, c.it.row_base + c.j * A6_BN);
    };
    auto load_v = [&](const Cursor& c, int g) {
      load_tile(sV + (g & 1) * A6_TILE_BYTES, &v_full[g & 1], 2 * D + c.it.head * A6_DH, c.it.row_base + c.j * A6_BN);
    };
    constexpr uint32_t idesc_s = make_idesc_bf16(A6_BM, A6_BN);
    constexpr uint32_t idesc_o = make_idesc_bf16(A6_BM, A6_DH) | (1u << 16);  // B (=V) MN-major
    const uint64_t q_desc0 = make_desc_k_sw128(smem_u32(sQ));
    const uint64_t k_desc0 = make_desc_k_sw128(smem_u32(sK));
    const uint64_t v_desc0 = desc_mn_sw128(smem_u32(sV), A6_BN * 128);
    auto issue_s = [&](const Cursor& c, int g) {     // S(g) = Q K^T
      mbar_wait(&k_full[g & 1], (uint32_t)((g >> 1) & 1));
      if (c.j == 0) mbar_wait(&q_full[c.seq & 1], (uint32_t)((c.seq >> 1) & 1));
      if (g > 0) mbar_wait(s_free, (uint32_t)((g - 1) & 1));   // S(g-1) is in the softmax registers
      tc_fence_after();
      const uint64_t qd = q_desc0 + (uint64_t)(((c.seq & 1) * A6_TILE_BYTES) >> 4);
      const uint64_t kd = k_desc0 + (uint64_t)(((g & 1) * A6_TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < A6_DH / 16; ++kk)
          tc_mma_ss(tS, qd + (uint64_t)((kk * 32) >> 4), kd + (uint64_t)((kk * 32) >> 4), idesc_s,
                    (uint32_t)(kk != 0));
        tc_commit(s_full);
      }
      __syncwarp();
    };

    if (Lc.pos.idx < n_items) {
      // prologue: tiles 0 and 1 in flight, S(0) issued
      load_q(Lc);
      load_k(Lc, 0);
      load_v(Lc, 0);
      cursor_next(Lc, walk);
      if (Lc.pos.idx < n_items) {
        if (Lc.j == 0) load_q(Lc);
        load_k(Lc, 1);
        load_v(Lc, 1);
        cursor_next(Lc, walk);
      }
      issue_s(Sc, 0);
      cursor_next(Sc, walk);

      for (int g = 0; Pc.pos.idx < n_items; ++g) {
        // invariant: Pc = tile g, Sc = tile g+1, Lc = tile g+2
        if (Sc.pos.idx < n_items) {
          issue_s(Sc, g + 1);
          cursor_next(Sc, walk);
          // S(g) has retired (S(g+1) was issued after s_free(g)), so K buffer g&1 can be refilled
          if (Lc.pos.idx < n_items) load_k(Lc, g + 2);
        }
        mbar_wait(&v_full[g & 1], (uint32_t)((g >> 1) & 1));
        mbar_wait(p_full, (uint32_t)(g & 1));        // P(g) in TMEM, O rescaled if it had to be
        if (Pc.j == 0 && Pc.seq > 0)                 // O of the previous item is in the epilogue's registers
          mbar_wait(o_free, (uint32_t)((Pc.seq - 1) & 1));
        tc_fence_after();
        {
          const uint64_t vd = v_desc0 + (uint64_t)(((g & 1) * A6_TILE_BYTES) >> 4);
          const uint32_t first = (uint32_t)(Pc.j != 0);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < A6_BN / 16; ++kk)
              tc_mma_ts(tO, tmem_base + A6_P_COL + (uint32_t)(kk * 8), vd + (uint64_t)((kk * 16 * 128) >> 4),
                        idesc_o, (kk != 0) ? 1u : first);
            tc_commit(pv_done);
          }
          __syncwarp();
        }
        cursor_next(Pc, walk);
        if (Lc.pos.idx < n_items) {
          if (Lc.j == 0) load_q(Lc);                 // (after PV(g) was issued: see load_q)
          mbar_wait(pv_done, (uint32_t)(g & 1));     // V buffer g&1 is free once PV(g) retired
          load_v(Lc, g + 2);
          cursor_next(Lc, walk);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax: warp = 16 rows x 128 keys
    const int quarter = warp & 3;                    // TMEM lane quarter (hardware: warp id % 4)
    const int hf = warp >> 2;                        // which 16 rows of the quarter
    const int row_lo = quarter * 32 + hf * 16 + (lane >> 2);   // this thread's query rows in the tile:
    const int row_hi = row_lo + 8;                              // row_lo and row_lo + 8
    const int cpair = 2 * (lane & 3);                // its columns: 8k + cpair, 8k + cpair + 1
    const uint32_t lane_off = (uint32_t)(quarter * 32 + hf * 16) << 16;
    const uint32_t tS_mine = tS + lane_off;
    const uint32_t tO_mine = tO + lane_off;
    const uint32_t tP_mine = tmem_base + A6_P_COL + lane_off;
    int g = 0;                                       // position in the CTA's tile stream
    int seq = 0;                                     // non-empty items so far

    Pos pos = pos_first(blockIdx.x, n_qt, heads);
    for (; pos.idx < n_items; pos_advance(pos, step, stride, n_qt, heads)) {
      const Item it = make_item(pos, R, kv_len);
      if (it.n_tiles == 0) {                         // no valid key at all: zeros
        const int r = threadIdx.x >> 1, h2 = threadIdx.x & 1;   // 256 threads: row x 64-byte half
        const int row = it.q0 + r;
        __nv_bfloat16* out = ctx + ((long long)(it.row_base + row)) * D + it.head * A6_DH + h2 * 32;
        if (row < R) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      float m_lo = 0.f, m_hi = 0.f;                  // running max (log2 units) used in exponents
      float l_lo = 0.f, l_hi = 0.f;                  // this thread's part of the two row sums

      for (int j = 0; j < it.n_tiles; ++j, ++g) {
        A6_T(0);
        mbar_wait(s_full, (uint32_t)(g & 1));
        A6_T(1);
        tc_fence_after();
        float s[64];                                 // s[4k+e]: row (e>>1 ? hi : lo), column 8k+cpair+(e&1)
        tmem_ld_16x256b_x16(tS_mine, s);
        tc_wait_ld();
        tc_fence_before();                           // S(g) is in registers: release the S columns
        __syncwarp();                                // for QK^T of tile g+1 right away
        if (lane == 0) mbar_arrive(s_free);
        if (j == 0 && seq > 0 && threadIdx.x == 0) {   // previous item's output store has left its staging
          bulk_wait_group_read0();                     // tile (= that item's Q buffer): hand it back to
          mbar_arrive(stage_free);                     // the producer (it refills it for item seq+1)
        }
        A6_T(2);
        const int n_valid = it.klen - j * A6_BN;     // >= 1; >= 128 except for the last tile
        if (n_valid < A6_BN) {                       // masked keys -> -inf (exp2 gives exactly 0)
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if ((i >> 2) * 8 + cpair + (i & 1) >= n_valid) s[i] = -INFINITY;
        }
        // Row max of the tile (exact, from the registers).
        auto tile_max = [&](float& mx_lo, float& mx_hi) {
          float a0 = max3(s[0], s[1], s[4]), a1 = max3(s[5], s[8], s[9]);
          float b0 = max3(s[2], s[3], s[6]), b1 = max3(s[7], s[10], s[11]);
#pragma unroll
          for (int k = 3; k < 15; k += 2) {
            a0 = max3(a0, s[4 * k], s[4 * k + 1]);
            a1 = max3(a1, s[4 * k + 4], s[4 * k + 5]);
            b0 = max3(b0, s[4 * k + 2], s[4 * k + 3]);
            b1 = max3(b1, s[4 * k + 6], s[4 * k + 7]);
          }
          a0 = max3(a0, s[60], s[61]);
          b0 = max3(b0, s[62], s[63]);
          mx_lo = quad_max(fmaxf(a0, a1)) * scale_log2;
          mx_hi = quad_max(fmaxf(b0, b1)) * scale_log2;
        };
        // p = 2^(s*scale - m): one FFMA + one MUFU per element; P -> TMEM as packed bf16x2, the A
        // operand of P V (pk[2k+e]: row lo/hi, packed column 4k + t%4)
        float sl0 = 0.f, sl1 = 0.f, sh0 = 0.f, sh1 = 0.f;
        uint32_t pk[32];
        auto exp_group = [&](int k, float nm_lo, float nm_hi) {
          const float e0 = ex2a(fmaf(s[4 * k + 0], scale_log2, nm_lo));
          const float e1 = ex2a(fmaf(s[4 * k + 1], scale_log2, nm_lo));
          const float e2 = ex2a(fmaf(s[4 * k + 2], scale_log2, nm_hi));
          const float e3 = ex2a(fmaf(s[4 * k + 3], scale_log2, nm_hi));
          sl0 += e0; sl1 += e1; sh0 += e2; sh1 += e3;
          pk[2 * k + 0] = pack_bf16x2(e0, e1);
          pk[2 * k + 1] = pack_bf16x2(e2, e3);
        };
        // With A6_OPT_GROUPS > 0, tiles after the first of an item start their first exponentials with the
        // running max before this tile's own row max is known (the max tree, the quad shuffles and the rescale
        // vote, ~500 cycles of dependent latency per tile, would then run on the ALU pipe while the XU pipe
        // already works); if the vote says the running max moved, those groups are recomputed.
        constexpr int A6_OPT = A6_OPT_GROUPS;
        float mx_lo, mx_hi;
        float f_lo = 1.f, f_hi = 1.f;
        bool need = false;
        bool pv_seen = (g == 0);
        A6_T(3);
        if (j == 0) {
          tile_max(mx_lo, mx_hi);
          m_lo = mx_lo;
          m_hi = mx_hi;
        } else {
          {
            const float nm_lo = -m_lo, nm_hi = -m_hi;
#pragma unroll
            for (int k = 0; k < A6_OPT; ++k) exp_group(k, nm_lo, nm_hi);
          }
          tile_max(mx_lo, mx_hi);
          if (mx_lo > m_lo + A6_RESCALE_THRESHOLD) { f_lo = ex2a(m_lo - mx_lo); m_lo = mx_lo; l_lo *= f_lo; need = true; }
          if (mx_hi > m_hi + A6_RESCALE_THRESHOLD) { f_hi = ex2a(m_hi - mx_hi); m_hi = mx_hi; l_hi *= f_hi; need = true; }
          // PV(g-1) must have retired before P is overwritten or O is rescaled. The wait sits right
          // before the P store (after the exponentials), except in the rare rescale branch.
          if (__any_sync(0xffffffffu, need)) {
            if (!pv_seen) {
              mbar_wait(pv_done, (uint32_t)((g - 1) & 1));
              tc_fence_after();
              pv_seen = true;
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {              // 16 columns at a time: keeps the scores in registers
              uint32_t o[8];
              tmem_ld_16x256b_x2(tO_mine + (uint32_t)(c * 16), o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                o[i] = __float_as_uint(__uint_as_float(o[i]) * ((i & 2) ? f_hi : f_lo));
              tmem_st_16x256b_x2(tO_mine + (uint32_t)(c * 16), o);
            }
            tc_wait_st();
            // redo the optimistic groups with the new running max (rare)
            sl0 = sl1 = sh0 = sh1 = 0.f;
            const float nm_lo = -m_lo, nm_hi = -m_hi;
#pragma unroll
            for (int k = 0; k < A6_OPT; ++k) exp_group(k, nm_lo, nm_hi);
          }
        }
        A6_T(4);
        {
          const float nm_lo = -m_lo, nm_hi = -m_hi;
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (j == 0 || k >= A6_OPT) exp_group(k, nm_lo, nm_hi);
          if (!pv_seen) {
            mbar_wait(pv_done, (uint32_t)((g - 1) & 1));
            tc_fence_after();
          }
          tmem_st_16x128b_x16(tP_mine, pk);
        }
        l_lo += sl0 + sl1;
        l_hi += sh0 + sh1;
        A6_T(5);
        tc_wait_st();
        tc_fence_before();                           // P (and a rescaled O) ordered before the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        A6_T(6);
      }

      // ---- item epilogue: O / l -> bf16 -> ctx
      mbar_wait(pv_done, (uint32_t)((g - 1) & 1));
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_16x256b_x8(tO_mine, o);
      const float inv_lo = __fdividef(1.f, quad_sum(l_lo));   // one MUFU.RCP; the result is rounded to bf16
      const float inv_hi = __fdividef(1.f, quad_sum(l_hi));
      tc_wait_ld();
      tc_fence_before();                             // O is in registers: PV of the next item may overwrite it
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      // The tile leaves through shared memory (128-byte-swizzled rows) and ONE TMA store per item:
      // row-strided stores straight from the registers cost ~2000 cycles per item here. The staging
      // tile is this item's own Q buffer (every QK^T of the item has retired); the 3-D map clips
      // rows >= R of the last query tile.
      uint8_t* stage = sQ + (seq & 1) * A6_TILE_BYTES;
      const uint32_t st_lo = smem_u32(stage) + (uint32_t)(row_lo * 128 + cpair * 2);
      const uint32_t st_hi = smem_u32(stage) + (uint32_t)(row_hi * 128 + cpair * 2);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t v0 = pack_bf16x2(__uint_as_float(o[4 * k + 0]) * inv_lo, __uint_as_float(o[4 * k + 1]) * inv_lo);
        const uint32_t v1 = pack_bf16x2(__uint_as_float(o[4 * k + 2]) * inv_hi, __uint_as_float(o[4 * k + 3]) * inv_hi);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(st_lo + (uint32_t)((k ^ (row_lo & 7)) << 4)), "r"(v0) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(st_hi + (uint32_t)((k ^ (row_hi & 7)) << 4)), "r"(v1) : "memory");
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        tma_store_3d(&tmap_out, stage, it.head * A6_DH, it.q0, it.b);
        bulk_commit_group();
      }
      ++seq;
    }
    if (threadIdx.x == 0) bulk_wait_group_read0();   // shared memory must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A6_CTRL_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A6_TMEM_COLS);
  }
#ifdef A6_TRACE
  if ((blockIdx.x == 5 || blockIdx.x == 100 || blockIdx.x == 153 || blockIdx.x == 290) && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("cta %d on sm %u: %lld cycles\n", (int)blockIdx.x, smid, clock64() - t_cta0);
    if (blockIdx.x == 5) printf("  item 2 epilogue: last arrive %lld pv_done %lld o_free %lld pair %lld staged %lld bar256 %lld stored %lld\n", g_ep[0], g_ep[1], g_ep[2], g_ep[3], g_ep[4], g_ep[5], g_ep[6]);
    const int w = blockIdx.x == 5 ? 0 : blockIdx.x == 100 ? 1 : blockIdx.x == 153 ? 2 : 3;
    for (int t = 0; t < 56; t += 8)
      printf("  cta %d tiles %d..: %lld %lld %lld %lld %lld %lld %lld %lld\n", (int)blockIdx.x, t, g_tile_t[w][t],
             g_tile_t[w][t + 1], g_tile_t[w][t + 2], g_tile_t[w][t + 3], g_tile_t[w][t + 4], g_tile_t[w][t + 5],
             g_tile_t[w][t + 6], g_tile_t[w][t + 7]);
  }
  if (blockIdx.x == A6_TRACE_CTA && threadIdx.x == 0)
    for (int t = 0; t < 8; ++t)
      printf("tile %d: wait_s %lld ld %lld max %lld pv_wait %lld exp %lld st_wait+arrive %lld | period %lld\n", 20 + t,
             g_trace[t][1] - g_trace[t][0], g_trace[t][2] - g_trace[t][1], g_trace[t][3] - g_trace[t][2],
             g_trace[t][4] - g_trace[t][3], g_trace[t][5] - g_trace[t][4], g_trace[t][6] - g_trace[t][5],
             t ? g_trace[t][0] - g_trace[t - 1][0] : 0ll);
#endif
}

}  // namespace

int attention_tc64_launch(const __nv_bfloat16* qkv, int B, int R, int heads, const int32_t* kv_len,
                          float scale, __nv_bfloat16* ctx, cudaStream_t s) {
  if (B <= 0 || R <= 0) return 0;
  const int D = heads * A6_DH;
  CUtensorMap tm;
  W2V_TRY(make_tmap_2d_bf16(&tm, qkv, (uint64_t)3 * D, (uint64_t)B * R, (uint64_t)3 * D, 64, A6_BN));
  CUtensorMap tm_out;   // [B][R][D]: a box over the end of a window is clipped, not spilled into the next
  W2V_TRY(make_tmap_3d_bf16(&tm_out, ctx, (uint64_t)D, (uint64_t)R, (uint64_t)B, (uint64_t)D, (uint64_t)R * D, 64, A6_BM));
  const float scale_log2 = scale * 1.4426950408889634f;
  W2V_ONCE_BEGIN
  W2V_CHECK_CUDA(cudaFuncSetAttribute(attention_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      A6_SMEM_BYTES));
  W2V_ONCE_END
  const int n_qt = (R + A6_BM - 1) / A6_BM;
  const long long n_items = (long long)n_qt * heads * B;
  W2V_REQUIRE(n_items < (1ll << 30), "attention: too many work items");
  W2V_REQUIRE(B <= A6_MAX_B, "attention: at most %d windows per launch (got %d)", A6_MAX_B, B);
  const long long slots = 2ll * num_sms();           // persistent: two CTAs per SM
  const int grid = (int)(n_items < slots ? n_items : slots);
  Step step;
  step.qt = grid % n_qt;
  step.head = (grid / n_qt) % heads;
  step.b = (grid / n_qt) / heads;
  ProfScope ps(s, "attention_d64");
  attention_tc64_kernel<<<grid, A6_THREADS, A6_SMEM_BYTES, s>>>(tm, tm_out, R, heads, n_qt, (int)n_items, B, step, kv_len,
                                                                scale_log2, ctx);
  W2V_CHECK_LAUNCH();
  return 0;
}

}  // namespace w2v
