// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA into 128B-swizzled shared memory), persistent over output tiles, with the
// epilogues the SFC forward pass needs fused in:
//
//   out_bf16[r, c] = act(acc + bias[c])                       (QKV, FFN-up, conv pre-norm, ...)
//   out_f32 [r, c] = resid[r, c] + act(acc + bias[c])         (out-proj, FFN-down, pos-conv)
//
// C[M,N] = A[M,K] * W[N,K]^T, both operands K-major. The A operand is described only by a TMA
// tensor map, so the same kernel runs
//   * plain Linear layers                       (A = activations [M,K]),
//   * the strided Conv1d layers as implicit GEMM (A = overlapping-row im2col VIEW of the
//     channels-last activations: row stride 2*512, row length k*512 — no im2col copy), and
//   * the 128-tap grouped positional conv       (a_mode 1: K-block j reads rows shifted by j).
//
// Reference ops replaced: torch.nn.Linear / Conv1d inside HF Wav2Vec2 (HF:281-292, 429-434,
// 360-368, 500-549, 566-573), lib/models.py:383-387 (adapter), torch TransformerEncoderLayer
// projections (lib/models.py:291-300).
#pragma once
#include "common.h"

namespace w2v {

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct GemmProblem {
  // operands
  const __nv_bfloat16* A;   // base of the A view
  int64_t a_rows;           // rows visible through the A view (TMA zero-fills beyond)
  int64_t a_row_stride;     // elements between consecutive A rows (may be < a_cols: im2col view)
  int64_t a_cols;           // contiguous extent of one A row (== K except a_mode 1)
  const __nv_bfloat16* W;   // [N, K] row-major (K contiguous)
  int N, K;
  // output row space: num_groups groups of rows_per_group rows
  int num_groups;           // 1 = flat
  int rows_per_group;
  int64_t a_group_rows;     // A-row offset between consecutive groups
  int64_t o_group_rows;     // output-row offset between consecutive groups
  int a_mode;               // 0: A tile = (row0, kb*64) ; 1: A tile = (row0 + kb, nb*BLOCK_N)
  // epilogue
  const float* bias;        // [N] or null
  int act_split;            // columns <  act_split use act_lo, columns >= act_split use act_hi
  int act_lo, act_hi;
  const float* resid;       // fp32 [*, ld_resid] or null (only with out_f32)
  int64_t ld_resid;
  void* out;
  int64_t ld_out;
  int out_f32;              // 1: fp32 output, 0: bf16 output
  const int* mask_len;      // optional, device: rows with (row % mask_period) >= mask_len[row /
  int mask_period;          //   mask_period] are written as zeros (frame mask, HF:753-756)
};

// Launches on `stream`. block_n in {64, 128, 256}; N % block_n == 0; K % 64 == 0.
int gemm_tc_launch(const GemmProblem& p, int block_n, cudaStream_t stream);

// CTA-pair (cta_group::2) 256x256-tile variant: a_mode 0, N % 256 == 0 (gemm_tc2.cu).
int gemm_tc2_launch(const GemmProblem& p, cudaStream_t stream);

// Positional conv: same problem description as gemm_tc_launch(p, 64, ..) with a_mode 1, but the
// 256-row input block of a tile stays resident in shared memory for all taps (posconv_tc.cu).
int posconv_tc_launch(const GemmProblem& p, cudaStream_t stream);

}  // namespace w2v
