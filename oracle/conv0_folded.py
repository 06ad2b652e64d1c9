"""TEST INFRASTRUCTURE ONLY — numpy restatement of the algebra behind the tcgen05 conv-layer-0 kernel
(wav2vecsegmenter_b200/csrc/conv0_tc.cu), so that the derivation is pinned on the CPU against the plain
`Conv1d(1, 512, k=10, stride=5) -> LayerNorm(512) -> GELU` of the reference (HF:281-299).

With ONE input channel the LayerNorm statistics of a frame are a closed form of its 10 samples:
  z = [x_0..x_9, 1],  wt[c] = [w[c,:] - mean_c w, b_c - mean_c b]      (channel-centred taps + bias)
  y_c - mean_c(y) = wt[c] . z            var_c(y) = z^T G z,  G = wt^T wt / 512 = U^T U
and the LayerNorm output is a K = 12 dot product of
  A row (frame)   = [rstd*x_0 .. rstd*x_9, rstd, 1]          (rounded to fp16 by the kernel)
  W row (channel) = [gamma_c*wt[c,0..10], beta_c]            (rounded to fp16 by the kernel)
"""
from __future__ import annotations

import numpy as np


def pack(w: np.ndarray, bias: np.ndarray, gamma: np.ndarray, beta: np.ndarray):
    """w [512, 10], bias/gamma/beta [512] -> (W rows fp16 [512, 12], U upper-triangular fp32 [11, 11]);
    mirrors conv0_pack_kernel: fp64 centring and Gram matrix, Cholesky with zero pivots -> zero rows"""
    v = np.concatenate([w.astype(np.float64), bias.astype(np.float64)[:, None]], axis=1)      # [512, 11]
    v = v - v.mean(axis=0, keepdims=True)
    G = v.T @ v / v.shape[0]
    U = np.zeros((11, 11))
    dmax = float(np.max(np.diag(G)))
    for i in range(11):
        d = G[i, i] - np.sum(U[:i, i] ** 2)
        if d > 1e-13 * dmax and d > 0.0:
            U[i, i] = np.sqrt(d)
            for j in range(i + 1, 11):
                U[i, j] = (G[i, j] - np.sum(U[:i, i] * U[:i, j])) / U[i, i]
    ga = gamma.astype(np.float32)[:, None]
    rows = np.concatenate([ga * v.astype(np.float32), beta.astype(np.float32)[:, None]], axis=1)
    return rows.astype(np.float16), U.astype(np.float32)


def forward(x: np.ndarray, rows_f16: np.ndarray, U: np.ndarray, eps: float = 1e-5):
    """x fp32 [n_samples] (already normalised) -> LayerNorm output BEFORE the GELU, fp32 [T, 512], computed
    the way the kernel does: fp32 variance from |U z|^2, fp16 operands, fp32 accumulation."""
    T = (len(x) - 10) // 5 + 1
    idx = 5 * np.arange(T)[:, None] + np.arange(10)[None, :]
    frames = x.astype(np.float32)[idx]                                        # [T, 10]
    z = np.concatenate([frames, np.ones((T, 1), np.float32)], axis=1)        # [T, 11]
    s = z @ U.T.astype(np.float32)                                            # rows of U z
    var = np.sum(s * s, axis=1, dtype=np.float32)
    rstd = (1.0 / np.sqrt(var + np.float32(eps))).astype(np.float32)
    a = np.concatenate([np.clip(frames * rstd[:, None], -60000.0, 60000.0), rstd[:, None],
                        np.ones((T, 1), np.float32)], axis=1)   # clip = the kernel's fp16 range guard
    a16 = a.astype(np.float16)
    return a16.astype(np.float32) @ rows_f16.astype(np.float32).T            # fp32 accumulate
