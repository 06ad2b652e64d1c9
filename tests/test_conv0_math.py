"""CPU: the algebra of the tcgen05 conv-layer-0 kernel (LayerNorm folded into a K=12 fp16 dot product,
analytic per-frame variance from a Cholesky factor; csrc/conv0_tc.cu) restated in numpy
(oracle/conv0_folded.py) against torch Conv1d -> LayerNorm in fp64 (HF:281-299)."""
import math

import numpy as np
import pytest
import torch

from oracle import conv0_folded


def _reference(x, w, b, gamma, beta):
    y = torch.nn.functional.conv1d(torch.from_numpy(x).double()[None, None], torch.from_numpy(w).double()[:, None, :],
                                   torch.from_numpy(b).double(), stride=5)[0].T          # [T, 512]
    return torch.nn.functional.layer_norm(y, (512,), torch.from_numpy(gamma).double(),
                                          torch.from_numpy(beta).double(), 1e-5).numpy()


@pytest.mark.parametrize("case", ["random", "const_bias_dead_tap", "quiet_input", "zero_input"])
def test_folded_layernorm_matches_conv_then_layernorm(case):
    rng = np.random.default_rng(7)
    w = (rng.standard_normal((512, 10)) * 1.4 / math.sqrt(10)).astype(np.float32)
    b = (rng.standard_normal(512) * 0.05).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(512)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(512)).astype(np.float32)
    x = rng.standard_normal(16000).astype(np.float32)
    if case == "const_bias_dead_tap":        # Gram matrix only semi-definite: two zero pivots
        b[:] = 0.25
        w[:, 3] = 0.0
    if case == "quiet_input":                # frames whose variance is dominated by the bias spread / eps
        x *= 1e-3
    if case == "zero_input":                 # samples beyond a window's length read as 0
        x[:] = 0.0
    rows, U = conv0_folded.pack(w, b, gamma, beta)
    got = conv0_folded.forward(x, rows, U)
    ref = _reference(x, w, b, gamma, beta)
    assert got.shape == ref.shape
    err = np.abs(got - ref)
    # fp16 operands: 2^-11 per factor over a 12-term dot product; far below the bf16 rounding (2^-9
    # relative) the kernel applies to its output
    assert err.max() < 6e-3 and err.mean() < 6e-4, (err.max(), err.mean())


def test_variance_is_a_sum_of_squares():
    """|U z|^2 equals the biased channel variance of the conv output (what LayerNorm divides by)"""
    rng = np.random.default_rng(3)
    w = rng.standard_normal((512, 10)).astype(np.float32)
    b = rng.standard_normal(512).astype(np.float32)
    _, U = conv0_folded.pack(w, b, np.ones(512, np.float32), np.zeros(512, np.float32))
    x = rng.standard_normal(10).astype(np.float64)
    y = w.astype(np.float64) @ x + b
    z = np.concatenate([x, [1.0]])
    assert abs(np.sum((U.astype(np.float64) @ z) ** 2) - y.var()) < 1e-5 * y.var()


def test_fp16_range_guard_on_a_dead_tap():
    """a huge sample along a direction the taps ignore must not turn the frame into NaNs (inf * 0)"""
    rng = np.random.default_rng(11)
    w = (rng.standard_normal((512, 10)) * 0.4).astype(np.float32)
    w[:, 3] = 0.0
    b = np.full(512, 0.1, np.float32)
    gamma = np.ones(512, np.float32)
    beta = np.zeros(512, np.float32)
    x = np.zeros(400, np.float32)
    x[3::5] = 1000.0                     # only tap 3 (and 8) of each frame see it; tap 8 is live
    x[8::5] = 0.0
    x[3] = 5e4
    rows, U = conv0_folded.pack(w, b, gamma, beta)
    got = conv0_folded.forward(x, rows, U)
    assert np.isfinite(got).all()
