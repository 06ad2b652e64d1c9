"""Per-kernel parity on the GPU, through the C ABI, against plain PyTorch fp32 of the same op.

Tolerances: operands are bf16 (rel. 2^-9), accumulation fp32. Outputs that are stored as bf16 are
compared with atol/rtol a few bf16 ulps; fp32 outputs with the error a bf16-input GEMM must have.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from wav2vecsegmenter_b200 import _native

    lib = _native.load()
    _native.check(lib.w2vseg_device_ok(), "device_ok")
    return lib


def _gemm(lib, A, W, bias, act, resid, out_f32, block_n):
    from wav2vecsegmenter_b200 import _native as n

    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    n.check(
        lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, n.ptr(bias), act, n.ptr(resid), n.ptr(out),
                        int(out_f32), block_n, n.current_stream_ptr()),
        "gemm",
    )
    torch.cuda.synchronize()
    return out


def _ref_act(x, act):
    if act == 1:
        return torch.nn.functional.gelu(x)
    if act == 2:
        return torch.relu(x)
    return x


@pytest.mark.parametrize("block_n", [512, 256, 128, 64])
@pytest.mark.parametrize(
    "M,N,K,act,out_f32,use_resid",
    [
        (128, 256, 64, 0, 1, False),      # one tile, one k-block
        (128, 256, 256, 0, 1, False),     # one tile, several k-blocks (descriptor K advance)
        (300, 512, 1024, 0, 1, False),    # M tail (TMA zero fill + row predicate)
        (1000, 1024, 1024, 0, 1, True),   # residual epilogue
        (999, 3072, 1024, 0, 0, False),   # bf16 out
        (2000, 4096, 1024, 1, 0, False),  # GELU
        (777, 1024, 4096, 2, 0, False),   # ReLU, long K (pipeline wrap-around)
        (14000, 1024, 512, 0, 1, True),   # many tiles: persistent loop + TMEM double buffering
    ],
)
def test_gemm(lib, M, N, K, act, out_f32, use_resid, block_n):
    if block_n == 64 and not out_f32:
        pytest.skip("bf16 output is only built for block_n >= 128")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + act)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if use_resid else None
    out = _gemm(lib, A, W, bias, act, resid, out_f32, block_n)
    ref = _ref_act(A.float() @ W.float().t() + bias, act)
    if use_resid:
        ref = ref + resid
    err = (out.float() - ref).abs().max().item()
    tol = 2e-3 if out_f32 else 3e-2
    assert err < tol, f"max abs err {err}"


def test_gemm_gelu_large_arguments(lib):
    """GELU in the GEMM epilogue over the whole fp32-relevant range: |x| up to ~60 (outlier FFN
    units of trained checkpoints); gelu(x) -> x for large x, -> 0 for large negative x."""
    M, N, K = 256, 256, 64
    A = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    A[:, 0] = 1.0
    W = torch.zeros(N, K, device="cuda", dtype=torch.bfloat16)
    bias = torch.linspace(-60.0, 60.0, N, device="cuda")     # pre-activation == bias
    out = _gemm(lib, A, W, bias, 1, None, 1, 256)
    ref = torch.nn.functional.gelu(bias)[None].expand(M, N)
    err = (out - ref).abs()
    assert (err <= 3e-5 + 6e-4 * ref.abs()).all(), f"max abs err {err.max().item()}"


def test_gemm_inplace_residual(lib):
    from wav2vecsegmenter_b200 import _native as n

    M, N, K = 1500, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / 32).bfloat16()
    h = torch.randn(M, N, device="cuda", generator=g)
    ref = h + A.float() @ W.float().t()
    n.check(lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, None, 0, n.ptr(h), n.ptr(h), 1, 256,
                            n.current_stream_ptr()))
    torch.cuda.synchronize()
    assert (h - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_attention_random_ragged_batches(lib, seed):
    """persistent d=64 kernel: every CTA walks ~3-6 work items; random key lengths (empty, one-tile,
    partial and full windows in any order) against the independent mma.sync implementation"""
    from wav2vecsegmenter_b200 import _native as n

    rng = torch.Generator().manual_seed(seed)
    B, heads, dh = 14, 16, 64
    R = int(torch.randint(300, 900, (1,), generator=rng))
    lens = torch.randint(0, R + 1, (B,), generator=rng)
    lens[torch.rand(B, generator=rng) < 0.3] = int(torch.randint(1, 129, (1,), generator=rng))   # one-tile windows
    lens[torch.rand(B, generator=rng) < 0.1] = 0
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(B * R, 3 * D, device="cuda", generator=g).bfloat16()
    kv_len = lens.to(device="cuda", dtype=torch.int32)
    outs = []
    for impl in ("w2vseg_attention", "w2vseg_attention_mma"):
        ctx = torch.full((B * R, D), float("nan"), device="cuda", dtype=torch.bfloat16)
        n.check(getattr(lib, impl)(n.ptr(qkv), B, R, heads, dh, n.ptr(kv_len), dh ** -0.5, n.ptr(ctx),
                                   n.current_stream_ptr()))
        torch.cuda.synchronize()
        outs.append(ctx.float())
    assert torch.isfinite(outs[0]).all()
    assert (outs[0] - outs[1]).abs().max().item() < 3e-2


@pytest.mark.parametrize("kw,rows_out", [(3, 1999), (2, 999), (3, 31999)])
def test_conv_as_implicit_gemm(lib, kw, rows_out):
    """stride-2 Conv1d over channels-last activations == GEMM over an overlapping-row TMA view"""
    from wav2vecsegmenter_b200 import _native as n

    C = 512
    rows_in = 2 * rows_out + kw  # slack rows included
    g = torch.Generator(device="cuda").manual_seed(kw * 100 + rows_out)
    x = (torch.randn(rows_in, C, device="cuda", generator=g)).bfloat16()
    w = (torch.randn(C, C, kw, device="cuda", generator=g) / math.sqrt(C * kw)).bfloat16()  # [O, I, J]
    bias = torch.randn(C, device="cuda", generator=g)
    wp = w.permute(0, 2, 1).contiguous().view(C, kw * C)  # K index = j*C + i
    out = torch.empty(rows_out, C, device="cuda", dtype=torch.bfloat16)
    n.check(lib.w2vseg_conv_gemm(n.ptr(x), rows_out, C, kw, 2, n.ptr(wp), C, n.ptr(bias), n.ptr(out),
                                 n.current_stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv1d(x.float().t()[None], w.float(), bias, stride=2)[0].t()[:rows_out]
    assert (out.float() - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("B,R,D,taps", [(2, 300, 1024, 128), (3, 129, 128, 16), (1, 1000, 256, 128), (4, 50, 128, 128)])
def test_posconv(lib, B, R, D, taps, impl):
    """grouped positional conv + GELU + residual (HF:326-379, 764-765): the resident-A kernel of the
    forward pass (impl 0) and the generic shifted-row GEMM (impl 1) against torch conv1d"""
    from wav2vecsegmenter_b200 import _native as n

    halo, G = taps // 2, D // 64
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + R + taps)
    x = torch.randn(B, R, D, device="cuda", generator=g).bfloat16()                 # conv input
    w = (torch.randn(D, 64, taps, device="cuda", generator=g) / math.sqrt(64 * taps)).bfloat16()  # [O, I/g, J]
    bias = torch.randn(D, device="cuda", generator=g) * 0.1
    h0 = torch.randn(B * R, D, device="cuda", generator=g)
    zpad = torch.zeros(B * (R + 2 * halo) + 2 * halo, D, device="cuda", dtype=torch.bfloat16)
    for b in range(B):
        zpad[b * (R + 2 * halo) + halo: b * (R + 2 * halo) + halo + R] = x[b]
    wp = w.permute(0, 2, 1).contiguous().view(D, taps * 64)                          # K index = j*64 + i
    h = h0.clone()
    n.check(lib.w2vseg_posconv(n.ptr(zpad), n.ptr(wp), n.ptr(bias), B, R, D, taps, n.ptr(h), impl,
                               n.current_stream_ptr()))
    torch.cuda.synchronize()
    y = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=halo, groups=G)
    y = y[:, :, :R].transpose(1, 2).reshape(B * R, D)                                # SamePad drops the last frame
    ref = h0 + torch.nn.functional.gelu(y)
    err = (h - ref).abs().max().item()
    assert err < 5e-3, f"max abs err {err}"


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize(
    "R0,lens,case",
    [
        (640, [3205, 3205], "plain"),                 # 5 frame blocks per window, all samples valid
        (320, [1605, 700, 0, 1605], "ragged"),        # 2.5 blocks per window (row predicate), short + empty windows
        (64000, [320000], "full"),                    # one 20 s window: persistent loop, TMEM double buffering
        (1280, [6405, 6000, 6405], "const_bias"),     # constant bias + a dead tap: semi-definite Gram matrix
    ],
)
def test_conv0(lib, R0, lens, case, impl):
    """conv layer 0 + LayerNorm(512) + GELU with on-the-fly window normalisation (HF:281-299): the
    tcgen05 kernel with the LayerNorm folded into the MMA (impl 0) and the CUDA-core kernel (impl 1)
    against torch conv1d -> layer_norm -> gelu in fp32."""
    from wav2vecsegmenter_b200 import _native as n

    B = len(lens)
    g = torch.Generator(device="cuda").manual_seed(R0 + B)
    stride = R0 * 5 + 8
    audio = torch.randn(B, stride, device="cuda", generator=g) * 0.3 + 0.05
    w = torch.randn(512, 10, device="cuda", generator=g) / math.sqrt(10)
    bias = torch.randn(512, device="cuda", generator=g) * 0.2
    gamma = torch.randn(512, device="cuda", generator=g)
    beta = torch.randn(512, device="cuda", generator=g) * 0.5
    if case == "const_bias":
        bias.fill_(0.25)
        w[:, 3] = 0.0
    slen = torch.tensor(lens, device="cuda", dtype=torch.int32)
    stats = torch.stack([torch.full((B,), 0.05, device="cuda"), torch.full((B,), 1 / 0.3, device="cuda")], dim=1).contiguous()
    out = torch.full((B * R0, 512), float("nan"), device="cuda", dtype=torch.bfloat16)
    scratch = torch.empty(65536, device="cuda", dtype=torch.uint8)
    n.check(lib.w2vseg_conv0(n.ptr(audio), stride, n.ptr(slen), n.ptr(stats), n.ptr(w), n.ptr(bias), n.ptr(gamma),
                             n.ptr(beta), 1e-5, n.ptr(out), B, R0, impl, n.ptr(scratch), scratch.numel(),
                             n.current_stream_ptr()), "conv0")
    torch.cuda.synchronize()
    idx = torch.arange(R0 * 5 + 5, device="cuda")
    xn = torch.where(idx[None, :] < slen[:, None], (audio[:, : R0 * 5 + 5] - 0.05) * (1 / 0.3), torch.zeros((), device="cuda"))
    y = torch.nn.functional.conv1d(xn[:, None, :].double(), w[:, None, :].double(), bias.double(), stride=5)   # [B, 512, R0]
    assert y.shape[2] == R0
    y = torch.nn.functional.layer_norm(y.transpose(1, 2), (512,), gamma.double(), beta.double(), 1e-5)
    ref = torch.nn.functional.gelu(y).float().reshape(B * R0, 512)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs()
    # bf16 output (2^-9 relative) + fp16 operands of the folded MMA (2^-11 per factor) + MUFU.TANH
    assert (err <= 1e-2 + 8e-3 * ref.abs()).all(), f"max abs err {err.max().item()}"
    assert err.mean().item() < 2e-3


@pytest.mark.parametrize("C,in_f32,act,rows", [(1024, True, 0, 4099), (512, False, 0, 4099), (512, False, 1, 4099),
                                               (512, True, 0, 4099), (512, False, 1, 23017), (512, False, 0, 7),
                                               (1024, True, 0, 23017)])
def test_layernorm(lib, C, in_f32, act, rows):
    from wav2vecsegmenter_b200 import _native as n

    g = torch.Generator(device="cuda").manual_seed(C + act)
    x = torch.randn(rows, C, device="cuda", generator=g) * 3 + 1
    if not in_f32:
        x = x.bfloat16()
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    out = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
    n.check(lib.w2vseg_layernorm(n.ptr(x), int(in_f32), rows, C, n.ptr(gamma), n.ptr(beta), 1e-5, act,
                                 n.ptr(out), n.current_stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    if act:
        ref = torch.nn.functional.gelu(ref)
    # bf16 output: half an ulp is 2^-9 relative; GELU adds <= 5e-4 relative (MUFU.TANH)
    err = (out.float() - ref).abs()
    assert (err <= 1e-2 + 6e-3 * ref.abs()).all(), f"max abs err {err.max().item()}"
    assert err.mean().item() < 3e-3


@pytest.mark.parametrize("impl", ["w2vseg_attention", "w2vseg_attention_mma"])
@pytest.mark.parametrize("heads,dh", [(16, 64), (8, 128)])
@pytest.mark.parametrize("R,lens", [(1000, [999, 999]), (333, [333, 1, 200]), (130, [64, 65, 0, 130]),
                                    (1100, [1099, 128, 129, 257]),
                                    # many items per persistent CTA, one-tile / empty items between long ones
                                    # (ragged talk tails: a one-tile item between two others once deadlocked)
                                    (1000, [999, 100, 999, 64, 128, 999, 1, 999, 129, 999, 50, 0, 999, 120]),
                                    (640, [100] * 9 + [640, 100, 100, 300, 100] * 3)])
def test_attention(lib, heads, dh, R, lens, impl):
    from wav2vecsegmenter_b200 import _native as n

    B = len(lens)
    D = heads * dh
    g = torch.Generator(device="cuda").manual_seed(R + dh)
    qkv = torch.randn(B * R, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 3.0  # sharper softmax: running-max updates (and the lazy O rescale) get exercised
    qkv = qkv.bfloat16()
    kv_len = torch.tensor(lens, device="cuda", dtype=torch.int32)
    ctx = torch.empty(B * R, D, device="cuda", dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(dh)
    n.check(getattr(lib, impl)(n.ptr(qkv), B, R, heads, dh, n.ptr(kv_len), scale, n.ptr(ctx),
                               n.current_stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, R, 3, heads, dh).permute(2, 0, 3, 1, 4)  # [B, H, R, dh]
    s = (q @ k.transpose(-1, -2)) * scale
    mask = torch.arange(R, device="cuda")[None, :] >= kv_len[:, None]  # [B, R] keys
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)  # windows with zero valid keys -> zeros
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * R, D)
    assert (ctx.float() - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("impl", ["w2vseg_attention", "w2vseg_attention_mma"])
@pytest.mark.parametrize("heads,dh,R,lens", [(16, 64, 1000, [999, 640]), (8, 128, 700, [700, 333]), (16, 64, 300, [300, 129, 1])])
def test_attention_sharply_peaked_scores(lib, heads, dh, R, lens, impl):
    """scores spread over +-35 (q and k scaled by 3: softmax close to an arg-max, the running max moves by more than
    the lazy-rescale threshold 2^8 many times per row, most probabilities underflow to 0 in bf16) against fp32 torch on
    the SAME bf16 inputs. (End to end such a regime cannot be compared with an fp32 reference: a 0.4 % bf16 rounding
    of q / k moves a score of 60 by 0.24, i.e. the attention weights by 27 % — for any bf16 implementation.)"""
    from wav2vecsegmenter_b200 import _native as n

    B, D = len(lens), heads * dh
    g = torch.Generator(device="cuda").manual_seed(7 * R + dh)
    qkv = torch.randn(B * R, 3 * D, device="cuda", generator=g)
    qkv[:, : 2 * D] *= 3.0
    qkv = qkv.bfloat16()
    kv_len = torch.tensor(lens, device="cuda", dtype=torch.int32)
    ctx = torch.empty(B * R, D, device="cuda", dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(dh)
    n.check(getattr(lib, impl)(n.ptr(qkv), B, R, heads, dh, n.ptr(kv_len), scale, n.ptr(ctx), n.current_stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, R, 3, heads, dh).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    assert s.abs().max().item() > 30
    mask = torch.arange(R, device="cuda")[None, :] >= kv_len[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * R, D)
    assert torch.isfinite(ctx.float()).all()
    assert (ctx.float() - ref).abs().max().item() < 4e-2
