"""ncu target: the three dominant GEMM shapes of the large (24/24) model at batch 14, a few
launches each (run plain first, then under ncu — see profiles/README)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
M = 14000
shapes = [("ffn_up", 4608, 1024, 1, 0), ("ffn_down", 1024, 4608, 0, 1), ("qkv", 3072, 1024, 0, 0)]
if os.environ.get("MORE"):
    shapes += [("attn_out", 1024, 1024, 0, 1), ("attn_out_inplace", 1024, 1024, 0, 2), ("ffn_down_inplace", 1024, 4608, 0, 2),
               ("attn_out_bf16out", 1024, 1024, 0, 0), ("conv_k3_l1", 512, 1536, 0, 0)]
g = torch.Generator(device="cuda").manual_seed(0)
BN = int(os.environ.get("BLOCK_N", "256"))
for name, N, K, act, f32 in shapes:
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / 32).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if f32 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    if f32 == 2:  # the engine's form: h += A W^T + b (out aliases resid)
        out, f32 = resid, 1
    NW, NT = (1, 1) if os.environ.get('NCU') else (3, 10)
    for _ in range(NW):
        n.check(lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, n.ptr(bias), act, n.ptr(resid), n.ptr(out),
                                f32, BN, n.current_stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(NT):
        lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, n.ptr(bias), act, n.ptr(resid), n.ptr(out), f32, BN,
                        n.current_stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / NT
    print(f"{name}: M={M} N={N} K={K} {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
