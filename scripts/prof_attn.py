"""micro-benchmark / ncu target: attention at the large-model shape (B=14, R=1000, 16 x 64)"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
B, R = 14, 1000
for heads, dh, name in ((16, 64, "enc"), (8, 128, "head")):
    D = heads * dh
    qkv = torch.randn(B * R, 3 * D, device="cuda").bfloat16()
    kv = torch.full((B,), 999, dtype=torch.int32, device="cuda")
    ctx = torch.empty(B * R, D, device="cuda", dtype=torch.bfloat16)
    for impl in ("w2vseg_attention", "w2vseg_attention_mma"):
        fn = getattr(lib, impl)
        NW, NT = (1, 1) if os.environ.get("NCU") else (3, 20)
        for _ in range(NW):
            n.check(fn(n.ptr(qkv), B, R, heads, dh, n.ptr(kv), dh ** -0.5, n.ptr(ctx), n.current_stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(NT):
            fn(n.ptr(qkv), B, R, heads, dh, n.ptr(kv), dh ** -0.5, n.ptr(ctx), n.current_stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / NT
        print(f"{name} {impl}: {ms*1e3:.1f} us  {4*999*999*D*B/ms/1e9:.1f} TFLOP/s")
