"""micro-benchmark / ncu target: positional conv at the large-model shape (B=14, R=1000, D=1024, 128 taps)"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
B, R, D, taps = 14, 1000, 1024, 128
halo = taps // 2
zpad = torch.randn(B * (R + 2 * halo) + 2 * halo, D, device="cuda").bfloat16()
wp = (torch.randn(D, taps * 64, device="cuda") / math.sqrt(64 * taps)).bfloat16()
bias = torch.randn(D, device="cuda")
h = torch.randn(B * R, D, device="cuda")
for impl, name in ((0, "resident-A kernel"), (1, "generic shifted-row GEMM")):
    for _ in range(3):
        n.check(lib.w2vseg_posconv(n.ptr(zpad), n.ptr(wp), n.ptr(bias), B, R, D, taps, n.ptr(h), impl, n.current_stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.w2vseg_posconv(n.ptr(zpad), n.ptr(wp), n.ptr(bias), B, R, D, taps, n.ptr(h), impl, n.current_stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"posconv {name}: {ms*1e3:.1f} us  {2*B*R*D*64*taps/ms/1e9:.1f} TFLOP/s")
