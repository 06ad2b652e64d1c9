"""micro-benchmark / ncu target: conv layer 0 + LayerNorm + GELU at the bench shape (14 x 20 s windows).
Prints the time of both implementations and the achieved HBM write bandwidth (the layer's floor is
writing its 917 MB bf16 output)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
B, R0 = 14, 64000
stride = 320000
audio = torch.randn(B, stride, device="cuda")
slen = torch.full((B,), stride, device="cuda", dtype=torch.int32)
stats = torch.tensor([[0.0, 1.0]] * B, device="cuda")
w = torch.randn(512, 10, device="cuda") / math.sqrt(10)
bias = torch.randn(512, device="cuda") * 0.1
gamma = torch.ones(512, device="cuda")
beta = torch.zeros(512, device="cuda")
out = torch.empty(B * R0 + 4, 512, device="cuda", dtype=torch.bfloat16)
scratch = torch.empty(65536, device="cuda", dtype=torch.uint8)
only = sys.argv[1:] and int(sys.argv[1])
for impl, name in ((0, "tcgen05 (LN folded into the MMA)"), (1, "CUDA cores")):
    if sys.argv[1:] and impl != only:
        continue
    args = (n.ptr(audio), stride, n.ptr(slen), n.ptr(stats), n.ptr(w), n.ptr(bias), n.ptr(gamma), n.ptr(beta),
            1e-5, n.ptr(out), B, R0, impl, n.ptr(scratch), scratch.numel(), n.current_stream_ptr())
    for _ in range(3):
        n.check(lib.w2vseg_conv0(*args))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.w2vseg_conv0(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10   # includes the (tiny) weight-pack / transpose launch of the test entry
    byts = B * R0 * 512 * 2 + B * stride * 4
    print(f"conv0 {name}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s (algorithmic bytes {byts/1e6:.0f} MB)")
