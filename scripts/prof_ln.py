"""micro-benchmark: LayerNorm(1024) of the fp32 residual stream at the bench shape (14 000 rows).
W2VSEG_LN=v1 selects the one-row-per-warp kernel for A/B runs. The input rotates over 4 buffers
(4 x 57 MB > L2) so that every launch reads from HBM like in the forward pass."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
rows, C = 14000, 1024
xs = [torch.randn(rows, C, device="cuda") for _ in range(4)]
gamma = torch.randn(C, device="cuda")
beta = torch.randn(C, device="cuda")
out = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
st = n.current_stream_ptr()
for i in range(8):
    n.check(lib.w2vseg_layernorm(n.ptr(xs[i % 4]), 1, rows, C, n.ptr(gamma), n.ptr(beta), 1e-5, 0, n.ptr(out), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 40
for i in range(N):
    lib.w2vseg_layernorm(n.ptr(xs[i % 4]), 1, rows, C, n.ptr(gamma), n.ptr(beta), 1e-5, 0, n.ptr(out), st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
byts = rows * C * 6
print(f"layernorm1024 [{os.environ.get('W2VSEG_LN', 'stream')}]: {ms*1e3:.2f} us  {byts/ms/1e6:.0f} GB/s (4 B read + 2 B written per element)")
