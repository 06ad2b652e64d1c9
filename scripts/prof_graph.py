"""experiment: the steady-state forward replayed from a CUDA graph vs launched kernel by kernel
(large 24/24 + adapters, 14 x 20 s, inputs rotating over 4 batches). Same timing as bench.py."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402

spec = synth.LARGE_ALL
eng = SFCEngine(spec)
eng.load_state_dict(synth.random_state_dict(spec, 0))
B, L = 14, 320000
audio = [torch.randn(B, L, device="cuda") * 0.1 for _ in range(4)]
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
ol = torch.full((B,), 999, dtype=torch.int32, device="cuda")
R = eng.frame_stride(L)
logits = torch.empty(B, R, device="cuda")
probs = torch.empty(B, R, device="cuda")


def step(i):
    eng.sfc_forward(audio[i % 4], lens, lens, ol, L, logits, probs)


for i in range(4):
    step(i)
torch.cuda.synchronize()
graphs = []
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(4):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step(i)
        graphs.append(g)
torch.cuda.synchronize()


def timed(fn, n=20):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rep in range(2):
    a = timed(step)
    b = timed(lambda i: graphs[i % 4].replay())
    print(f"kernel-by-kernel {a:.3f} ms/step, CUDA graph {b:.3f} ms/step")
