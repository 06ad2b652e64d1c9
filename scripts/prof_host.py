"""host-side cost of ONE forward (enqueue only: the CUDA work is queued asynchronously) — what a small-batch or
ragged-tail step is bound by. Prints us per forward for batch 14 and batch 1."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402

spec = synth.LARGE_ALL
eng = SFCEngine(spec)
eng.load_state_dict(synth.random_state_dict(spec, 0))
for B in (14, 1):
    L = 320000
    audio = torch.randn(B, L, device="cuda") * 0.1
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    ol = torch.full((B,), 999, dtype=torch.int32, device="cuda")
    R = eng.frame_stride(L)
    lg, pr = torch.empty(B, R, device="cuda"), torch.empty(B, R, device="cuda")
    for _ in range(3):
        eng.sfc_forward(audio, lens, lens, ol, L, lg, pr)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        eng.sfc_forward(audio, lens, lens, ol, L, lg, pr)
    t_enq = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    print(f"B={B}: host enqueue {t_enq * 1e6:.0f} us per forward, {t_all * 1e3:.2f} ms per forward incl. GPU")
for B in (1, 2):
    g = eng.graphed(B, 320000)
    g.audio.normal_(0, 0.1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    print(f"B={B}: CUDA-graph replay {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms per forward")
