"""where the end-to-end step (TalkRunner.run on a 280 s talk in pinned host memory) spends its time"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402
from wav2vecsegmenter_b200.pipeline import TalkRunner  # noqa: E402

spec = synth.LARGE_ALL
eng = SFCEngine(spec)
eng.load_state_dict(synth.random_state_dict(spec, 0))
runner = TalkRunner(eng, batch_size=14, segment_sec=20, inference_times=1)
talk = (torch.randn(14 * 320000) * 0.1).pin_memory().numpy()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print(f"run() total            {timed(lambda: runner.run([talk])):.3f} ms")
wins, n_frames = runner.plan([talk])
print(f"plan() host            {timed(lambda: runner.plan([talk])):.3f} ms")
from wav2vecsegmenter_b200.pipeline import LazyWave  # noqa: E402

dev = LazyWave(talk, eng.device, runner._side())
dev.upload_to(len(talk))
torch.cuda.synchronize()
dev.waited = dev.hi   # resident: nothing left to wait for
print(f"H2D 17.9 MB            {timed(lambda: torch.from_numpy(talk).to(eng.device, non_blocking=True)):.3f} ms")
r_max = max(eng.frame_stride(max(w.n_samples, 400)) for w in wins)
print(f"_forward_rows          {timed(lambda: runner._forward_rows({0: dev}, wins, r_max)):.3f} ms")
rows = runner._forward_rows({0: dev}, wins, r_max)
print(f"reduce() (+D2H)        {timed(lambda: runner.reduce(rows, wins, n_frames)):.3f} ms")
audio = dev.dev.view(14, 320000)
lens = torch.full((14,), 320000, dtype=torch.int32, device="cuda")
ol = torch.full((14,), 999, dtype=torch.int32, device="cuda")
print(f"sfc_forward (resident) {timed(lambda: eng.sfc_forward(audio, lens, lens, ol, 320000)):.3f} ms")

# host-side (enqueue) cost of one step: no synchronisation inside the loop
def host_only(fn, n=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3


print(f"host enqueue: _forward_rows {host_only(lambda: runner._forward_rows({0: dev}, wins, r_max)):.3f} ms, "
      f"reduce_device {host_only(lambda: runner.reduce_device(rows, wins, n_frames)):.3f} ms, "
      f"sfc_forward {host_only(lambda: eng.sfc_forward(audio, lens, lens, ol, 320000)):.3f} ms")
t = timed(lambda: [None for _ in runner.run_stream((talk for _ in range(10)), depth=2)], n=3) / 10
print(f"run_stream per talk    {t:.3f} ms")
