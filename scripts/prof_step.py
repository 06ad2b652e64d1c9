"""ncu target: ONE steady-state SFC forward (large 24/24 + adapters, batch 14 x 20 s) bracketed by
cudaProfilerStart/Stop, after weight upload and two warm-up passes. Run plain first, then
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
"""
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
audio = torch.randn(B, L, device="cuda") * 0.1
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
ol = torch.full((B,), 999, dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.sfc_forward(audio, lens, lens, ol, L)
torch.cuda.synchronize()
torch.cuda.profiler.start()
_, p = eng.sfc_forward(audio, lens, lens, ol, L)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(p.mean()))
