"""A/B of the step time (large 24/24 + adapters, 14 x 20 s) with PDL off / on, for the library named by W2VSEG_LIB.
Needs a library built with experiments/pdl_r02.patch.txt applied (it adds w2vseg_set_pdl); the shipped library does not
use programmatic dependent launch (profiles/experiments_r02.md: no gain under the power cap)."""
import os
import torch
from wav2vecsegmenter_b200 import synth
from wav2vecsegmenter_b200.engine import SFCEngine

spec = synth.LARGE_ALL
eng = SFCEngine(spec)
eng.load_state_dict(synth.random_state_dict(spec, 2))
lib = eng.lib
B, L = 14, 320000
audio = torch.stack([synth.synthetic_audio(L, 900 + i) for i in range(B)]).cuda()
sl = torch.full((B,), L, dtype=torch.int32, device="cuda")
ol = torch.full((B,), eng.num_frames(L), dtype=torch.int32, device="cuda")
ms = {}
for mode in (0, 1, 0, 1):
    lib.w2vseg_set_pdl(mode)
    for _ in range(3):
        eng.sfc_forward(audio, sl, sl, ol, L)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.sfc_forward(audio, sl, sl, ol, L)
    e1.record()
    torch.cuda.synchronize()
    ms.setdefault(mode, []).append(round(e0.elapsed_time(e1) / 20, 3))
print(os.environ.get("W2VSEG_LIB", "default").split("/")[-2], "PDL off", ms[0], "on", ms[1])
