"""times the head-only training step (frozen encoder) at the benchmark shape: large (24 layers, no adapters:
the reference's frozen 'large (0/24)' model), batch 14 x 20 s. Prints ms for the encoder forward and the head
forward+loss+backward, and the per-kernel breakdown of the latter."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as nat  # noqa: E402
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402
from wav2vecsegmenter_b200.train import HeadTrainer  # noqa: E402

spec = synth.ModelSpec(keep_layers=24, adapter_layers=0)
sd = synth.random_state_dict(spec, 0)
eng = SFCEngine(spec)
eng.load_state_dict(sd)
head = {k[len("seg_model."):]: v for k, v in sd.items() if k.startswith("seg_model.")}
tr = HeadTrainer(eng, head)
B, L = 14, 320_000
T = eng.num_frames(L)
audio = torch.randn(B, L, device="cuda") * 0.1
lens = [L] * B
target = (torch.rand(B, T, device="cuda") > 0.5).float()
hidden, _ = eng.encode(audio, lens, lens, L)
hid = hidden[:, :T].contiguous()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_enc = timed(lambda: eng.encode(audio, lens, lens, L))
t_head = timed(lambda: tr.step_hidden(hid, [T] * B, target, 0.7))
t_sync = timed(lambda: tr.sync())
print(f"encoder forward {t_enc:.2f} ms | head fwd+loss+bwd {t_head:.2f} ms | parameter re-upload {t_sync:.2f} ms "
      f"| step total {t_enc + t_head + t_sync:.2f} ms = {B * 20 / (t_enc + t_head + t_sync) * 1e3:.0f} audio-s/s trained")
eng.lib.w2vseg_profile_enable(1)
tr.step_hidden(hid, [T] * B, target, 0.7)
buf = nat.C.create_string_buffer(1 << 16)
eng.lib.w2vseg_profile_collect(buf, len(buf))
eng.lib.w2vseg_profile_enable(0)
for line in buf.value.decode().splitlines():
    print("   ", line)
