"""prints max / mean abs probability error of the CUDA path vs every reference golden batch"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from util import load_gold, make_batch, spec_of  # noqa: E402
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402

import os

for name, corr in [(n, c) for n in ["tiny_batch", "middle_window", "middle_half_batch", "large_batch"] for c in ("0", "1")]:
    os.environ["W2VSEG_BIAS_CORRECTION"] = corr
    print(f"--- {name}, bias correction {'on' if corr == '1' else 'off'}")
    g = load_gold(name)
    spec = spec_of(g)
    lens = [int(x) for x in g["lens"]]
    eng = SFCEngine(spec)
    eng.load_state_dict(synth.random_state_dict(spec, int(g["seed"])))
    audio = make_batch(lens, int(g["audio_seed"])).cuda()
    lmax = max(lens)
    out_mask = g["out_mask"]
    logits, probs = eng.sfc_forward(audio, lens, [lmax] * len(lens), out_mask.sum(1).tolist(), lmax)
    lg = logits[:, : out_mask.shape[1]].cpu().numpy()
    le = (lg - g["logits"])[out_mask]
    print(f"   logits: ref std {g['logits'][out_mask].std():.3f}  err mean {le.mean():+.4f}  err std {le.std():.4f}  max |err| {np.abs(le).max():.4f}")
    p = probs[:, : out_mask.shape[1]].cpu().numpy()
    err = np.abs(p - g["probs"])[out_mask]
    hid, _ = eng.encode(audio, lens, [lmax] * len(lens), lmax)
    h = hid[:, : int(g["hidden_T"])][:, g["hidden_frames"]].cpu().numpy()
    rel = np.abs(h - g["hidden"]).max() / np.abs(g["hidden"]).max()
    rms = np.sqrt(((h - g["hidden"]) ** 2).mean() / (g["hidden"] ** 2).mean())
    print(f"{name:18s} layers {spec.keep_layers:2d}+{spec.adapter_layers:2d}ad  prob err max {err.max():.4f} mean {err.mean():.5f} "
          f"p99 {np.quantile(err, 0.99):.4f} | hidden rel-max {rel:.4f} rel-rms {rms:.4f} | same side of 0.5: "
          f"{((p > 0.5) == (g['probs'] > 0.5))[out_mask].mean():.4f}")
    eng.close()
    del eng
    torch.cuda.empty_cache()
