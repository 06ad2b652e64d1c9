"""helpers shared by the parity tests"""
from pathlib import Path

import numpy as np
import torch

from wav2vecsegmenter_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"


def load_gold(name):
    return np.load(GOLD / f"{name}.npz", allow_pickle=False)


def spec_of(g):
    k, a, hl, hh = [int(x) for x in g["spec"][:4]]
    extra = {}
    if len(g["spec"]) > 4:
        extra = {"feat_norm": "group" if int(g["spec"][4]) else "layer", "conv_bias": bool(int(g["spec"][5]))}
        if len(g["spec"]) > 6:
            extra["post_ln"] = bool(int(g["spec"][6]))
    return synth.ModelSpec(keep_layers=k, adapter_layers=a, head_layers=hl, head_heads=hh, **extra)


def make_batch(lens, audio_seed):
    """raw zero-padded batch exactly as CollateFn sees it before normalisation"""
    lmax = max(lens)
    audio = torch.zeros(len(lens), lmax)
    for i, n in enumerate(lens):
        audio[i, :n] = synth.synthetic_audio(int(n), int(audio_seed) + i)
    return audio


def out_lens_ref(lens):
    """frames per window as lib/dataset.py:665-666 computes end-start for a window starting at 0"""
    return [int(np.round((n + 1e-6) * 49.95 / 16000)) for n in lens]
