"""TEST / ANALYSIS INFRASTRUCTURE (oracle side, never imported by the product path). CPU study: does per-layer BIAS CORRECTION (b' = b - (bf16(W) - W) @ mean_x, mean_x from a calibration
signal) remove the frame-independent logit offset that bf16 weight rounding causes — on a DIFFERENT signal
than the one calibrated on? fp32 arithmetic, weights rounded to bf16 (the CUDA path's configuration)."""
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent   # oracle/ -> repo root
sys.path.insert(0, str(ROOT))
from oracle import sfc_oracle as so  # noqa: E402
from wav2vecsegmenter_b200 import synth  # noqa: E402

torch.set_num_threads(8)
spec = synth.LARGE_ALL if len(sys.argv) < 2 else getattr(synth, sys.argv[1])
sd = synth.random_state_dict(spec, 0)
W2V = so.W2V


class Lin:
    """linear layers by name; mode 'fp32' | 'bf16' | 'bf16+corr'; records mean input when calibrating"""

    def __init__(self, sd):
        self.sd, self.mode, self.calib, self.mean_x, self.corr = sd, "fp32", False, {}, {}

    def __call__(self, name, x, w, b):
        if self.calib:
            self.mean_x[name] = x.reshape(-1, x.shape[-1]).mean(0)
        if self.mode == "fp32":
            return x @ w.t() + b
        wq = w.to(torch.bfloat16).float()
        y = x @ wq.t() + b
        if self.mode == "bf16+corr":
            if name not in self.corr:
                self.corr[name] = (wq - w) @ self.mean_x[name]
            y = y - self.corr[name]
        return y


def forward(lin, audio, n):
    g = lambda k: sd[k]
    h = audio[:, None, :]
    for l, s in enumerate(so.CONV_STRIDE):
        p = f"{W2V}feature_extractor.conv_layers.{l}."
        w = g(p + "conv.weight")
        k = w.shape[-1]
        cols = h.unfold(2, k, s).permute(0, 2, 1, 3).reshape(h.shape[0], -1, w.shape[1] * k)   # [B, T, Cin*k]
        y = lin(f"conv{l}", cols, w.reshape(w.shape[0], -1), g(p + "conv.bias"))
        y = F.layer_norm(y, (y.shape[-1],), g(p + "layer_norm.weight"), g(p + "layer_norm.bias"), 1e-5)
        h = F.gelu(y).transpose(1, 2)
    f = h.transpose(1, 2)
    B, T, _ = f.shape
    p = W2V + "feature_projection."
    z = F.layer_norm(f, (512,), g(p + "layer_norm.weight"), g(p + "layer_norm.bias"), 1e-5)
    z = lin("proj", z, g(p + "projection.weight"), g(p + "projection.bias"))
    w = so.pos_conv_weight(sd, torch.float32)
    if lin.mode != "fp32":
        w = w.to(torch.bfloat16).float()
    pc = F.conv1d(z.transpose(1, 2), w, g(W2V + "encoder.pos_conv_embed.conv.bias"), padding=64, groups=16)[:, :, :-1]
    z = z + F.gelu(pc).transpose(1, 2)
    valid = torch.ones(B, T, dtype=torch.bool)
    for i in range(spec.keep_layers):
        p = f"{W2V}encoder.layers.{i}."
        u = F.layer_norm(z, (1024,), g(p + "layer_norm.weight"), g(p + "layer_norm.bias"), 1e-5)
        q = lin(f"q{i}", u, g(p + "attention.q_proj.weight"), g(p + "attention.q_proj.bias"))
        k = lin(f"k{i}", u, g(p + "attention.k_proj.weight"), g(p + "attention.k_proj.bias"))
        v = lin(f"v{i}", u, g(p + "attention.v_proj.weight"), g(p + "attention.v_proj.bias"))
        sh = lambda t: t.view(B, T, 16, 64).transpose(1, 2)
        a = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / 8.0, -1) @ sh(v)
        z = z + lin(f"o{i}", a.transpose(1, 2).reshape(B, T, 1024), g(p + "attention.out_proj.weight"), g(p + "attention.out_proj.bias"))
        u = F.layer_norm(z, (1024,), g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias"), 1e-5)
        mid = F.gelu(lin(f"f1_{i}", u, g(p + "feed_forward.intermediate_dense.weight"), g(p + "feed_forward.intermediate_dense.bias")))
        ff = lin(f"f2_{i}", mid, g(p + "feed_forward.output_dense.weight"), g(p + "feed_forward.output_dense.bias"))
        if p + "ffn_adapter.down_proj.weight" in sd:
            a2 = torch.relu(lin(f"ad{i}", u, g(p + "ffn_adapter.down_proj.weight"), g(p + "ffn_adapter.down_proj.bias")))
            ff = ff + 4.0 * lin(f"au{i}", a2, g(p + "ffn_adapter.up_proj.weight"), g(p + "ffn_adapter.up_proj.bias"))
        z = z + ff
    return so.head(sd, z, torch.ones(B, T, dtype=torch.bool), spec.head_heads)[0]


def sig(n, seed, kind):
    x = synth.synthetic_audio(n, seed) if kind == "noise" else synth.speech_like_audio(n, seed)[0]
    return so.normalize_rows(x[None, :], [True])


n = 160_000
lin = Lin(sd)
with torch.no_grad():
    lin.calib = True
    forward(lin, sig(n, 777, "speech"), n)          # calibration signal: speech-like bursts, its own seed
    lin.calib = False
    for kind, seed in (("noise", 40), ("speech", 62), ("noise", 41)):
        x = sig(n, seed, kind)
        lin.mode = "fp32"; ref = forward(lin, x, n)
        lin.mode = "bf16"; q = forward(lin, x, n) - ref
        lin.mode = "bf16+corr"; c = forward(lin, x, n) - ref
        print(f"test signal {kind}/{seed}: bf16 weights: logit err mean {q.mean():+.4f} std {q.std():.4f} max {q.abs().max():.4f}"
              f"  | + bias correction: mean {c.mean():+.4f} std {c.std():.4f} max {c.abs().max():.4f}")
