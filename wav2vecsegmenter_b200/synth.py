"""Deterministic random-init checkpoints and synthetic audio for tests and benchmarks.

There is no network (no pretrained XLS-R, no released SFC checkpoints), so every parity test and
benchmark uses seeded random weights *in the reference's checkpoint layout* (keys
`wav2vec_model.model.*` + `seg_model.*`, as produced by train.py:596-604) and seeded synthetic
16 kHz audio. The generator is pure torch-CPU so the same seed gives the same tensors in the build
container (where the reference generates golden vectors) and on the GPU box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)


@dataclass(frozen=True)
class ModelSpec:
    """arithmetic-relevant architecture of one SHAS model (lib/models.py:173-212)"""

    keep_layers: int = 24          # wav2vec_keep_layers
    adapter_layers: int = 0        # layers [keep-adapter_layers, keep) carry a ScaledParallelAdapter
    hidden: int = 1024
    heads: int = 16
    ffn: int = 4096
    adapter_dim: int = 512
    adapter_scale: float = 4.0
    conv_dim: int = 512
    pos_kernel: int = 128
    pos_groups: int = 16
    head_layers: int = 1           # n_transformer_enc_layers
    head_heads: int = 8            # n_transformer_enc_heads
    head_ffn: int = 2048
    ln_eps: float = 1e-5
    # feature-extractor variant (HF config.feat_extract_norm): "layer" = LayerNorm + GELU after every conv
    # (XLS-R, HF:275-299), "group" = GroupNorm(512 groups) after conv 0 only (HF:302-323, 388-391)
    feat_norm: str = "layer"
    conv_bias: bool = True
    # encoder layer variant (HF config.do_stable_layer_norm): False = pre-LN "stable layer norm" layers (XLS-R),
    # True = post-LN layers h = LN(h + Attn(h)); h = LN(h + FFN(h)) (wav2vec2-base / -large-960h, HF:575-609)
    post_ln: bool = False

    @staticmethod
    def from_shas_kwargs(wav2vec_keep_layers, finetune_wav2vec, wav2vec_ft_layers, ffn_adapter,
                         n_transformer_enc_layers=1, n_transformer_enc_heads=8, **_):
        """same selection rule as SHAS.__init__ (lib/models.py:188) / HFWav2Vec2WithAdapter
        (lib/models.py:445-461): adapters exist iff finetune_wav2vec and ffn_adapter, in the last
        wav2vec_ft_layers kept layers"""
        n_ad = 0
        if finetune_wav2vec and ffn_adapter:
            n_ad = max(0, min(int(wav2vec_keep_layers), int(wav2vec_ft_layers)))
        return ModelSpec(keep_layers=int(wav2vec_keep_layers), adapter_layers=n_ad,
                         head_layers=int(n_transformer_enc_layers),
                         head_heads=int(n_transformer_enc_heads))


# the benchmark / parity configurations of BASELINE.json
LARGE_ALL = ModelSpec(keep_layers=24, adapter_layers=24)      # large (24/24) + adapters
MIDDLE = ModelSpec(keep_layers=16, adapter_layers=0)          # middle (0/16), frozen encoder
MIDDLE_HALF = ModelSpec(keep_layers=16, adapter_layers=8)     # middle+half (8/16)
TINY = ModelSpec(keep_layers=2, adapter_layers=1)             # test-sized
TINY_GN = ModelSpec(keep_layers=2, adapter_layers=1, feat_norm="group")                     # GroupNorm extractor
TINY_GN_NOBIAS = ModelSpec(keep_layers=2, adapter_layers=0, feat_norm="group", conv_bias=False)
# the wav2vec2-large-960h architecture family: GroupNorm extractor without conv bias + post-LN encoder layers
TINY_POSTLN = ModelSpec(keep_layers=3, adapter_layers=0, feat_norm="group", conv_bias=False, post_ln=True)


def random_state_dict(spec: ModelSpec, seed: int = 0, logit_std: float = 2.0) -> dict:
    """full-model state dict (finetune_wav2vec layout). Weight-norm uses the torch>=2.1 key
    spelling (`parametrizations.weight.original0/1`); the loader also accepts weight_g/weight_v."""
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def linear(sd, prefix, out_f, in_f, gain=1.0):
        sd[prefix + ".weight"] = randn(out_f, in_f, std=gain / math.sqrt(in_f))
        sd[prefix + ".bias"] = randn(out_f, std=0.05)

    def lnorm(sd, prefix, n):
        sd[prefix + ".weight"] = 1.0 + randn(n, std=0.1)
        sd[prefix + ".bias"] = randn(n, std=0.1)

    sd = {}
    w = "wav2vec_model.model."
    D, C = spec.hidden, spec.conv_dim
    sd[w + "masked_spec_embed"] = torch.rand(D, generator=g)
    cin = 1
    for l, k in enumerate(CONV_KERNEL):
        p = f"{w}feature_extractor.conv_layers.{l}"
        sd[p + ".conv.weight"] = randn(C, cin, k, std=1.4 / math.sqrt(cin * k))
        if spec.conv_bias:
            sd[p + ".conv.bias"] = randn(C, std=0.05)
        if spec.feat_norm == "layer" or l == 0:     # "group": GroupNorm affine after conv 0 only (same key names)
            lnorm(sd, p + ".layer_norm", C)
        cin = C
    lnorm(sd, w + "feature_projection.layer_norm", C)
    linear(sd, w + "feature_projection.projection", D, C)
    gc = D // spec.pos_groups
    v = randn(D, gc, spec.pos_kernel, std=1.0 / math.sqrt(gc * spec.pos_kernel))
    norm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    sd[w + "encoder.pos_conv_embed.conv.bias"] = randn(D, std=0.05)
    sd[w + "encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = norm * (
        1.0 + randn(1, 1, spec.pos_kernel, std=0.1))
    sd[w + "encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    for i in range(spec.keep_layers):
        p = f"{w}encoder.layers.{i}"
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            linear(sd, f"{p}.attention.{nm}", D, D)
        lnorm(sd, p + ".layer_norm", D)
        linear(sd, p + ".feed_forward.intermediate_dense", spec.ffn, D)
        linear(sd, p + ".feed_forward.output_dense", D, spec.ffn)
        lnorm(sd, p + ".final_layer_norm", D)
        if i >= spec.keep_layers - spec.adapter_layers:
            linear(sd, p + ".ffn_adapter.down_proj", spec.adapter_dim, D)
            linear(sd, p + ".ffn_adapter.up_proj", D, spec.adapter_dim, gain=0.25)
    sd.update(random_head_state_dict(spec, seed + 1, logit_std, prefix="seg_model."))
    return sd


def random_head_state_dict(spec: ModelSpec, seed: int = 1, logit_std: float = 2.0,
                           prefix: str = "") -> dict:
    """seg_model-only state dict (the frozen-encoder checkpoint layout, train.py:605-613)"""
    g = torch.Generator().manual_seed(seed)
    D = spec.hidden

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    if spec.head_layers:
        p = prefix + "transformer.layers.0."
        sd[p + "self_attn.in_proj_weight"] = randn(3 * D, D, std=1.0 / math.sqrt(D))
        sd[p + "self_attn.in_proj_bias"] = randn(3 * D, std=0.05)
        sd[p + "self_attn.out_proj.weight"] = randn(D, D, std=1.0 / math.sqrt(D))
        sd[p + "self_attn.out_proj.bias"] = randn(D, std=0.05)
        sd[p + "linear1.weight"] = randn(spec.head_ffn, D, std=1.0 / math.sqrt(D))
        sd[p + "linear1.bias"] = randn(spec.head_ffn, std=0.05)
        sd[p + "linear2.weight"] = randn(D, spec.head_ffn, std=1.0 / math.sqrt(spec.head_ffn))
        sd[p + "linear2.bias"] = randn(D, std=0.05)
        for nm in ("norm1", "norm2"):
            sd[p + nm + ".weight"] = 1.0 + randn(D, std=0.1)
            sd[p + nm + ".bias"] = randn(D, std=0.1)
    sd[prefix + "layer_norm.weight"] = 1.0 + randn(D, std=0.1)
    sd[prefix + "layer_norm.bias"] = randn(D, std=0.1)
    sd[prefix + "output_layer.weight"] = randn(1, D, std=logit_std / math.sqrt(D))
    sd[prefix + "output_layer.bias"] = randn(1, std=0.3)
    return sd


def synthetic_audio(n_samples: int, seed: int = 0, modulated: bool = True) -> torch.Tensor:
    """fp32 mono 16 kHz 'speech-like' signal in [-1, 1): band-limited noise whose amplitude is
    gated at a syllable/pause time scale, so frame probabilities are not flat"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_samples, generator=g)
    if modulated and n_samples >= 3200:
        n_env = n_samples // 1600 + 2  # 0.1 s envelope grid
        env = torch.rand(n_env, generator=g)
        env = (env > 0.35).float() * (0.3 + 0.7 * torch.rand(n_env, generator=g))
        env = torch.nn.functional.interpolate(env[None, None], size=n_samples, mode="linear",
                                              align_corners=False)[0, 0]
        x = x * (0.02 + env)
    x = 0.25 * x / x.abs().max().clamp_min(1e-6) * 3.0
    return x.clamp_(-0.999, 0.999)


def speech_like_audio(n_samples: int, seed: int = 0):
    """fp32 mono 16 kHz signal with talk-like structure: noise bursts of 1.5-9 s ("speech", random
    level) separated by 0.25-1.5 s of near-silence ("pauses"), 20 ms ramps at the edges. Returns
    (samples, is_speech per 20 ms output frame at 49.95 Hz). Used by the boundary-parity fixture:
    with a calibrated output layer the frame probabilities are decisive like a trained model's."""
    g = torch.Generator().manual_seed(seed)
    env = torch.zeros(n_samples)
    pos = 0
    while pos < n_samples:
        sp = int((1.5 + 7.5 * torch.rand(1, generator=g).item()) * 16000)
        pa = int((0.25 + 1.25 * torch.rand(1, generator=g).item()) * 16000)
        lvl = 0.3 + 0.7 * torch.rand(1, generator=g).item()
        env[pos:pos + sp] = lvl
        pos += sp + pa
    k = torch.ones(1, 1, 321) / 321.0   # 20 ms moving average = linear ramps
    env_s = torch.nn.functional.conv1d(env[None, None], k, padding=160)[0, 0]
    x = torch.randn(n_samples, generator=g) * (0.002 + env_s) * 0.25
    n_frames = int(round(n_samples * 49.95 / 16000))
    centers = ((torch.arange(n_frames).double() + 0.5) * 16000 / 49.95).long().clamp_(0, n_samples - 1)
    return x.clamp_(-0.999, 0.999), (env[centers] > 0).numpy()
