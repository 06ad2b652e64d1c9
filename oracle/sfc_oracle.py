"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (wav2vecsegmenter_b200/,
lib/, segment.py, inference.py); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it, and only as the checker.

A plain, dependency-free (torch CPU ops only, no transformers, no reference import) restatement of
the SFC forward pass of ahclab/Wav2VecSegmenter, computed in fp32 (or fp64) straight from a
state dict in the reference's checkpoint layout. Each function cites the reference lines it
follows; `HF:` = transformers/models/wav2vec2/modeling_wav2vec2.py (v5.5.0 as installed; the
reference pins 4.36.1, same arithmetic).

Pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c) — upstream parity is unpinned. This oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF: oracle/make_golden.py imports the unmodified reference (HF Wav2Vec2Model +
lib.models.SHAS + lib.evaluate.infer + lib.segment) in the build container, runs it on seeded
weights/audio and commits the results under tests/golden/; tests/test_oracle.py checks this file
against those vectors on every CPU run.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
W2V = "wav2vec_model.model."
SEG = "seg_model."


def conv_out_frames(n: int) -> int:
    """HF:1005-1024 _get_feat_extract_output_lengths"""
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        if n < k:
            return 0
        n = (n - k) // s + 1
    return int(n)


def _get(sd, key, dtype):
    return sd[key].to(dtype)


def pos_conv_weight(sd, dtype):
    """weight_norm(dim=2): w = g * v / ||v||, norm over dims (0, 1) per tap (HF:343-355).
    Accepts both key spellings (torch 1.13 weight_g/weight_v, torch>=2.1 parametrizations)."""
    p = W2V + "encoder.pos_conv_embed.conv."
    if p + "weight" in sd:
        return _get(sd, p + "weight", dtype)
    if p + "weight_g" in sd:
        g, v = _get(sd, p + "weight_g", dtype), _get(sd, p + "weight_v", dtype)
    else:
        g = _get(sd, p + "parametrizations.weight.original0", dtype)
        v = _get(sd, p + "parametrizations.weight.original1", dtype)
    return g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()


def normalize_rows(audio: torch.Tensor, included) -> torch.Tensor:
    """CollateFn's per-row normalisation over the (already zero-padded) row, unbiased std
    (lib/datautils.py:122-125); rows with included=False are left untouched (:88)."""
    out = audio.clone()
    for i, inc in enumerate(included):
        if inc:
            row = audio[i]
            out[i] = (row - row.mean()) / row.std()
    return out


def feature_extractor(sd, x: torch.Tensor) -> torch.Tensor:
    """x [B, L] -> [B, T, 512]. feat_extract_norm "layer": 7 x [Conv1d -> LayerNorm over channels -> GELU]
    (HF:275-299, 382-419). "group" (recognised by the absence of conv_layers.1.layer_norm): conv 0 ->
    GroupNorm(num_groups = channels, i.e. per channel over TIME of the padded row) -> GELU, conv 1..6 -> GELU
    (HF:302-323, 249-272, 388-391). Conv biases are optional (config.conv_bias)."""
    dtype = x.dtype
    h = x[:, None, :]
    group = f"{W2V}feature_extractor.conv_layers.1.layer_norm.weight" not in sd
    for l, s in enumerate(CONV_STRIDE):
        p = f"{W2V}feature_extractor.conv_layers.{l}."
        bias = _get(sd, p + "conv.bias", dtype) if p + "conv.bias" in sd else None
        h = F.conv1d(h, _get(sd, p + "conv.weight", dtype), bias, stride=s)
        if group:
            if l == 0:
                h = F.group_norm(h, h.shape[1], _get(sd, p + "layer_norm.weight", dtype),
                                 _get(sd, p + "layer_norm.bias", dtype), 1e-5)
            h = F.gelu(h)
            continue
        h = h.transpose(1, 2)
        h = F.layer_norm(h, (h.shape[-1],), _get(sd, p + "layer_norm.weight", dtype),
                         _get(sd, p + "layer_norm.bias", dtype), 1e-5)
        h = F.gelu(h.transpose(1, 2))
    return h.transpose(1, 2)


def _mha(x, wq, bq, wk, bk, wv, bv, wo, bo, n_heads, key_valid):
    """softmax(QK^T/sqrt(d) + key mask) V, heads split on the feature axis (HF:500-549;
    torch.nn.MultiheadAttention). key_valid bool [B, T]."""
    B, T, D = x.shape
    d = D // n_heads
    q = (x @ wq.t() + bq).view(B, T, n_heads, d).transpose(1, 2)
    k = (x @ wk.t() + bk).view(B, T, n_heads, d).transpose(1, 2)
    v = (x @ wv.t() + bv).view(B, T, n_heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    s = s.masked_fill(~key_valid[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1) @ v
    return a.transpose(1, 2).reshape(B, T, D) @ wo.t() + bo


def encoder(sd, audio: torch.Tensor, sample_lens, keep_layers: int, n_heads: int = 16,
            post_ln: bool = False) -> torch.Tensor:
    """model.wav2vec_model(audio, in_mask) of lib/evaluate.py:59, i.e. HFWav2Vec2[WithAdapter]
    (lib/models.py:322-368, 431-485) around Wav2Vec2Model.forward (HF:1327-1383):
    audio [B, L] (already normalised), sample_lens[b] = in_mask[b].sum(). Returns [B, T, 1024]."""
    dtype = audio.dtype
    f = feature_extractor(sd, audio)                                          # HF:1348-1349
    B, T, _ = f.shape
    valid = torch.zeros(B, T, dtype=torch.bool, device=audio.device)        # HF:1026-1044
    for b, n in enumerate(sample_lens):
        valid[b, : conv_out_frames(int(n))] = True
    p = W2V + "feature_projection."
    z = F.layer_norm(f, (f.shape[-1],), _get(sd, p + "layer_norm.weight", dtype),
                     _get(sd, p + "layer_norm.bias", dtype), 1e-5)            # HF:429-434
    z = z @ _get(sd, p + "projection.weight", dtype).t() + _get(sd, p + "projection.bias", dtype)
    z = z * valid[:, :, None].to(dtype)                                       # HF:753-756
    w = pos_conv_weight(sd, dtype)                                            # HF:360-368
    groups = z.shape[-1] // w.shape[1]
    pc = F.conv1d(z.transpose(1, 2), w, _get(sd, W2V + "encoder.pos_conv_embed.conv.bias", dtype),
                  padding=w.shape[-1] // 2, groups=groups)
    if w.shape[-1] % 2 == 0:
        pc = pc[:, :, :-1]
    z = z + F.gelu(pc).transpose(1, 2)                                        # HF:764-765
    for i in range(keep_layers):                                              # HF:770-784
        p = f"{W2V}encoder.layers.{i}."
        g = lambda k: _get(sd, p + k, dtype)  # noqa: E731
        if post_ln:
            # do_stable_layer_norm = False (wav2vec2-base / -large-960h): HF Wav2Vec2EncoderLayer.forward;
            # encoder.layer_norm (applied BEFORE the layers in this variant) is an Identity in the reference
            # (lib/models.py:349), so the first layer sees the un-normalised positional-conv output
            z = z + _mha(z, g("attention.q_proj.weight"), g("attention.q_proj.bias"),
                         g("attention.k_proj.weight"), g("attention.k_proj.bias"),
                         g("attention.v_proj.weight"), g("attention.v_proj.bias"),
                         g("attention.out_proj.weight"), g("attention.out_proj.bias"), n_heads, valid)
            z = F.layer_norm(z, (z.shape[-1],), g("layer_norm.weight"), g("layer_norm.bias"), 1e-5)
            ff = F.gelu(z @ g("feed_forward.intermediate_dense.weight").t() + g("feed_forward.intermediate_dense.bias"))
            z = z + ff @ g("feed_forward.output_dense.weight").t() + g("feed_forward.output_dense.bias")
            z = F.layer_norm(z, (z.shape[-1],), g("final_layer_norm.weight"), g("final_layer_norm.bias"), 1e-5)
            continue
        u = F.layer_norm(z, (z.shape[-1],), g("layer_norm.weight"), g("layer_norm.bias"), 1e-5)
        z = z + _mha(u, g("attention.q_proj.weight"), g("attention.q_proj.bias"),
                     g("attention.k_proj.weight"), g("attention.k_proj.bias"),
                     g("attention.v_proj.weight"), g("attention.v_proj.bias"),
                     g("attention.out_proj.weight"), g("attention.out_proj.bias"), n_heads, valid)
        u = F.layer_norm(z, (z.shape[-1],), g("final_layer_norm.weight"), g("final_layer_norm.bias"), 1e-5)
        ff = F.gelu(u @ g("feed_forward.intermediate_dense.weight").t() + g("feed_forward.intermediate_dense.bias"))
        ff = ff @ g("feed_forward.output_dense.weight").t() + g("feed_forward.output_dense.bias")
        if p + "ffn_adapter.down_proj.weight" in sd:                          # lib/models.py:383-387, 415-421
            a = torch.relu(u @ g("ffn_adapter.down_proj.weight").t() + g("ffn_adapter.down_proj.bias"))
            ff = ff + 4.0 * (a @ g("ffn_adapter.up_proj.weight").t() + g("ffn_adapter.up_proj.bias"))
        z = z + ff
    return z  # final encoder LayerNorm removed -> Identity (lib/models.py:349, 463)


def head(sd, hidden: torch.Tensor, out_mask: torch.Tensor, n_heads: int = 8, prefix: str = SEG) -> torch.Tensor:
    """SegmentationFrameClassifier.forward (lib/models.py:307-319): one pre-LN
    TransformerEncoderLayer (gelu, dff 2048) with key-padding mask ~out_mask, LayerNorm,
    Linear(1024 -> 1), squeeze. hidden [B, T, 1024], out_mask bool [B, T] -> logits [B, T]."""
    dtype = hidden.dtype
    g = lambda k: _get(sd, prefix + k, dtype)  # noqa: E731
    x = hidden
    D = x.shape[-1]
    if prefix + "transformer.layers.0.self_attn.in_proj_weight" in sd:
        p = "transformer.layers.0."
        w_in, b_in = g(p + "self_attn.in_proj_weight"), g(p + "self_attn.in_proj_bias")
        u = F.layer_norm(x, (D,), g(p + "norm1.weight"), g(p + "norm1.bias"), 1e-5)
        x = x + _mha(u, w_in[:D], b_in[:D], w_in[D:2 * D], b_in[D:2 * D], w_in[2 * D:], b_in[2 * D:],
                     g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"), n_heads,
                     out_mask.bool())
        u = F.layer_norm(x, (D,), g(p + "norm2.weight"), g(p + "norm2.bias"), 1e-5)
        x = x + (F.gelu(u @ g(p + "linear1.weight").t() + g(p + "linear1.bias")) @ g(p + "linear2.weight").t()
                 + g(p + "linear2.bias"))
    x = F.layer_norm(x, (D,), g("layer_norm.weight"), g("layer_norm.bias"), 1e-5)
    return (x @ g("output_layer.weight").t() + g("output_layer.bias")).squeeze(-1)


def batch_probs(sd, audio, sample_lens, out_mask, keep_layers, head_heads=8, post_ln=False):
    """the per-batch body of lib.evaluate.infer (lib/evaluate.py:58-91) for loss_tag 'bce':
    returns (probs [B, T'], logits [B, T'], out_mask', ends_shift) where ends_shift = 1 if the
    reference decrements every `end` of the batch (:66-68)."""
    hidden = encoder(sd, audio, sample_lens, keep_layers, post_ln=post_ln)
    ends_shift = 0
    size1, size2 = hidden.shape[1], out_mask.shape[1]
    if size1 != size2:
        if size1 < size2:
            out_mask = out_mask[:, :-1]
            ends_shift = 1
        else:
            hidden = hidden[:, :-1, :]
    logits = head(sd, hidden, out_mask, head_heads)
    probs = torch.sigmoid(logits)
    probs = probs.masked_fill(~out_mask, 0.0)
    logits = logits.masked_fill(~out_mask, 0.0)
    return probs, logits, out_mask, ends_shift
