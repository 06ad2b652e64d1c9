"""SHAS model (drop-in for the reference's lib.models.SHAS, lib/models.py:172-235) backed by the
sm_100a CUDA engine instead of transformers.Wav2Vec2Model + torch.nn.TransformerEncoder.

Kept from the reference: module path / class name / the 10 constructor kwargs (so a saved Hydra
config with `_target_: lib.models.SHAS` instantiates this class), `.wav2vec_model(audio, in_mask)
-> (None, hidden)`, `.seg_model(hidden, out_mask) -> logits`, `.to()`, `.eval()`,
`load_state_dict` on the model (full checkpoint) or on `.seg_model` (frozen-encoder checkpoint),
and the checkpoint key names (both weight-norm spellings of the positional conv).
Inference only: there is no autograd through the CUDA path and no CPU fallback.
"""
from __future__ import annotations

import dataclasses
import os

import torch
from torch import nn

from wav2vecsegmenter_b200.engine import SFCEngine
from wav2vecsegmenter_b200.synth import ModelSpec, random_state_dict

try:
    from constants import HIDDEN_SIZE
except ImportError:  # pragma: no cover
    from lib.constants import HIDDEN_SIZE


class _EngineHolder:
    """shared by the two sub-modules; creates the engine on the first .to(cuda)"""

    def __init__(self, spec: ModelSpec):
        self.spec = spec
        self.engine: SFCEngine | None = None
        self.pending: dict = {}
        self.dirty = False
        self.device = None

    def create(self, device=None) -> None:
        """`.to(device)`: only remembers the device. The reference moves the module BEFORE it loads the
        checkpoint (segment.py:41-51); the engine itself is built by the first forward / `.engine` access,
        when the weights — and with them the feature-extractor variant (LayerNorm vs GroupNorm convs, conv
        bias or not: HF config.feat_extract_norm / conv_bias) — are known."""
        dev = torch.device(device if device is not None else "cuda:0")
        if dev.type != "cuda":
            raise RuntimeError("lib.models.SHAS (B200 build) runs on CUDA only: there is no CPU path")
        self.device = dev

    def ensure(self, device=None) -> SFCEngine:
        """engine with every pending weight uploaded and finalised: called by the first forward /
        `.engine` access, i.e. after load_state_dict. Missing tensors surface here (W2VSegError)."""
        if self.engine is None:
            if device is not None or self.device is None:
                self.create(device)
            fe = "wav2vec_model.model.feature_extractor.conv_layers."
            if any(k.startswith(fe) for k in self.pending):      # variant as the checkpoint has it
                self.spec = dataclasses.replace(
                    self.spec,
                    feat_norm="layer" if fe + "1.layer_norm.weight" in self.pending else "group",
                    conv_bias=fe + "0.conv.bias" in self.pending)
            self.engine = SFCEngine(self.spec, self.device)
            self.dirty = True
        eng = self.engine
        if self.dirty:
            eng.load_encoder_state(self.pending, "wav2vec_model.model.")
            eng.load_head_state(self.pending, "seg_model.")
            eng.finalize()
            self.dirty = False
        return eng


def _encoder_is_post_ln(name: str) -> bool:
    """HF config.do_stable_layer_norm of the pretrained encoder (the checkpoint keys do not reveal it: pre-LN and
    post-LN layers have the same parameters). W2VSEG_POST_LN=0/1 overrides (offline / random-init runs)."""
    env = os.environ.get("W2VSEG_POST_LN")
    if env is not None:
        return env == "1"
    if os.environ.get("W2VSEG_RANDOM_INIT", "0") == "1":
        return False
    try:
        from transformers import Wav2Vec2Config

        return not Wav2Vec2Config.from_pretrained(name).do_stable_layer_norm
    except Exception:       # no local copy of the config: the reference's models are all XLS-R (pre-LN)
        return False


def _pretrained_encoder_state(name: str, spec: ModelSpec) -> dict:
    """weights of the pretrained wav2vec 2.0 / XLS-R encoder (the reference downloads them in
    HFWav2Vec2.__init__, lib/models.py:334). Sources, in order: a local directory / HF cache via
    transformers (weights only, no compute), or seeded random init when W2VSEG_RANDOM_INIT=1."""
    if os.environ.get("W2VSEG_RANDOM_INIT", "0") == "1":
        spec = dataclasses.replace(spec, feat_norm=os.environ.get("W2VSEG_FEAT_NORM", "layer"))
        sd = random_state_dict(spec, seed=int(os.environ.get("W2VSEG_SEED", "0")))
        return {k: v for k, v in sd.items() if k.startswith("wav2vec_model.model.")}
    from transformers import Wav2Vec2Model

    hf = Wav2Vec2Model.from_pretrained(name)
    if hf.config.hidden_size != spec.hidden:
        raise NotImplementedError(f"hidden size {hf.config.hidden_size} is not supported by the CUDA path (1024 only: "
                                  "XLS-R-300m / wav2vec2-large geometry)")
    return {"wav2vec_model.model." + k: v for k, v in hf.state_dict().items()}


class _Wav2VecModule(nn.Module):
    """callable like HFWav2Vec2[WithAdapter].forward (lib/models.py:367-368, 484-485)"""

    def __init__(self, holder: _EngineHolder):
        super().__init__()
        self._holder = holder

    @torch.no_grad()
    def forward(self, audio, attention_mask):
        eng = self._holder.ensure(audio.device if audio.is_cuda else None)
        audio = audio.to(eng.device, torch.float32)
        lens = attention_mask.to(eng.device).sum(dim=1).to(torch.int32)
        L = audio.shape[1]
        hidden, _ = eng.encode(audio, lens, None, L)   # audio is already normalised by CollateFn
        return None, hidden[:, : eng.num_frames(L)]


class SegmentationFrameClassifier(nn.Module):
    """callable like the reference's classifier head (lib/models.py:279-319)"""

    def __init__(self, holder: _EngineHolder):
        super().__init__()
        self._holder = holder

    @torch.no_grad()
    def forward(self, x, attention_mask):
        eng = self._holder.ensure(x.device if x.is_cuda else None)
        out_len = attention_mask.to(eng.device).bool().sum(dim=1).to(torch.int32)
        logits, _ = eng.head(x.to(eng.device, torch.float32), out_len)
        return logits

    def load_state_dict(self, state_dict, strict: bool = True):
        h = self._holder
        for k, v in state_dict.items():
            h.pending["seg_model." + k] = v.detach().cpu()
        h.dirty = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self, *a, **k):
        return {key[len("seg_model."):]: v for key, v in self._holder.pending.items() if key.startswith("seg_model.")}


class SHAS(nn.Module):
    def __init__(self, wav2vec_model_name, wav2vec_keep_layers, finetune_wav2vec, wav2vec_ft_layers,
                 finetune_w2v_feat_enc, finetune_w2v_ffn, ffn_adapter, n_transformer_enc_layers,
                 n_transformer_enc_heads, init_dropout) -> None:
        super().__init__()
        spec = ModelSpec.from_shas_kwargs(wav2vec_keep_layers, finetune_wav2vec, wav2vec_ft_layers,
                                          ffn_adapter, n_transformer_enc_layers, n_transformer_enc_heads)
        if _encoder_is_post_ln(wav2vec_model_name):
            # wav2vec2-base / -large-960h layers; the reference's adapter layer class exists for the stable-LN
            # layer only (lib/models.py:390-428), so finetune + ffn_adapter cannot be combined with them
            if spec.adapter_layers:
                raise NotImplementedError("FFN adapters need a stable-LayerNorm (pre-LN) encoder such as XLS-R")
            spec = dataclasses.replace(spec, post_ln=True)
        assert spec.hidden == HIDDEN_SIZE
        self.spec = spec
        self._holder = _EngineHolder(spec)
        self.wav2vec_model = _Wav2VecModule(self._holder)
        self.seg_model = SegmentationFrameClassifier(self._holder)
        self._finetune = bool(finetune_wav2vec)
        if not self._finetune:
            # frozen encoder: its weights are the pretrained ones, only the head comes from the ckpt
            self._holder.pending.update(_pretrained_encoder_state(wav2vec_model_name, spec))
            self._holder.dirty = True

    def to(self, device=None, *args, **kwargs):
        if device is not None and torch.device(device).type == "cuda":
            self._holder.create(device)
        elif device is not None:
            raise RuntimeError("lib.models.SHAS (B200 build) runs on CUDA only: there is no CPU path")
        return self

    def load_state_dict(self, state_dict, strict: bool = True):
        h = self._holder
        for k, v in state_dict.items():
            h.pending[k] = v.detach().cpu()
        h.dirty = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self, *a, **k):
        return dict(self._holder.pending)

    @property
    def engine(self) -> SFCEngine:
        return self._holder.ensure()

    @torch.no_grad()
    def forward(self, audio, in_mask, out_mask):
        """reference lib/models.py:214-235 (inference only)"""
        _, h = self.wav2vec_model(audio, in_mask)
        if h.shape[1] != out_mask.shape[1]:
            if h.shape[1] < out_mask.shape[1]:
                out_mask = out_mask[:, :-1]
            else:
                h = h[:, :-1, :]
        return self.seg_model(h, out_mask)
