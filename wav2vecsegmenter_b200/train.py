"""Head-only training on the CUDA path (frozen encoder: the reference's `finetune_wav2vec=False` setting,
train.py:381-480 around lib/models.py:214-235, 279-319). The encoder forward is the inference kernel sequence
(no gradient), the head forward / loss / backward is `w2vseg_head_train_step`; the optimiser is the caller's
(torch.optim on fp32 master parameters kept here), and after every step the head parameters are re-uploaded
into the handle (bf16 matrices, fp32 vectors).

Dropout: the reference trains the head in train() mode — `init_dropout` (conf/task/shas.yaml:14, 0.1) on the encoder
output (lib/models.py:302,312) and the TransformerEncoderLayer's default 0.1 at its four sites. `HeadTrainer(...,
init_dropout=0.1, layer_dropout=0.1)` applies both with counter-hash masks (csrc/dropout.cuh); `dropout_masks()`
below rebuilds the same masks on the host side for the parity test. The defaults (0, 0) give the deterministic
eval-mode gradient.
"""
from __future__ import annotations

import torch

from . import _native as nat
from .engine import _HEAD_RULES, _canonical, SFCEngine


class HeadTrainer:
    """fp32 master copy of the seg_model parameters + one CUDA training step per call.

        trainer = HeadTrainer(engine, head_state_dict)            # keys as in SegmentationFrameClassifier
        opt = torch.optim.AdamW(trainer.parameters(), lr=...)
        loss = trainer.step(audio, sample_len, norm_len, out_len, target, pos_weight); opt.step(); trainer.sync()
    """

    def __init__(self, engine: SFCEngine, head_state: dict, init_dropout: float = 0.0, layer_dropout: float = 0.0,
                 seed: int = 0):
        self.init_dropout, self.layer_dropout = float(init_dropout), float(layer_dropout)
        self.seed = int(seed) & 0xFFFFFFFF       # seed of the NEXT step; advanced by every step
        self.engine = engine
        self.lib = engine.lib
        self.params: dict[str, torch.nn.Parameter] = {}
        self.names: dict[str, str] = {}          # state-dict key -> canonical head.* name
        for k, v in head_state.items():
            name = _canonical(k, _HEAD_RULES)
            if name is None:
                raise nat.W2VSegError(f"unrecognised head parameter '{k}'")
            self.params[k] = torch.nn.Parameter(v.detach().to(engine.device, torch.float32).clone())
            self.names[k] = name
        n = int(self.lib.w2vseg_head_grad_floats(engine._h))
        self._grads = torch.zeros(n, dtype=torch.float32, device=engine.device)
        self._loss = torch.zeros(1, dtype=torch.float32, device=engine.device)
        self._ws = None
        self._views = {}
        for k, name in self.names.items():
            numel = nat.C.c_int64(0)
            off = int(self.lib.w2vseg_head_grad_offset(engine._h, name.encode(), nat.C.byref(numel)))
            if off < 0 or numel.value != self.params[k].numel():
                raise nat.W2VSegError(f"no gradient slot for '{k}' ({name})")
            self._views[k] = self._grads[off: off + numel.value].view_as(self.params[k])
        self.sync()

    def parameters(self):
        return list(self.params.values())

    def state_dict(self) -> dict:
        return {k: p.detach().clone() for k, p in self.params.items()}

    def sync(self) -> None:
        """upload the current master parameters into the handle (after an optimiser step)"""
        eng = self.engine
        for k, p in self.params.items():
            src = p.detach().contiguous()
            nat.check(self.lib.w2vseg_set_weight(eng._h, self.names[k].encode(), src.data_ptr(), src.numel(), eng._stream()),
                      f"set_weight({self.names[k]})")

    def step_hidden(self, hidden: torch.Tensor, out_len, target: torch.Tensor, pos_weight: float = 1.0,
                    logits_out: torch.Tensor | None = None) -> torch.Tensor:
        """hidden fp32 [B, T, 1024] (encoder output, no gradient), target fp32 [B, T]. Fills `.grad` of every
        parameter (accumulating, like loss.backward()) and returns the loss as a 0-dim device tensor."""
        eng = self.engine
        assert hidden.is_cuda and hidden.dtype == torch.float32 and hidden.dim() == 3
        if hidden.stride(2) != 1 or hidden.stride(1) != hidden.shape[2]:
            hidden = hidden.contiguous()
        B, T, _ = hidden.shape
        ol = eng._i32(out_len, eng.device)
        tg = target.to(eng.device, torch.float32).contiguous()
        assert tg.shape == (B, T)
        need = int(self.lib.w2vseg_head_train_workspace_bytes(eng._h, B, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=eng.device)
        nat.check(self.lib.w2vseg_head_train_step(eng._h, hidden.data_ptr(), hidden.stride(0), T, ol.data_ptr(),
                                                  tg.data_ptr(), float(pos_weight), B, self._loss.data_ptr(),
                                                  nat.ptr(logits_out), self._grads.data_ptr(), self._grads.numel(),
                                                  self.init_dropout, self.layer_dropout, self.seed,
                                                  self._ws.data_ptr(), self._ws.numel(), eng._stream()),
                  "w2vseg_head_train_step")
        self.seed = (self.seed + 1) & 0xFFFFFFFF
        for k, p in self.params.items():
            g = self._views[k]
            p.grad = g.clone() if p.grad is None else p.grad + g
        return self._loss[0].clone()

    def step(self, audio: torch.Tensor, sample_len, norm_len, out_len, target: torch.Tensor,
             pos_weight: float = 1.0) -> torch.Tensor:
        """model(audio, in_mask, out_mask) + loss + backward of train.py:399-462 for the frozen-encoder model:
        raw audio fp32 [B, L] on the device -> loss; `.grad` filled. Applies the reference's +-1 frame fix-up
        (lib/models.py:224-231, train.py:419-440) between the encoder length and the target length."""
        eng = self.engine
        L = audio.shape[1]
        hidden, _ = eng.encode(audio, sample_len, norm_len, L)
        Th, Tt = eng.num_frames(L), target.shape[1]
        ol = eng._i32(out_len, eng.device)
        if Th < Tt:                       # logits shorter than the target: drop the target's last frame
            target = target[:, :-1]
            ol = torch.clamp(ol, max=Tt - 1)
            Tt -= 1
        return self.step_hidden(hidden[:, :Tt], ol, target, pos_weight)


def _lowbias32(x: torch.Tensor) -> torch.Tensor:
    """the integer hash of csrc/dropout.cuh on int64 tensors holding uint32 values"""
    m = 0xFFFFFFFF
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & m
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & m
    return x ^ (x >> 16)


def dropout_mask(p: float, seed: int, site: int, numel: int, device="cpu") -> torch.Tensor:
    """factor (0 or 1/(1-p)) the CUDA step applies to flat element index 0..numel-1 of dropout site `site`
    (0 encoder output [B*T, D]; 1 attention weights [B, heads, T, T]; 2 attention-block output [B*T, D];
    3 FFN inner activation [B*T, F]; 4 FFN output [B*T, D])"""
    if p <= 0:
        return torch.ones(numel, device=device)
    key = int(_lowbias32(torch.tensor([(seed * 0x9E3779B9 + site) & 0xFFFFFFFF], dtype=torch.int64))[0])
    idx = torch.arange(numel, dtype=torch.int64, device=device) & 0xFFFFFFFF
    keep = _lowbias32(idx ^ key) >= int(p * 4294967296.0)
    return keep.float() / (1.0 - p)
