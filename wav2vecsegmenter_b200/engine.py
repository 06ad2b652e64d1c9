"""Host-side owner of one SFC model replica on one B200: wraps the opaque handle of libw2vseg.so,
maps reference checkpoint keys onto the library's canonical tensor names, owns the scratch
workspace (a torch uint8 tensor — PyTorch is only the allocator / stream provider here) and
exposes the three forward entry points plus the talk-level reduction kernels.

Reference boundary this mirrors: lib/models.py:172-235 (SHAS), lib/evaluate.py:58-91.
"""
from __future__ import annotations

import os
import re

import torch

from . import _native as nat
from .synth import ModelSpec

_W2V_RULES = [
    (r"^feature_extractor\.conv_layers\.(\d)\.conv\.(weight|bias)$", r"fe.conv\1.\2"),
    (r"^feature_extractor\.conv_layers\.(\d)\.layer_norm\.(weight|bias)$", r"fe.conv\1.ln.\2"),
    (r"^feature_projection\.layer_norm\.(weight|bias)$", r"fp.ln.\1"),
    (r"^feature_projection\.projection\.(weight|bias)$", r"fp.proj.\1"),
    (r"^encoder\.pos_conv_embed\.conv\.bias$", "pos.bias"),
    (r"^encoder\.pos_conv_embed\.conv\.(weight_g|parametrizations\.weight\.original0)$", "pos.weight_g"),
    (r"^encoder\.pos_conv_embed\.conv\.(weight_v|parametrizations\.weight\.original1)$", "pos.weight_v"),
    (r"^encoder\.pos_conv_embed\.conv\.weight$", "pos.weight"),
    (r"^encoder\.layers\.(\d+)\.attention\.([qkv])_proj\.(weight|bias)$", r"enc.\1.\2.\3"),
    (r"^encoder\.layers\.(\d+)\.attention\.out_proj\.(weight|bias)$", r"enc.\1.o.\2"),
    (r"^encoder\.layers\.(\d+)\.layer_norm\.(weight|bias)$", r"enc.\1.ln1.\2"),
    (r"^encoder\.layers\.(\d+)\.final_layer_norm\.(weight|bias)$", r"enc.\1.ln2.\2"),
    (r"^encoder\.layers\.(\d+)\.feed_forward\.intermediate_dense\.(weight|bias)$", r"enc.\1.ff1.\2"),
    (r"^encoder\.layers\.(\d+)\.feed_forward\.output_dense\.(weight|bias)$", r"enc.\1.ff2.\2"),
    (r"^encoder\.layers\.(\d+)\.ffn_adapter\.down_proj\.(weight|bias)$", r"enc.\1.ad_down.\2"),
    (r"^encoder\.layers\.(\d+)\.ffn_adapter\.up_proj\.(weight|bias)$", r"enc.\1.ad_up.\2"),
]
_W2V_IGNORED = re.compile(r"^(masked_spec_embed|encoder\.layer_norm\.(weight|bias))$")

_HEAD_RULES = [
    (r"^transformer\.layers\.0\.self_attn\.in_proj_(weight|bias)$", r"head.in_proj.\1"),
    (r"^transformer\.layers\.0\.self_attn\.out_proj\.(weight|bias)$", r"head.o.\1"),
    (r"^transformer\.layers\.0\.linear1\.(weight|bias)$", r"head.ff1.\1"),
    (r"^transformer\.layers\.0\.linear2\.(weight|bias)$", r"head.ff2.\1"),
    (r"^transformer\.layers\.0\.norm1\.(weight|bias)$", r"head.ln1.\1"),
    (r"^transformer\.layers\.0\.norm2\.(weight|bias)$", r"head.ln2.\1"),
    (r"^layer_norm\.(weight|bias)$", r"head.ln_f.\1"),
    (r"^output_layer\.(weight|bias)$", r"head.out.\1"),
]


def _canonical(key: str, rules):
    for pat, rep in rules:
        if re.match(pat, key):
            return re.sub(pat, rep, key)
    return None


class SFCEngine:
    def __init__(self, spec: ModelSpec, device: str | torch.device = "cuda:0"):
        self.lib = nat.load()
        self.spec = spec
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise nat.W2VSegError("SFCEngine needs a CUDA (sm_100a) device; there is no CPU fallback")
        torch.cuda.set_device(self.device)
        cfg = nat.Config(
            n_layers=spec.keep_layers, n_adapter_layers=spec.adapter_layers, hidden=spec.hidden,
            heads=spec.heads, ffn=spec.ffn, adapter_dim=spec.adapter_dim,
            adapter_scale=spec.adapter_scale, conv_dim=spec.conv_dim, pos_kernel=spec.pos_kernel,
            pos_groups=spec.pos_groups, head_layers=spec.head_layers, head_heads=spec.head_heads,
            head_ffn=spec.head_ffn, ln_eps=spec.ln_eps,
            feat_group_norm=int(spec.feat_norm == "group"), conv_bias=int(spec.conv_bias),
            post_layer_norm=int(spec.post_ln),
        )
        h = nat.C.c_void_p()
        nat.check(self.lib.w2vseg_create(nat.C.byref(cfg), nat.C.byref(h)), "w2vseg_create")
        self._h = h
        self._ws = None
        self._finalized = False
        self._matrices = {}     # canonical name -> fp32 source tensor of every matrix uploaded since the last finalize

    def close(self):
        if getattr(self, "_h", None):
            self.lib.w2vseg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _set(self, name: str, t: torch.Tensor):
        src = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        nat.check(
            self.lib.w2vseg_set_weight(self._h, name.encode(), src.data_ptr(), src.numel(), self._stream()),
            f"set_weight({name})",
        )
        # src may be a temporary: the packing kernel is queued on the current stream, and the
        # caching allocator only reuses the block for later work on that same stream -> safe.
        self._finalized = False
        if t.dim() >= 2:
            self._matrices[name] = t

    def load_encoder_state(self, sd: dict, prefix: str = "wav2vec_model.model."):
        """HF Wav2Vec2Model parameters (keys after `prefix`); layers >= keep_layers are ignored
        like lib/models.py:340-346 drops them."""
        for k, v in sd.items():
            if not k.startswith(prefix):
                continue
            key = k[len(prefix):]
            if _W2V_IGNORED.match(key):
                continue
            m = re.match(r"^encoder\.layers\.(\d+)\.", key)
            if m and int(m.group(1)) >= self.spec.keep_layers:
                continue
            name = _canonical(key, _W2V_RULES)
            if name is None:
                raise nat.W2VSegError(f"unrecognised encoder parameter '{k}'")
            if re.match(r"^enc\.(\d+)\.ad_", name) and int(name.split(".")[1]) < self.spec.keep_layers - self.spec.adapter_layers:
                raise nat.W2VSegError(f"checkpoint has an adapter in layer {name.split('.')[1]} but the model spec does not")
            self._set(name, v)

    def load_head_state(self, sd: dict, prefix: str = "seg_model."):
        """SegmentationFrameClassifier parameters (lib/models.py:279-305 layout)"""
        for k, v in sd.items():
            if not k.startswith(prefix):
                continue
            name = _canonical(k[len(prefix):], _HEAD_RULES)
            if name is None:
                raise nat.W2VSegError(f"unrecognised head parameter '{k}'")
            self._set(name, v)

    def load_state_dict(self, sd: dict):
        """full-model checkpoint layout (train.py:596-604)"""
        self.load_encoder_state(sd, "wav2vec_model.model.")
        self.load_head_state(sd, "seg_model.")
        self.finalize()

    def finalize(self, bias_correction: bool | None = None):
        """completeness check + weight folding; then (default on, W2VSEG_BIAS_CORRECTION=0 turns it off)
        the bias correction for the bf16 rounding of the matrices (include/w2vseg.h): one calibration
        forward over a fixed synthetic speech-like signal + one small kernel per matrix."""
        nat.check(self.lib.w2vseg_finalize_weights(self._h, self._stream()), "finalize_weights")
        self._finalized = True
        if bias_correction is None:
            bias_correction = os.environ.get("W2VSEG_BIAS_CORRECTION", "1") != "0"
        if bias_correction:
            self.correct_biases()
        self._matrices = {}

    def correct_biases(self, calib_audio: torch.Tensor | None = None):
        """calib_audio: fp32 [B, L] raw samples on any device (default: two 20 s windows of seeded noise
        bursts and pauses, wav2vecsegmenter_b200.synth.speech_like_audio, seeds unrelated to any test)"""
        from .synth import speech_like_audio

        if calib_audio is None:
            L = 320_000
            calib_audio = torch.stack([speech_like_audio(L, 7770 + i)[0] for i in range(2)])
        audio = calib_audio.to(self.device, torch.float32).contiguous()
        B, L = audio.shape
        lens = torch.full((B,), L, dtype=torch.int32, device=self.device)
        out_len = torch.full((B,), self.num_frames(L), dtype=torch.int32, device=self.device)
        ws = self._workspace(B, L)
        nat.check(self.lib.w2vseg_calibrate(self._h, audio.data_ptr(), audio.stride(0), lens.data_ptr(), lens.data_ptr(),
                                            out_len.data_ptr(), B, L, ws.data_ptr(), ws.numel(), self._stream()),
                  "w2vseg_calibrate")
        for name, t in self._matrices.items():
            src = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
            nat.check(self.lib.w2vseg_correct_bias(self._h, name.encode(), src.data_ptr(), src.numel(), self._stream()),
                      f"correct_bias({name})")

    # ------------------------------------------------------------------ geometry / scratch
    def frame_stride(self, l_max: int) -> int:
        return int(self.lib.w2vseg_frame_stride(int(l_max)))

    def num_frames(self, n_samples: int) -> int:
        return int(self.lib.w2vseg_num_frames(int(n_samples)))

    def _workspace(self, B: int, l_max: int) -> torch.Tensor:
        need = int(self.lib.w2vseg_workspace_bytes(self._h, B, int(l_max)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _i32(x, device):
        if isinstance(x, torch.Tensor):
            if x.dtype == torch.int32 and x.is_cuda and x.is_contiguous():
                return x
            return x.to(device=device, dtype=torch.int32).contiguous()
        return torch.tensor(list(x), dtype=torch.int32).to(device, non_blocking=True)

    # ------------------------------------------------------------------ forward
    def encode(self, audio: torch.Tensor, sample_len, norm_len=None, l_max: int | None = None):
        """audio fp32 [B, L] on device (raw if norm_len given, else treated as already normalised).
        Returns (hidden fp32 [B, R, 1024], enc_len int32 [B])."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2
        audio = audio if audio.stride(1) == 1 else audio.contiguous()
        B = audio.shape[0]
        sl = self._i32(sample_len, self.device)
        nl = torch.zeros_like(sl) if norm_len is None else self._i32(norm_len, self.device)
        l_max = int(l_max if l_max is not None else sl.max().item())
        R = self.frame_stride(l_max)
        hidden = torch.empty(B, R, self.spec.hidden, dtype=torch.float32, device=self.device)
        enc_len = torch.empty(B, dtype=torch.int32, device=self.device)
        ws = self._workspace(B, l_max)
        nat.check(
            self.lib.w2vseg_encode(self._h, audio.data_ptr(), audio.stride(0), sl.data_ptr(), nl.data_ptr(),
                                   B, l_max, hidden.data_ptr(), enc_len.data_ptr(), None, ws.data_ptr(),
                                   ws.numel(), self._stream()),
            "w2vseg_encode",
        )
        return hidden, enc_len

    def head(self, hidden: torch.Tensor, out_len):
        """hidden fp32 [B, T, 1024] (any batch stride, rows contiguous). Returns (logits, probs) [B, T]."""
        assert hidden.is_cuda and hidden.dtype == torch.float32 and hidden.dim() == 3
        if hidden.stride(2) != 1 or hidden.stride(1) != hidden.shape[2]:
            hidden = hidden.contiguous()
        B, T, _ = hidden.shape
        ol = self._i32(out_len, self.device)
        logits = torch.empty(B, T, dtype=torch.float32, device=self.device)
        probs = torch.empty(B, T, dtype=torch.float32, device=self.device)
        ws = self._workspace(B, max(T * 320, 400))
        nat.check(
            self.lib.w2vseg_head(self._h, hidden.data_ptr(), hidden.stride(0), T, ol.data_ptr(), B,
                                 logits.data_ptr(), probs.data_ptr(), ws.data_ptr(), ws.numel(),
                                 self._stream()),
            "w2vseg_head",
        )
        return logits, probs

    def sfc_forward(self, audio: torch.Tensor, sample_len, norm_len, out_len, l_max: int,
                    logits_out: torch.Tensor | None = None, probs_out: torch.Tensor | None = None,
                    included_out: torch.Tensor | None = None):
        """fused encode + head. Returns (logits, probs), each fp32 [B, R]. included_out (optional
        int32 [B] on the device) receives CollateFn's `included` flag, decided on the device."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
        B = audio.shape[0]
        sl = self._i32(sample_len, self.device)
        nl = self._i32(norm_len, self.device)
        ol = self._i32(out_len, self.device)
        R = self.frame_stride(l_max)
        if logits_out is None:
            logits_out = torch.empty(B, R, dtype=torch.float32, device=self.device)
        if probs_out is None:
            probs_out = torch.empty(B, R, dtype=torch.float32, device=self.device)
        ws = self._workspace(B, l_max)
        nat.check(
            self.lib.w2vseg_sfc_forward(self._h, audio.data_ptr(), audio.stride(0), sl.data_ptr(),
                                        nl.data_ptr(), ol.data_ptr(), B, int(l_max),
                                        logits_out.data_ptr(), probs_out.data_ptr(), nat.ptr(included_out),
                                        ws.data_ptr(), ws.numel(), self._stream()),
            "w2vseg_sfc_forward",
        )
        return logits_out, probs_out

    def sfc_forward_rows(self, audio: torch.Tensor, sample_len, norm_len, out_len, l_max: int,
                         rows_out: torch.Tensor, flag_col: int = -1):
        """fused encode + head writing probabilities (and the `included` flag in column flag_col)
        straight into `rows_out` fp32 [B, cols] (row stride = rows_out.stride(0)): the row matrix
        scatter_rows consumes. Columns beyond the batch's frame stride are zeroed."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
        assert rows_out.is_cuda and rows_out.dtype == torch.float32 and rows_out.dim() == 2 and rows_out.stride(1) == 1
        B = audio.shape[0]
        assert rows_out.shape[0] == B
        sl = self._i32(sample_len, self.device)
        nl = self._i32(norm_len, self.device)
        ol = self._i32(out_len, self.device)
        ws = self._workspace(B, l_max)
        nat.check(
            self.lib.w2vseg_sfc_forward_rows(self._h, audio.data_ptr(), audio.stride(0), sl.data_ptr(),
                                             nl.data_ptr(), ol.data_ptr(), B, int(l_max), rows_out.data_ptr(),
                                             rows_out.stride(0), rows_out.shape[1], int(flag_col),
                                             ws.data_ptr(), ws.numel(), self._stream()),
            "w2vseg_sfc_forward_rows",
        )
        return rows_out

    def graphed(self, B: int, l_max: int) -> "GraphedForward":
        """the fused forward for a FIXED batch shape captured into a CUDA graph (see GraphedForward)"""
        return GraphedForward(self, B, l_max)

    # ------------------------------------------------------------------ talk-level reductions
    def scatter_rows(self, rows: torch.Tensor, start, count, n_frames: int, flag_col: int = -1) -> torch.Tensor:
        """rows fp32 [W, stride] -> talk vector fp64 [n_frames] (NaN where no window wrote);
        flag_col >= 0: that column of each row holds the window's `included` flag"""
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.stride(-1) == 1
        st = self._i32(start, self.device)
        ct = self._i32(count, self.device)
        talk = torch.empty(n_frames, dtype=torch.float64, device=self.device)
        nat.check(
            self.lib.w2vseg_scatter_rows(rows.data_ptr(), rows.stride(0), st.data_ptr(), ct.data_ptr(),
                                         st.numel(), talk.data_ptr(), n_frames, int(flag_col), self._stream()),
            "w2vseg_scatter_rows",
        )
        return talk

    def nanfill(self, talk: torch.Tensor, idx) -> torch.Tensor:
        ix = self._i32(idx, self.device)
        if ix.numel():
            nat.check(self.lib.w2vseg_nanfill(talk.data_ptr(), talk.numel(), ix.data_ptr(), ix.numel(),
                                              self._stream()), "w2vseg_nanfill")
        return talk

    def overlap_average(self, tilings: torch.Tensor) -> torch.Tensor:
        assert tilings.is_cuda and tilings.dtype == torch.float64 and tilings.dim() == 2 and tilings.is_contiguous()
        out = torch.empty(tilings.shape[1], dtype=torch.float64, device=self.device)
        nat.check(self.lib.w2vseg_overlap_average(tilings.data_ptr(), tilings.shape[0], tilings.shape[1],
                                                  out.data_ptr(), self._stream()), "w2vseg_overlap_average")
        return out

    def moving_average(self, arr: torch.Tensor, window: int) -> torch.Tensor:
        assert arr.is_cuda and arr.dtype == torch.float64 and arr.dim() == 1 and arr.is_contiguous()
        out = torch.empty_like(arr)
        nat.check(self.lib.w2vseg_moving_average(arr.data_ptr(), arr.numel(), int(window), out.data_ptr(),
                                                 self._stream()), "w2vseg_moving_average")
        return out


class GraphedForward:
    """`w2vseg_sfc_forward` for one fixed (B, l_max) captured into a CUDA graph: ONE graph launch instead of ~196
    kernel launches. The forward allocates nothing and never synchronises, so it is capturable as it is. This is
    the latency path: at batch 1-2 the step is bound by the host's launch rate (~12 us per launch, batch-1 forward
    3.1 ms direct), not by the GPU; at the benchmark's batch 14 the step is GPU-bound and a graph changes nothing
    (profiles/experiments_r01_s3.md). Usage: fill `.audio` / `.sample_len` / `.norm_len` / `.out_len` (static device
    tensors) with copy_(), call `.replay()`, read `.probs` / `.logits` ([B, R])."""

    def __init__(self, engine: SFCEngine, B: int, l_max: int):
        self.engine = engine
        dev = engine.device
        self.B, self.l_max = int(B), int(l_max)
        R = engine.frame_stride(l_max)
        self.audio = torch.zeros(B, l_max, dtype=torch.float32, device=dev)
        self.sample_len = torch.full((B,), l_max, dtype=torch.int32, device=dev)
        self.norm_len = torch.full((B,), l_max, dtype=torch.int32, device=dev)
        self.out_len = torch.full((B,), engine.num_frames(l_max), dtype=torch.int32, device=dev)
        self.logits = torch.zeros(B, R, dtype=torch.float32, device=dev)
        self.probs = torch.zeros(B, R, dtype=torch.float32, device=dev)
        engine._workspace(B, l_max)                       # sized before capture: no allocation inside the graph
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up outside the capture (function attributes, tensor maps)
            self._forward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._forward()

    def _forward(self):
        self.engine.sfc_forward(self.audio, self.sample_len, self.norm_len, self.out_len, self.l_max,
                                logits_out=self.logits, probs_out=self.probs)

    def replay(self):
        self.graph.replay()
        return self.probs


def moving_average_device(arr, window: int, device=None):
    """numpy/torch fp64 vector -> trailing moving average on the GPU, returned as numpy fp64.
    Stand-alone (no model handle needed): lib/segment.py:508-522."""
    import numpy as np

    lib = nat.load()
    if device is None:   # the process's current GPU (under torchrun: LOCAL_RANK's device, not GPU 0)
        device = torch.device("cuda", torch.cuda.current_device())
    a = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(device)
    out = torch.empty_like(a)
    if a.numel():
        nat.check(lib.w2vseg_moving_average(a.data_ptr(), a.numel(), int(window), out.data_ptr(),
                                            torch.cuda.current_stream(a.device).cuda_stream),
                  "w2vseg_moving_average")
    return out.cpu().numpy()
