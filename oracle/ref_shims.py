"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Makes the UNMODIFIED reference (/root/reference, read-only, only present in the build container)
importable on CPU so that its own code can (a) pin the oracle and (b) generate the golden vectors
under tests/golden/ (see oracle/make_golden.py). Five shims, exactly those SURVEY.md §8c lists:

  1. stub `hydra` / `hydra.utils.instantiate`      (imported, unused: lib/models.py:5)
  2. Wav2Vec2Model.from_pretrained -> random-init XLS-R-300m architecture (lib/models.py:334,443)
  3. Wav2Vec2Processor.from_pretrained -> dummy    (lib/datautils.py:7-9 downloads at import)
  4. torchaudio.info / torchaudio.backend.sox_io_backend.load re-created on top of `wave` and
     scipy (lib/dataset.py:596-598, 659-663; both APIs are gone in torchaudio 2.11)
  5. np.int = int                                   (lib/segment.py:431; removed in numpy 2)
  6. int(pandas.Series of length 1)                 (lib/dataset.py:372,414 on the dev-set scoring
     path; removed in pandas 3 — restored as Series.__int__ = int(self.iloc[0]))
"""
from __future__ import annotations

import sys
import types
import wave
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "lib" / "models.py").exists()


GROUP_NORM_NAME = "synthetic/xls-r-300m-group-norm"          # feat_extract_norm="group", conv_bias=True
GROUP_NORM_NOBIAS_NAME = "synthetic/xls-r-300m-group-norm-nobias"
POST_LN_NAME = "synthetic/wav2vec2-large-960h-like"         # group norm, no conv bias, post-LN encoder layers


def xlsr_config(feat_extract_norm="layer", conv_bias=True, stable_ln=True):
    from transformers import Wav2Vec2Config

    return Wav2Vec2Config(
        hidden_size=1024,
        num_hidden_layers=24,
        num_attention_heads=16,
        intermediate_size=4096,
        hidden_act="gelu",
        layer_norm_eps=1e-5,
        feat_extract_norm=feat_extract_norm,
        feat_extract_activation="gelu",
        conv_bias=conv_bias,
        conv_dim=(512,) * 7,
        conv_kernel=(10, 3, 3, 3, 3, 2, 2),
        conv_stride=(5, 2, 2, 2, 2, 2, 2),
        num_conv_pos_embeddings=128,
        num_conv_pos_embedding_groups=16,
        do_stable_layer_norm=stable_ln,
        mask_time_prob=0.075,
    )


_installed = False


def install():
    """idempotent; returns the reference's modules as a namespace"""
    global _installed
    if not reference_available():
        raise RuntimeError("/root/reference is not present (it only exists in the build container)")
    if not _installed:
        for p in (str(REFERENCE_ROOT / "lib"), str(REFERENCE_ROOT)):
            if p not in sys.path:
                sys.path.insert(0, p)
        # the repo has its own top-level `lib` package (the drop-in); make sure the reference's wins
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.") or k == "constants"]:
            del sys.modules[k]
        if not hasattr(np, "int"):
            np.int = int  # noqa: NPY001
        import pandas as pd

        try:
            int(pd.Series([3]))
        except TypeError:
            pd.Series.__int__ = lambda self: int(self.iloc[0])
        h = types.ModuleType("hydra")
        hu = types.ModuleType("hydra.utils")
        hu.instantiate = lambda *a, **k: None
        h.utils = hu
        sys.modules.setdefault("hydra", h)
        sys.modules.setdefault("hydra.utils", hu)

        from transformers import Wav2Vec2Model, Wav2Vec2Processor

        def _cfg(name):   # the architecture is chosen by the (fake) model name SHAS passes through
            if name == GROUP_NORM_NAME:
                return xlsr_config("group", True)
            if name == GROUP_NORM_NOBIAS_NAME:
                return xlsr_config("group", False)
            if name == POST_LN_NAME:
                return xlsr_config("group", False, stable_ln=False)
            return xlsr_config()

        Wav2Vec2Model.from_pretrained = classmethod(lambda cls, name, *a, **k: cls(_cfg(name)))
        Wav2Vec2Processor.from_pretrained = classmethod(
            lambda cls, *a, **k: types.SimpleNamespace(
                tokenizer=types.SimpleNamespace(get_vocab=lambda: {})
            )
        )

        import torch
        import torchaudio

        def _info(path):
            with wave.open(str(path), "rb") as w:
                return types.SimpleNamespace(num_frames=w.getnframes(), sample_rate=w.getframerate())

        def _load(path, frame_offset=0, num_frames=-1):
            from scipy.io import wavfile

            sr, data = wavfile.read(str(path), mmap=True)
            end = None if num_frames < 0 else frame_offset + num_frames
            seg = np.asarray(data[frame_offset:end])
            if seg.dtype == np.int16:
                seg = seg.astype(np.float32) / 32768.0
            return torch.from_numpy(seg.astype(np.float32))[None], sr

        torchaudio.info = _info
        backend = types.ModuleType("torchaudio.backend")
        sox = types.ModuleType("torchaudio.backend.sox_io_backend")
        sox.load = _load
        backend.sox_io_backend = sox
        torchaudio.backend = backend
        sys.modules["torchaudio.backend"] = backend
        sys.modules["torchaudio.backend.sox_io_backend"] = sox
        _installed = True

    import importlib

    ns = types.SimpleNamespace()
    ns.models = importlib.import_module("lib.models")
    ns.dataset = importlib.import_module("lib.dataset")
    ns.datautils = importlib.import_module("lib.datautils")
    ns.evaluate = importlib.import_module("lib.evaluate")
    ns.segment = importlib.import_module("lib.segment")
    assert str(REFERENCE_ROOT) in ns.models.__file__, ns.models.__file__
    return ns
