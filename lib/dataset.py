"""Fixed-length windowing of one wav for SFC inference (drop-in for the reference's
lib.dataset.FixedSegmentationDatasetNoTarget, lib/dataset.py:571-668).

Differences in mechanism, not in results: the wav is decoded ONCE with the stdlib `wave` module
(16-bit PCM -> float32 / 32768, what the reference's sox loader yields) and windows are views of
that array; the window plan comes from wav2vecsegmenter_b200.pipeline (bit-exact integer logic).
"""
from __future__ import annotations

import os
import sys
import wave
from typing import Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

sys.path.append(os.path.dirname(__file__))  # the reference makes `from constants import ...` work this way

from constants import INPUT_SAMPLE_RATE, TARGET_SAMPLE_RATE  # noqa: E402
from wav2vecsegmenter_b200 import pipeline  # noqa: E402


def read_wav(path) -> tuple[np.ndarray, int]:
    """mono 16-bit PCM wav -> (float32 samples in [-1, 1), sample rate)"""
    with wave.open(str(path), "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM wav is supported (sample width {w.getsampwidth()})")
        n, ch, sr = w.getnframes(), w.getnchannels(), w.getframerate()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2")
    if ch > 1:
        pcm = pcm.reshape(-1, ch)[:, 0]
    return pcm.astype(np.float32) / np.float32(32768.0), sr


class FixedSegmentationDatasetNoTarget(Dataset):
    def __init__(self, path_to_wav: str, segment_length: int = 20, inference_times: int = 1) -> None:
        super().__init__()
        self.input_sr = INPUT_SAMPLE_RATE
        self.target_sr = TARGET_SAMPLE_RATE
        self.in_trg_ratio = self.input_sr / self.target_sr
        self.trg_in_ratio = 1 / self.in_trg_ratio
        self.segment_length = segment_length
        self.segment_length_inframes = int(np.round(segment_length * self.input_sr).astype(int))
        self.inference_times = inference_times
        self.path_to_wav = path_to_wav
        self.wave, self.sample_rate = read_wav(path_to_wav)
        assert self.sample_rate == self.input_sr, f"Audio needs to have sample rate of {self.input_sr}"
        self.duration_inframes = len(self.wave)
        self.duration_outframes = pipeline.samples_to_frames(self.duration_inframes)
        self.starts, self.ends = [], []

    def fixed_length_segmentation(self, i: int) -> None:
        """tiling i of inference_times (0 <= i < inference_times)"""
        s, e = pipeline.tiling_bounds(self.duration_inframes, self.segment_length, self.inference_times, i)
        self.starts, self.ends = np.asarray(s), np.asarray(e)

    def __len__(self) -> int:
        return len(self.starts)

    def __getitem__(self, index: int) -> Tuple[torch.FloatTensor, None, int, int]:
        a, b = int(self.starts[index]), int(self.ends[index])
        return (torch.from_numpy(self.wave[a:b]), None,
                pipeline.samples_to_frames(a + 1e-6), pipeline.samples_to_frames(b + 1e-6))
