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


# ---------------------------------------------------------------------------------------------
# Dev-set scoring path (reference lib/dataset.py:19-146, 335-498, 737-813): fixed-length windows of
# LABELLED talks with per-frame targets. Same tsv inputs, same window plan, same target strings and
# tensors; each wav is decoded once and windows are views of it.
# ---------------------------------------------------------------------------------------------
import pandas as pd  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from datautils import CollateFn  # noqa: E402


class SegmentationDataset(Dataset):
    """talk / true-segment tables and the target construction shared by the labelled datasets
    (reference lib/dataset.py:19-146; only the binary `bce` labels, vocab must be None)"""

    def __init__(self, talk_list: str, segments_list: str, vocab=None) -> None:
        super().__init__()
        if vocab is not None:
            raise NotImplementedError("token-level (ce / ssl) targets are outside the accelerated path")
        self.input_sr = INPUT_SAMPLE_RATE
        self.target_sr = TARGET_SAMPLE_RATE
        self.in_trg_ratio = self.input_sr / self.target_sr
        self.trg_in_ratio = 1 / self.in_trg_ratio
        self.talks_df = pd.read_csv(talk_list, sep="\t", index_col=0)
        self.segments_df = pd.read_csv(segments_list, sep="\t", index_col=0)
        self.vocab = vocab
        self.columns = ["talk_id", "start", "end", "duration", "included"]
        self.n_pos, self.n_all = 0, 0     # running counts behind pos_class_percentage
        self._labels = {}                 # talk_id -> binary label per input sample
        self._waves = {}                  # wav path -> decoded samples

    def _secs_to_outframes(self, x):
        return np.round(x * self.target_sr).astype(int)

    def _outframes_to_inframes(self, x):
        return np.round(x * self.in_trg_ratio).astype(int)

    def _inframes_to_outframes(self, x):
        return np.round(x * self.trg_in_ratio).astype(int)

    def _secs_to_inframes(self, x):
        return np.round(x * self.input_sr).astype(int)

    def _talk_row(self, talk_id):
        return self.talks_df.loc[self.talks_df["id"] == talk_id].iloc[0]

    def _talk_labels(self, talk_id) -> np.ndarray:
        if talk_id not in self._labels:
            lab = np.zeros(int(self._talk_row(talk_id)["total_frames"]))
            true = self.segments_df.loc[self.segments_df.talk_id == talk_id]
            for a, b in zip(true.start.to_numpy(), true.end.to_numpy()):
                lab[int(a): int(b)] = 1
            self._labels[talk_id] = lab
        return self._labels[talk_id]

    def _get_targets_for_segment(self, true_points: np.ndarray) -> list[list[int]]:
        """runs of ones in the window's input-space labels -> [start, end) pairs in output frames; a
        run that would start on the previous run's last frame is moved one frame later"""
        n = len(true_points)
        change = np.flatnonzero(true_points[1:] != true_points[:-1]) + 1
        bounds = np.concatenate(([0], change, [n])) if n else np.array([0, 0])
        targets: list[list[int]] = []
        for s, e in zip(bounds[:-1], bounds[1:]):
            if n and true_points[s] == 1:
                s, e = int(self._inframes_to_outframes(s)), int(self._inframes_to_outframes(e))
                if targets and s <= targets[-1][-1]:
                    s += 1
                targets.append([s, e])
                self.n_pos += e - s
        self.n_all += int(self._inframes_to_outframes(n))
        return targets

    def _get_targets_for_talk(self, sgm_df: pd.DataFrame, talk_id: str) -> pd.DataFrame:
        lab = self._talk_labels(talk_id)
        inc = []
        for a, b in zip(sgm_df.start.to_numpy(), sgm_df.end.to_numpy()):
            t = self._get_targets_for_segment(lab[int(a): int(b)])
            inc.append(",".join(f"{s}:{e}" for s, e in t) if t else "NA")
        sgm_df["included"] = inc
        return sgm_df

    def _construct_target(self, segment) -> torch.FloatTensor:
        target_len = int(self._inframes_to_outframes(segment.duration))
        target = torch.zeros(target_len, dtype=torch.float)
        if segment.included != "NA":
            for s_e in segment.included.split(","):
                s, e = s_e.split(":")
                target[int(s): min(int(e), target_len + 1)] = 1
        return target

    def _wave(self, path) -> np.ndarray:
        if path not in self._waves:
            x, sr = read_wav(path)
            assert sr == self.input_sr, f"Audio needs to have sample rate of {self.input_sr}"
            self._waves[path] = x
        return self._waves[path]


class FixedSegmentationDataset(SegmentationDataset):
    """fixed-length windows of labelled talks (reference lib/dataset.py:335-498)"""

    def __init__(self, talk_list, segments_list, segment_length, inference_times, vocab=None) -> None:
        super().__init__(talk_list, segments_list, vocab)
        self.segment_length = segment_length
        self.segment_length_inframes = int(self._secs_to_inframes(segment_length))
        self.inference_times = inference_times
        self.fixed_segments_df = pd.DataFrame(columns=self.columns)

    def _plan(self, talk_id: str, i: int) -> pd.DataFrame:
        row = self._talk_row(talk_id)
        self.talk_path = row["path"]
        self.duration_inframes = int(row["total_frames"])
        self.duration_outframes = int(self._inframes_to_outframes(self.duration_inframes))
        starts, ends = pipeline.tiling_bounds(self.duration_inframes, self.segment_length, self.inference_times, i)
        df = pd.DataFrame({"talk_id": talk_id, "start": starts, "end": ends}, columns=self.columns)
        df["duration"] = df.end - df.start
        return self._get_targets_for_talk(df, talk_id)

    def generate_fixed_segments(self, talk_id: str, i: int) -> None:
        """tiling i (0 <= i < inference_times) of one talk"""
        self.fixed_segments_df = self._plan(talk_id, i)

    def generate_fixed_segments_all_talks(self, i: int) -> None:
        self.fixed_segments_df = pd.concat([self._plan(t, i) for t in self.talks_df["id"]], ignore_index=True)
        self.pos_class_percentage = self.n_pos / self.n_all

    def __len__(self) -> int:
        return len(self.fixed_segments_df)

    def __getitem__(self, index: int) -> Tuple[torch.FloatTensor, torch.FloatTensor, int, int]:
        seg = self.fixed_segments_df.iloc[index]
        if not pd.isna(seg.talk_id):
            self.talk_path = self._talk_row(seg.talk_id)["path"]
        a, b = int(seg.start), int(seg.end)
        wav = torch.from_numpy(self._wave(self.talk_path)[a:b])
        return (wav, self._construct_target(seg), int(self._inframes_to_outframes(a + 1e-6)),
                int(self._inframes_to_outframes(b + 1e-6)))


class FixedDataloaderGenerator:
    """dataloaders over the fixed-length tilings of a labelled wav collection
    (reference lib/dataset.py:737-813)"""

    def __init__(self, talk_list, segments_list, segment_length, batch_size, num_workers,
                 inference_times: int = 1, autoregression: bool = False, vocab=None) -> None:
        if autoregression or vocab:
            raise NotImplementedError("only the binary (bce) frame classifier is on the accelerated path")
        self.talk_list = talk_list
        self.segments_list = segments_list
        self.segment_length = segment_length
        self.batch_size = batch_size
        self.num_workers = num_workers
        self.autoregression = autoregression
        self.vocab = vocab
        self.collate_fn = CollateFn(pad_token_id=0)
        self.dataset = FixedSegmentationDataset(talk_list, segments_list, segment_length, inference_times, vocab)

    def generate(self, talk_id: str, i: int) -> DataLoader:
        if talk_id == "":
            self.dataset.generate_fixed_segments_all_talks(i)
        else:
            self.dataset.generate_fixed_segments(talk_id, i)
        # the wav is already decoded in this process: worker processes would only copy it around
        return DataLoader(self.dataset, batch_size=self.batch_size, num_workers=0, drop_last=False,
                          shuffle=False, collate_fn=self.collate_fn)

    def get_talk_ids(self) -> list:
        return self.dataset.talks_df["id"].tolist()
