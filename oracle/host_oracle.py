"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/sfc_oracle.py header for the rules and pinning).

Plain-loop restatement of the HOST side of the SFC path: the window plan, collate, the scatter of
batch rows into the per-talk probability vector, NaN fill, tiling average, moving average and the
three segmentation algorithms that consume the probabilities. Written for obviousness, not speed;
pinned against the reference's own functions by oracle/make_golden.py -> tests/golden/*.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

INPUT_SR = 16_000            # lib/constants.py:1
TARGET_SR = 49.95            # lib/constants.py:2
FRAME_SEC = 20 / 1000        # lib/constants.py:3 (WAV2VEC_FRAME_LEN ms)


# ----------------------------------------------------------------------------- window plan
def to_outframes(x) -> int:
    """lib/dataset.py:604-606 — numpy round-half-even of samples * 49.95/16000"""
    return int(np.round(x * (1 / (INPUT_SR / TARGET_SR))).astype(int))


def window_plan(duration_samples: int, segment_sec: int, inference_times: int, i: int):
    """lib/dataset.py:612-639 fixed_length_segmentation(i) -> (starts, ends) in samples"""
    seg = int(np.round(segment_sec * INPUT_SR).astype(int))
    start = round(seg / inference_times * i)
    if start > duration_samples:
        start = 0
    cuts = list(range(start, duration_samples, seg))
    if cuts[0] != 0:
        cuts = [0] + cuts
    if cuts[-1] != duration_samples:
        two_sec = int(np.round(2 * INPUT_SR).astype(int))
        if duration_samples - cuts[-1] < two_sec:
            cuts[-1] = duration_samples
        else:
            cuts.append(duration_samples)
    return cuts[:-1], cuts[1:]


def window_frames(start_sample: int, end_sample: int):
    """lib/dataset.py:665-666"""
    return to_outframes(start_sample + 1e-6), to_outframes(end_sample + 1e-6)


# ----------------------------------------------------------------------------- collate
def collate(waves, starts_f, ends_f):
    """lib/datautils.py:61-142 for target-less batches. waves: list of 1-D float32 arrays.
    Returns dict(audio [B, Lmax] f32 normalised, in_len, out_mask [B, max(out_len)] bool,
    included, starts, ends)."""
    B = len(waves)
    lmax = max(len(w) for w in waves)
    audio = np.zeros((B, lmax), dtype=np.float32)
    included = []
    for b, w in enumerate(waves):
        audio[b, : len(w)] = w
        included.append(bool(np.float32(w.sum(dtype=np.float32) if len(w) else 0.0)))
    out_len = [e - s for s, e in zip(starts_f, ends_f)]
    out_mask = np.zeros((B, max(out_len)), dtype=bool)
    for b in range(B):
        out_mask[b, : out_len[b]] = True
    return {
        "audio_raw": audio,
        "in_len": [len(w) for w in waves],
        "out_mask": out_mask,
        "included": included,
        "starts": list(starts_f),
        "ends": list(ends_f),
    }


# ----------------------------------------------------------------------------- talk vector
def scatter_batch(talk: np.ndarray, probs: np.ndarray, starts, ends, included, ends_shift: int):
    """lib/evaluate.py:100-111 (+ the `ends -= 1` of :68 passed as ends_shift)"""
    for i in range(len(probs)):
        start, end = starts[i], ends[i] - ends_shift
        if included[i] and end > start:
            talk[start:end] = probs[i, : end - start]
        elif not included[i]:
            talk[start:end] = 0


def nan_fill(talk: np.ndarray):
    """lib/evaluate.py:118-125 — sequential, in place"""
    n = len(talk)
    for j in np.where(np.isnan(talk))[0]:
        talk[j] = np.nanmean(talk[max(0, j - 2): min(n, j + 3)])


def average_tilings(per_tiling):
    """segment.py:101-108"""
    acc = per_tiling[0].copy()
    for p in per_tiling[1:]:
        acc += p
    acc /= len(per_tiling)
    return acc


def moving_average(arr: np.ndarray, window: int) -> np.ndarray:
    """lib/segment.py:508-522 — trailing mean with ramp-up, Python left-to-right float64 sum"""
    out = np.empty(len(arr))
    for i in range(len(arr)):
        lo = max(0, i - window + 1)
        s = 0
        for k in range(lo, i + 1):
            s = s + arr[k]
        out[i] = s / (i + 1 - lo)
    return out


# ----------------------------------------------------------------------------- segments
@dataclass
class Seg:
    """lib/segment.py:13-31 — start/end in frames (may be fractional), 6-decimal seconds"""
    start: float
    end: float

    @property
    def duration(self):
        return float(round((self.end - self.start) / TARGET_SR, 6))

    @property
    def offset(self):
        return float(round(self.start / TARGET_SR, 6))


def _trim(probs, a, b, thr):
    """lib/segment.py:34-53 on the half-open frame range [a, b)"""
    idx = np.where(probs[a:b] >= thr)[0]
    if len(idx) == 0:
        return a, a
    return a + int(idx[0]), a + int(idx[-1]) + 1


def pdac(probs, max_segment_length=18, min_segment_length=0.2, threshold=0.5):
    """lib/segment.py:186-235 (+ split_and_trim :113-134)"""
    out = []

    def dur(a, b):
        return Seg(a, b).duration

    def rec(a, b):
        if dur(a, b) < max_segment_length:
            out.append(Seg(a, b))
            return
        order = np.argsort(probs[a:b])
        for j in order:
            if probs[a + j] > threshold:
                out.append(Seg(a, b))
                return
            la, lb = _trim(probs, a, a + int(j), threshold)
            ra, rb = _trim(probs, a + int(j) + 1, b, threshold)
            if dur(la, lb) > min_segment_length and dur(ra, rb) > min_segment_length:
                rec(la, lb)
                rec(ra, rb)
                return
        out.append(Seg(a, b))

    a0, b0 = _trim(probs, 0, len(probs), threshold)
    rec(a0, b0)
    return out


def _split_strm(preds: str, max_len: int, min_len: int, min_pause: int):
    """lib/segment.py:454-505"""
    total = len(preds)
    start, leftover, pieces = 0, "", []
    while start < total:
        end = min(start + max_len - len(leftover), total)
        cur = leftover + preds[start:end]
        first, second = cur[:min_len], cur[min_len:]
        runs = re.findall(r"0{1,}", second)
        best = ""
        for r in runs:  # stable sort by length, take last == last of the longest runs
            if len(r) >= len(best):
                best = r
        if len(best) > min_pause:
            head_b, leftover = second.split(best, maxsplit=1)
            if set(first) == set("0") or first == "":
                pieces.append(first)
                if len(head_b):
                    pieces.append(head_b)
            else:
                pieces.append(first + head_b)
            pieces.append(best)
        else:
            pieces.append(cur)
            leftover = ""
        start = end
    return pieces


def strm(probs, max_segment_length=18, min_segment_length=0.2, min_pause_length=0.2, threshold=0.5):
    """lib/segment.py:419-443 + get_segments :389-416"""
    preds = "".join("1" if p > threshold else "0" for p in probs)
    pieces = _split_strm(preds, int(max_segment_length / FRAME_SEC), int(min_segment_length / FRAME_SEC),
                         int(min_pause_length / FRAME_SEC))
    total = len("".join(pieces))
    pad = TARGET_SR * 0.06
    out, off = [], 0
    for piece in pieces:
        if not (set(piece) == set("0") or piece == ""):
            out.append(Seg(max(0, off - pad), min(off + len(piece) + pad, total)))
        off += len(piece)
    return out


def pthr(probs, max_segment_length=18, min_segment_length=0.2, max_lerp_range=0, min_lerp_range=0,
         threshold=0.5, moving_average_window=0):
    """lib/segment.py:525-592"""
    max_steps = int(max_segment_length / FRAME_SEC)
    min_steps = int(min_segment_length / FRAME_SEC)
    max_lerp = int(max_lerp_range / FRAME_SEC)
    min_lerp = int(min_lerp_range / FRAME_SEC)
    thr = np.full((max_steps), threshold)
    thr[:min_steps] = 0
    thr[min_steps: min_steps + min_lerp] = np.arange(min_lerp, dtype=float) / (min_lerp / threshold)
    thr[max_steps - max_lerp: max_steps] = threshold + np.arange(max_lerp, dtype=float) / (max_lerp / threshold)
    if moving_average_window > 0:
        probs = moving_average(probs, int(moving_average_window / FRAME_SEC))
    n = len(probs)
    pad = TARGET_SR * 0.06
    out, start = [], 0
    while start < n:
        if probs[start] <= threshold:
            start += 1
            continue
        end = None
        for k in range(min(len(thr), n - start)):
            if probs[start + k] <= thr[k]:
                end = start + k
                break
        if end is None:
            end = min(start + len(thr), n - 1)
        out.append(Seg(max(0, start - pad), min(end + pad, n - 1)))
        start = end + 1
    return out


def yaml_records(segments, wav_name):
    """lib/segment.py:595-618"""
    return [{"duration": s.duration, "offset": s.offset, "rW": 0, "uW": 0, "speaker_id": "NA",
             "wav": wav_name} for s in segments]
