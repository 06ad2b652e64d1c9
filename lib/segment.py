"""Probability -> segments: pDAC, pSTRM, pTHR, the moving average and the yaml records.

Same function names, arguments and results as the reference's lib/segment.py (cited per
function); the implementations are new: O(1) trims from precomputed next/previous-above-threshold
tables and an explicit stack for pDAC, byte-string run scanning for pSTRM, index jumps for pTHR,
and the moving average on the GPU (w2vseg_moving_average, bit-identical fp64 left-to-right sums).
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

try:  # bare import like the reference (lib/ on sys.path) or package import
    from constants import TARGET_SAMPLE_RATE, WAV2VEC_FRAME_LEN
except ImportError:  # pragma: no cover
    from lib.constants import TARGET_SAMPLE_RATE, WAV2VEC_FRAME_LEN

_PAD_FRAMES = TARGET_SAMPLE_RATE * 0.06   # segments are widened by 0.06 s on both sides


@dataclass
class Segment:
    """frames [start, end) of a talk; seconds are frames / 49.95 rounded to `decimal` places
    (reference lib/segment.py:13-31)"""

    start: float
    end: float
    probs: np.ndarray = None
    logits: np.ndarray = None
    decimal: int = 6

    @property
    def duration(self) -> float:
        return float(round((self.end - self.start) / TARGET_SAMPLE_RATE, self.decimal))

    @property
    def offset(self) -> float:
        return float(round(self.start / TARGET_SAMPLE_RATE, self.decimal))

    @property
    def offset_plus_duration(self) -> float:
        return round(self.offset + self.duration, self.decimal)


def _seconds(a, b) -> float:
    return float(round((b - a) / TARGET_SAMPLE_RATE, 6))


def trim(sgm: Segment, threshold: float) -> Segment:
    """shrink to the first/last frame with p >= threshold (reference lib/segment.py:34-53)"""
    keep = np.flatnonzero(sgm.probs >= threshold)
    if keep.size == 0:
        return Segment(sgm.start, sgm.start, probs=np.empty([0]))
    lo, hi = int(keep[0]), int(keep[-1]) + 1
    return Segment(sgm.start + lo, sgm.start + hi, probs=sgm.probs[lo:hi])


def split_and_trim(sgm: Segment, split_idx: int, threshold: float):
    """drop frame split_idx, trim both sides (reference lib/segment.py:113-134)"""
    left = Segment(sgm.start, sgm.start + split_idx, sgm.probs[:split_idx])
    right = Segment(left.end + 1, sgm.end, sgm.probs[split_idx + 1:])
    return trim(left, threshold), trim(right, threshold)


class _TrimTable:
    """next / previous frame with p >= threshold for every position: trim of any [a, b) in O(1)"""

    def __init__(self, probs: np.ndarray, threshold: float):
        n = len(probs)
        ok = probs >= threshold
        idx = np.arange(n)
        nxt = np.where(ok, idx, n)
        self.nxt = np.minimum.accumulate(nxt[::-1])[::-1] if n else nxt
        prv = np.where(ok, idx, -1)
        self.prv = np.maximum.accumulate(prv) if n else prv

    def trim(self, a: int, b: int):
        if b <= a:
            return a, a
        lo = int(self.nxt[a])
        if lo >= b:
            return a, a
        return lo, int(self.prv[b - 1]) + 1


def pdac(probs: np.ndarray, max_segment_length: float = 18, min_segment_length: float = 0.2,
         threshold: float = 0.5) -> list[Segment]:
    """probabilistic divide-and-conquer (reference lib/segment.py:186-235): split a too-long
    segment at its lowest-probability frame whose two trimmed halves are both longer than
    min_segment_length; candidates are tried in ascending probability (np.argsort order, like
    the reference) and the search stops at the first candidate above the threshold."""
    probs = np.asarray(probs)
    table = _TrimTable(probs, threshold)
    out: list[Segment] = []
    a0, b0 = table.trim(0, len(probs))
    stack = [(a0, b0)]
    while stack:
        a, b = stack.pop()
        if _seconds(a, b) < max_segment_length:
            out.append(Segment(a, b, probs=probs[a:b]))
            continue
        seg = probs[a:b]
        done = False
        for j in np.argsort(seg):
            j = int(j)
            if seg[j] > threshold:
                break
            la, lb = table.trim(a, a + j)
            ra, rb = table.trim(a + j + 1, b)
            if _seconds(la, lb) > min_segment_length and _seconds(ra, rb) > min_segment_length:
                stack.append((ra, rb))   # right half is emitted after the whole left subtree
                stack.append((la, lb))
                done = True
                break
        if not done:
            out.append(Segment(a, b, probs=probs[a:b]))
    return out


def pdac_with_logits(*args, **kwargs):
    raise NotImplementedError("pdac_with_logits belongs to the reference's ce/ssl variants, which are "
                              "outside the accelerated SFC path")


# ----------------------------------------------------------------------------------- pSTRM
def is_pause(x: str) -> bool:
    return x == "" or x.count("0") == len(x)


def get_pauses(pred: str) -> list[str]:
    return re.findall(r"0+", pred)


def split_predictions_strm(preds: str, max_segm_len: int, min_segm_len: int, min_pause_len: int) -> list[str]:
    """streaming split of Gaido et al. 2021 (reference lib/segment.py:454-505): look at up to
    max_segm_len frames, protect the first min_segm_len, cut at the first longest pause of the rest
    if it is longer than min_pause_len, carry what follows the pause over to the next look."""
    total = len(preds)
    pieces, carry, pos = [], "", 0
    while pos < total:
        nxt = min(pos + max_segm_len - len(carry), total)
        window = carry + preds[pos:nxt]
        head, tail = window[:min_segm_len], window[min_segm_len:]
        longest = ""
        for m in re.finditer(r"0+", tail):
            if m.end() - m.start() >= len(longest):   # reference: stable sort by length, take last
                longest = m.group()
        if len(longest) > min_pause_len:
            cut = tail.find(longest)                   # == split(longest, maxsplit=1)
            before, carry = tail[:cut], tail[cut + len(longest):]
            if is_pause(head):
                pieces.append(head)
                if before:
                    pieces.append(before)
            else:
                pieces.append(head + before)
            pieces.append(longest)
        else:
            pieces.append(window)
            carry = ""
        pos = nxt
    return pieces


def get_segments(splitted_predictions: list[str], frame_length: float) -> list[Segment]:
    """speech pieces -> segments widened by 0.06 s (reference lib/segment.py:389-416)"""
    total = sum(len(p) for p in splitted_predictions)
    out, off = [], 0
    for piece in splitted_predictions:
        if not is_pause(piece):
            out.append(Segment(max(0, off - _PAD_FRAMES), min(off + len(piece) + _PAD_FRAMES, total)))
        off += len(piece)
    return out


def strm(probs: np.ndarray, max_segment_length: float = 18, min_segment_length: float = 0.2,
         min_pause_length: float = 0.2, threshold: float = 0.5) -> list[Segment]:
    """reference lib/segment.py:419-443"""
    frame = WAV2VEC_FRAME_LEN / 1000
    bits = (np.asarray(probs) > threshold).astype(np.uint8) + ord("0")
    preds = bits.tobytes().decode("ascii")
    pieces = split_predictions_strm(preds, int(max_segment_length / frame), int(min_segment_length / frame),
                                    int(min_pause_length / frame))
    return get_segments(pieces, frame)


# ----------------------------------------------------------------------------------- pTHR
def moving_average(arr: np.ndarray, window: int) -> np.ndarray:
    """trailing mean with ramp-up (reference lib/segment.py:508-522), computed by the CUDA kernel
    w2vseg_moving_average: every output is the same left-to-right fp64 sum the reference's Python
    `sum(part) / len(part)` produces, so results are bit-identical. Needs a GPU (no CPU fallback)."""
    from wav2vecsegmenter_b200.engine import moving_average_device

    arr = np.asarray(arr, dtype=np.float64)
    if window <= 0:
        raise ZeroDivisionError("moving_average window must be >= 1 frame")
    return moving_average_device(arr, int(window))


def pthr(probs: np.ndarray, max_segment_length: float = 18, min_segment_length: float = 0.2,
         max_lerp_range: float = 0, min_lerp_range: float = 0, threshold: float = 0.5,
         moving_average_window: float = 0) -> list[Segment]:
    """threshold segmentation with a length-dependent end threshold (reference
    lib/segment.py:525-592): a segment opens at the first frame with p > threshold and closes at
    the first frame k steps later with p <= curve[k]; curve is 0 for the first min steps, ramps up
    to `threshold`, stays there, and ramps to 2*threshold over the last max_lerp steps."""
    frame = WAV2VEC_FRAME_LEN / 1000
    n_max = int(max_segment_length / frame)
    n_min = int(min_segment_length / frame)
    n_up = int(max_lerp_range / frame)
    n_lo = int(min_lerp_range / frame)
    curve = np.full((n_max), threshold)
    curve[:n_min] = 0
    if n_lo:
        curve[n_min: n_min + n_lo] = np.arange(n_lo, dtype=float) / (n_lo / threshold)
    if n_up:
        curve[n_max - n_up: n_max] = threshold + np.arange(n_up, dtype=float) / (n_up / threshold)
    probs = np.asarray(probs)
    if moving_average_window > 0:
        probs = moving_average(probs, int(moving_average_window / frame))
    n = len(probs)
    opens = np.flatnonzero(probs > threshold)
    out, pos, k = [], 0, 0
    while k < len(opens):
        start = int(opens[k])
        if start < pos:
            k = int(np.searchsorted(opens, pos))
            continue
        part = probs[start: start + len(curve)]
        hits = np.flatnonzero(part <= curve[: len(part)])
        end = start + int(hits[0]) if hits.size else min(start + len(curve), n - 1)
        out.append(Segment(max(0, start - _PAD_FRAMES), min(end + _PAD_FRAMES, n - 1)))
        pos = end + 1
        k += 1
    return out


def update_yaml_content(yaml_content: list[dict], segments: list[Segment], wav_name: str) -> list[dict]:
    """append MuST-C style records (reference lib/segment.py:595-618)"""
    yaml_content.extend(
        {"duration": s.duration, "offset": s.offset, "rW": 0, "uW": 0, "speaker_id": "NA", "wav": wav_name}
        for s in segments
    )
    return yaml_content
