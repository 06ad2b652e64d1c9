"""Host logic of the SFC path around the CUDA forward: the window plan of a talk, the per-window
metadata that reproduces the reference's batch-dependent behaviour, device batching, sharding of
windows across ranks, and the talk-level reductions.

Reference behaviour reproduced (all host-side integer logic, bit-exact):
  * fixed-length tiling i of a talk                       lib/dataset.py:612-639
  * sample -> frame indices (numpy round-half-even)       lib/dataset.py:604-606, 665-666
  * CollateFn: `included`, padded length of the batch     lib/datautils.py:88, 98-103, 122-125
  * the +-1 frame fix-up incl. `ends -= 1` for the batch  lib/evaluate.py:63-70
  * scatter into the talk vector, NaN fill                lib/evaluate.py:100-125
  * average over tilings                                  segment.py:101-108
The reference batches `batch_size` CONSECUTIVE windows of one tiling; two things depend on that
grouping (the normalisation length and the fix-up), so each window carries them as metadata and
the device is free to batch / shard windows any other way.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

INPUT_SR = 16_000
TARGET_SR = 49.95
_CONV_K = (10, 3, 3, 3, 3, 2, 2)
_CONV_S = (5, 2, 2, 2, 2, 2, 2)


def num_frames(n: int) -> int:
    """conv-stack output length (same arithmetic as w2vseg_num_frames / HF:1005-1024)"""
    for k, s in zip(_CONV_K, _CONV_S):
        if n < k:
            return 0
        n = (n - k) // s + 1
    return int(n)


def samples_to_frames(x) -> int:
    return int(np.round(x * (1 / (INPUT_SR / TARGET_SR))).astype(int))


@dataclass
class Window:
    talk: int            # talk index
    tiling: int          # inference iteration i
    start: int           # samples
    end: int
    start_f: int         # frames (output space)
    end_f: int
    norm_len: int = 0    # padded length of the reference batch (0 = silent window, not normalised)
    out_len: int = 0     # leading true entries of the (possibly trimmed) out_mask row
    ends_shift: int = 0  # 1 if the reference decrements `ends` for this window's batch
    included: bool = True

    @property
    def n_samples(self) -> int:
        return self.end - self.start


def tiling_bounds(duration: int, segment_sec: float, inference_times: int, i: int):
    seg = int(np.round(segment_sec * INPUT_SR).astype(int))
    start = round(seg / inference_times * i)
    if start > duration:
        start = 0
    cuts = np.arange(start, duration, seg).astype(int)
    if cuts[0] != 0:
        cuts = np.insert(cuts, 0, 0)
    if cuts[-1] != duration:
        if duration - cuts[-1] < int(np.round(2 * INPUT_SR).astype(int)):
            cuts[-1] = duration
        else:
            cuts = np.append(cuts, duration)
    return [int(c) for c in cuts[:-1]], [int(c) for c in cuts[1:]]


def plan_tiling(duration: int, segment_sec: float, inference_times: int, i: int, batch_size: int,
                talk: int = 0, included=None) -> list[Window]:
    """windows of tiling i with the reference-batch metadata filled in. `included[k]` (optional)
    = whether window k has a non-zero sample sum (lib/datautils.py:88)."""
    starts, ends = tiling_bounds(duration, segment_sec, inference_times, i)
    wins = [Window(talk, i, s, e, samples_to_frames(s + 1e-6), samples_to_frames(e + 1e-6))
            for s, e in zip(starts, ends)]
    if included is not None:
        for w, inc in zip(wins, included):
            w.included = bool(inc)
    for b0 in range(0, len(wins), batch_size):
        group = wins[b0: b0 + batch_size]
        lmax = max(w.n_samples for w in group)
        t_hidden = num_frames(lmax)
        t_mask = max(w.end_f - w.start_f for w in group)
        shift = 0
        if t_hidden < t_mask:
            if t_mask - t_hidden != 1:
                raise ValueError("reference cannot run this batch: out_mask is more than one frame "
                                 f"longer than the encoder output ({t_mask} vs {t_hidden})")
            shift = 1
        elif t_hidden - t_mask > 1:
            raise ValueError("reference cannot run this batch: encoder output is more than one frame "
                             f"longer than out_mask ({t_hidden} vs {t_mask})")
        for w in group:
            w.norm_len = lmax if w.included else 0
            w.out_len = min(w.end_f - w.start_f, t_mask - shift)
            w.ends_shift = shift
    return wins


def scatter_plan(wins: list[Window], n_frames: int):
    """(start[], count[], nan_idx[]) for w2vseg_scatter_rows / w2vseg_nanfill: what
    lib/evaluate.py:100-111 writes, in window order, and which frames stay NaN afterwards."""
    start, count = [], []
    covered = np.zeros(n_frames, dtype=bool)
    for w in wins:
        s, e = w.start_f, w.end_f - w.ends_shift
        s_c, e_c = max(0, min(s, n_frames)), max(0, min(e, n_frames))
        if w.included and e > s:
            start.append(s)
            count.append(e - s)
            covered[s_c:e_c] = True
        elif not w.included:
            start.append(s)
            count.append(-(e - s) if e > s else 0)
            covered[s_c:e_c] = True
        else:
            start.append(s)
            count.append(0)
    return (np.asarray(start, dtype=np.int32), np.asarray(count, dtype=np.int32),
            np.flatnonzero(~covered).astype(np.int32))


def shard_ranges(n: int, world: int):
    """contiguous, balanced shards: rank r owns [lo, hi)"""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class LazyWave:
    """Device copy of one talk's samples, uploaded in chunks on a side stream just AHEAD of the device
    batches that read them, so that the host->device copy of a long talk overlaps its own forward
    passes (the reference copies batch by batch with blocking `.to(device)`, lib/evaluate.py:36-44;
    a single up-front copy of a 2 h talk is 460 MB). `lo` = first sample this rank needs."""

    # samples per copy (17.9 MB = one 14-window batch); W2VSEG_UPLOAD_CHUNK overrides it for A/B
    # measurements (a huge value = the whole talk in one copy before the first batch)
    CHUNK = int(os.environ.get("W2VSEG_UPLOAD_CHUNK", 14 * 320_000))

    def __init__(self, host_wave, device, side, buf=None, lo: int = 0):
        import torch

        self.host = torch.from_numpy(np.ascontiguousarray(host_wave, dtype=np.float32))
        self.n = self.host.numel()
        self.side = side
        with torch.cuda.stream(side):
            if buf is None or buf.numel() < self.n:
                buf = torch.empty(max(self.n, 1), dtype=torch.float32, device=device)
        self.buf = buf
        self.dev = buf[: self.n]
        self.hi = min(max(int(lo), 0), self.n)   # samples [lo, hi) are enqueued
        self.waited = self.hi                    # ... and [lo, waited) are ordered before the main stream
        self.events = []                         # (hi after the chunk, event), stream order

    def upload_to(self, upto: int) -> None:
        """enqueue copies (side stream) until sample `upto` is covered; does not touch the main stream"""
        import torch

        upto = min(int(upto), self.n)
        with torch.cuda.stream(self.side):
            while self.hi < upto:
                e = min(self.hi + self.CHUNK, self.n)
                self.dev[self.hi: e].copy_(self.host[self.hi: e], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.side)
                self.hi = e
                self.events.append((e, ev))

    def wait(self, upto: int, main) -> None:
        """make the main stream wait until samples below `upto` have arrived"""
        upto = min(int(upto), self.n)
        if upto <= self.waited:
            return
        self.upload_to(upto)
        while self.events and self.events[0][0] < upto:
            self.events.pop(0)
        hi, ev = self.events[0]
        main.wait_event(ev)
        self.waited = hi


@dataclass
class TalkResult:
    probs: np.ndarray                 # averaged over tilings, float64 [n_frames]
    per_tiling: list = field(default_factory=list)


class TalkRunner:
    """Runs whole talks through an SFCEngine. `device_batch` is the number of windows per forward
    call (independent of the reference `batch_size`, which only shapes the metadata)."""

    def __init__(self, engine, batch_size: int = 14, segment_sec: float = 20, inference_times: int = 1,
                 device_batch: int | None = None, dist_group=None):
        self.engine = engine
        self.batch_size = int(batch_size)
        self.segment_sec = segment_sec
        self.inference_times = int(inference_times)
        self.device_batch = int(device_batch or batch_size)
        self.dist_group = dist_group

    # -------------------------------------------------------------------------------------
    def plan(self, waves: list[np.ndarray]) -> tuple[list[Window], list[int]]:
        """global window list (talk-major, tiling-major). CollateFn's `included` flag
        (lib/datautils.py:88, "sample sum != 0") is decided on the device by the forward pass, so
        the host never has to touch the samples; every window is planned as included."""
        wins, n_frames = [], []
        for t, wave in enumerate(waves):
            dur = len(wave)
            n_frames.append(samples_to_frames(dur))
            for i in range(self.inference_times):
                wins += plan_tiling(dur, self.segment_sec, self.inference_times, i, self.batch_size, t)
        return wins, n_frames

    def _stage_meta(self, shard: list[Window], wins: list[Window], n_frames: list[int], stream=None):
        """All small integer inputs of a job — per device batch (sample_len | norm_len | out_len) and per
        (talk, tiling) the scatter plan (start | count | NaN frames) — packed into ONE pinned host array
        and sent with ONE copy on `stream` (the upload stream), instead of a handful of tiny pageable
        copies on the compute stream between the forward passes. Returns (event, batch_meta, plans):
        batch_meta[g] = int32 [3, len(group g)], plans[(talk, tiling)] = (start, count, nan_idx, lo, hi)
        with wins[lo:hi] the windows of that tiling."""
        import torch

        eng = self.engine
        parts, off = [], 0
        batch_slices, plan_slices = [], {}

        def push(a):
            nonlocal off
            a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
            parts.append(a)
            off += a.size
            return off - a.size, a.size

        for b0 in range(0, len(shard), self.device_batch):
            group = shard[b0: b0 + self.device_batch]
            o, _ = push([[w.n_samples for w in group], [w.norm_len for w in group], [w.out_len for w in group]])
            batch_slices.append((o, len(group)))
        k = 0
        while k < len(wins):   # wins are talk-major, tiling-major: one contiguous run per (talk, tiling)
            j = k
            while j < len(wins) and wins[j].talk == wins[k].talk and wins[j].tiling == wins[k].tiling:
                j += 1
            st, ct, nan_idx = scatter_plan(wins[k:j], n_frames[wins[k].talk])
            plan_slices[(wins[k].talk, wins[k].tiling)] = (push(st), push(ct), push(nan_idx), k, j)
            k = j
        total = max(off, 1)
        host = torch.empty(total, dtype=torch.int32, pin_memory=True)
        if off:
            host[:off] = torch.from_numpy(np.concatenate(parts))
        stream = stream or torch.cuda.current_stream(eng.device)
        with torch.cuda.stream(stream):
            dev = torch.empty(total, dtype=torch.int32, device=eng.device)
            dev.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        dev.record_stream(torch.cuda.current_stream(eng.device))
        batch_meta = [dev[o: o + 3 * n].view(3, n) for o, n in batch_slices]
        plans = {key: (dev[a[0]: a[0] + a[1]], dev[b[0]: b[0] + b[1]], dev[c[0]: c[0] + c[1]], lo, hi)
                 for key, (a, b, c, lo, hi) in plan_slices.items()}
        return ev, batch_meta, plans

    def _forward_rows(self, waves_dev, wins: list[Window], r_max: int, batch_meta=None):
        """probability rows fp32 [len(wins), r_max + 1] on the device for the given windows; the
        last column carries the window's `included` flag (1.0 / 0.0)"""
        import torch

        eng = self.engine
        rows = torch.empty(len(wins), r_max + 1, dtype=torch.float32, device=eng.device)
        main = torch.cuda.current_stream(eng.device)

        def need(group):   # per talk: one past the last sample the device batch reads
            out = {}
            for w in group:
                out[w.talk] = max(out.get(w.talk, 0), w.end)
            if len({w.talk for w in group}) == 1 and len(group) > 1:   # strided view runs to a full last row
                lm = max(w.n_samples for w in group)
                out[group[0].talk] = max(out[group[0].talk],
                                         min(group[0].start + (len(group) - 1) * (group[1].start - group[0].start) + lm,
                                             waves_dev[group[0].talk].n))
            return out

        groups = [wins[b0: b0 + self.device_batch] for b0 in range(0, len(wins), self.device_batch)]
        for gi, b0 in enumerate(range(0, len(wins), self.device_batch)):
            group = groups[gi]
            for t, upto in need(group).items():
                waves_dev[t].wait(upto, main)
            lmax = max(w.n_samples for w in group)
            if lmax < 400:
                rows[b0: b0 + len(group)].zero_()
                rows[b0: b0 + len(group), r_max] = 1.0
                continue  # shorter than one receptive field: no frames at all
            first = group[0]
            step = group[1].start - first.start if len(group) > 1 else lmax
            regular = (all(w.talk == first.talk for w in group) and step >= lmax
                       and all(w.start == first.start + k * step for k, w in enumerate(group))
                       and first.start + (len(group) - 1) * step + lmax <= waves_dev[first.talk].n)
            if regular:
                # consecutive windows of one talk: a strided VIEW of the talk's samples is the batch
                # (rows may run past a short last window: sample_len masks that)
                stage = waves_dev[first.talk].dev.as_strided((len(group), lmax), (step, 1), first.start)
            else:
                stage = torch.zeros(len(group), lmax, dtype=torch.float32, device=eng.device)
                for k, w in enumerate(group):
                    stage[k, : w.n_samples] = waves_dev[w.talk].dev[w.start: w.end]
            if batch_meta is not None:
                meta = batch_meta[gi]
            else:
                meta = torch.tensor([[w.n_samples for w in group], [w.norm_len for w in group],
                                     [w.out_len for w in group]], dtype=torch.int32).to(eng.device, non_blocking=True)
            # probabilities, zero tail and the `included` flag land in `rows` straight from the head kernel
            eng.sfc_forward_rows(stage, meta[0], meta[1], meta[2], lmax, rows[b0: b0 + len(group)], flag_col=r_max)
            if gi + 1 < len(groups):   # the next batch's samples travel while this one computes
                for t, upto in need(groups[gi + 1]).items():
                    waves_dev[t].upload_to(upto)
        return rows

    def _side(self):
        import torch

        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream(self.engine.device)
        return side

    def run(self, waves: list[np.ndarray], results_on: int | None = None) -> list[TalkResult] | None:
        """waves: one float32 array of raw samples per talk. Returns per-talk probabilities.
        With a process group, windows are sharded over the ranks and the probability rows are gathered;
        `results_on=r` makes only rank r assemble the talks and copy them to the host (what segment.py
        needs: rank 0 writes the yaml) — the other ranks return None and skip that work."""
        import torch

        eng = self.engine
        wins, n_frames = self.plan(waves)
        r_max = max([eng.frame_stride(max(w.n_samples, 400)) for w in wins] + [1])
        world, rank = 1, 0
        if self.dist_group is not None:
            import torch.distributed as dist

            world, rank = dist.get_world_size(self.dist_group), dist.get_rank(self.dist_group)
        lo, hi = shard_ranges(len(wins), world)[rank]
        side = self._side()
        side.wait_stream(torch.cuda.current_stream(eng.device))
        waves_dev = {}
        for w in wins[lo:hi]:
            if w.talk not in waves_dev:   # windows are talk-major, tiling-major: the first one starts lowest
                first = min(x.start for x in wins[lo:hi] if x.talk == w.talk)
                waves_dev[w.talk] = LazyWave(waves[w.talk], eng.device, side, lo=first)
        mine = results_on is None or results_on == rank      # does this rank assemble the talks?
        ev, batch_meta, plans = self._stage_meta(wins[lo:hi], wins if mine else [], n_frames, side)
        torch.cuda.current_stream(eng.device).wait_event(ev)
        rows = self._forward_rows(waves_dev, wins[lo:hi], r_max, batch_meta)
        for lw in waves_dev.values():
            lw.buf.record_stream(torch.cuda.current_stream(eng.device))
        if world > 1:
            rows = gather_rows(rows, len(wins), world, self.dist_group)
        return self.reduce(rows, wins, n_frames, plans) if mine else None

    def reduce_device(self, rows, wins: list[Window], n_frames: list[int], plans=None):
        """rows [len(wins), r] (device) -> per talk (avg float64 [n], tilings float64 [inference_times, n]) on
        the device: scatter, NaN fill and tiling average kernels, no host round trip. `plans` = the
        scatter plans already on the device (_stage_meta); without it they are computed and sent here."""
        import torch

        eng = self.engine
        out = []
        for t, n in enumerate(n_frames):
            tilings = torch.empty(self.inference_times, n, dtype=torch.float64, device=eng.device)
            for i in range(self.inference_times):
                if plans is not None and (t, i) in plans:
                    st, ct, nan_idx, k0, k1 = plans[(t, i)]
                    talk_rows = rows[k0:k1]
                else:
                    idx = [k for k, w in enumerate(wins) if w.talk == t and w.tiling == i]
                    sub = [wins[k] for k in idx]
                    st, ct, nan_idx = scatter_plan(sub, n)
                    talk_rows = rows[idx[0]: idx[-1] + 1] if idx else rows[:0]
                talk = eng.scatter_rows(talk_rows, st, ct, n, flag_col=talk_rows.shape[1] - 1)
                eng.nanfill(talk, nan_idx)
                tilings[i] = talk
            out.append((eng.overlap_average(tilings), tilings))
        return out

    def reduce(self, rows, wins: list[Window], n_frames: list[int], plans=None) -> list[TalkResult]:
        """rows [len(wins), r] (device) -> per-talk averaged probabilities on the host. All reduction
        kernels are queued first; the results then leave in ONE pass of async copies into one pinned
        buffer and one synchronisation (a `.cpu()` per talk and tiling costs a device sync each)."""
        import torch

        dev = self.reduce_device(rows, wins, n_frames, plans)
        it = self.inference_times
        total = sum(n for n in n_frames) * (1 + (it if it > 1 else 0))
        host = torch.empty(max(total, 1), dtype=torch.float64, pin_memory=True)
        out, off = [], 0
        for (avg, til), n in zip(dev, n_frames):
            a = host[off: off + n]
            a.copy_(avg, non_blocking=True)
            off += n
            parts = [a]
            if it > 1:
                t = host[off: off + it * n].view(it, n)
                t.copy_(til, non_blocking=True)
                off += it * n
                parts = [t[i] for i in range(it)]
            out.append((a, parts))
        torch.cuda.current_stream(self.engine.device).synchronize()
        return [TalkResult(a.numpy(), [p.numpy() for p in parts]) for a, parts in out]

    def run_stream(self, talks, depth: int = 2):
        """Throughput API: yields one TalkResult per input talk (same values as run([wave])[0]), with a
        `depth`-deep software pipeline across talks — the host->device copy of talk k+1 and the
        device->host copy of talk k-1 run on a side stream while the forward of talk k computes.
        `talks` is an iterable of float32 sample arrays (pinned host memory makes the copies truly
        asynchronous). Replaces the reference's per-talk loop with its blocking copies
        (segment.py:71-124, lib/evaluate.py:36-44,96-97)."""
        import torch

        eng = self.engine
        main = torch.cuda.current_stream(eng.device)
        side = self._side()
        down = getattr(self, "_down_stream", None)
        if down is None:
            down = self._down_stream = torch.cuda.Stream(eng.device)
        inflight = []      # (done_event, avg_host, tilings_host)
        slots = [None] * depth   # device wave buffers + the event after which they may be overwritten

        def finish(item):
            ev, avg_h, til_h = item
            ev.synchronize()
            return TalkResult(avg_h.numpy().copy(), [til_h[i].numpy().copy() for i in range(self.inference_times)])

        for k, wave in enumerate(talks):
            wave = np.ascontiguousarray(wave, dtype=np.float32)
            slot = k % depth
            wins, n_frames = self.plan([wave])
            r_max = max([eng.frame_stride(max(w.n_samples, 400)) for w in wins] + [1])
            world, rank = 1, 0
            if self.dist_group is not None:
                import torch.distributed as dist

                world, rank = dist.get_world_size(self.dist_group), dist.get_rank(self.dist_group)
            lo, hi = shard_ranges(len(wins), world)[rank]
            if slots[slot] is not None:
                side.wait_event(slots[slot][1])              # forward of talk k-depth has consumed the buffer
            first = min([w.start for w in wins[lo:hi]] + [len(wave)])
            lw = LazyWave(wave, eng.device, side, buf=slots[slot][0] if slots[slot] is not None else None, lo=first)
            lw.upload_to(first + 2 * LazyWave.CHUNK)         # the rest follows batch by batch (_forward_rows)
            buf = lw.buf
            ev, batch_meta, plans = self._stage_meta(wins[lo:hi], wins, n_frames, side)
            main.wait_event(ev)
            rows = self._forward_rows({0: lw}, wins[lo:hi], r_max, batch_meta)
            if world > 1:
                rows = gather_rows(rows, len(wins), world, self.dist_group)
            avg, til = self.reduce_device(rows, wins, n_frames, plans)[0]
            fwd = torch.cuda.Event()
            fwd.record(main)
            slots[slot] = (buf, fwd)
            # results leave on their OWN stream: on the upload stream the wait for this forward would
            # hold back the next talk's host->device copies until the forward has finished
            with torch.cuda.stream(down):
                down.wait_event(fwd)
                avg_h = torch.empty(avg.shape, dtype=avg.dtype, pin_memory=True)
                til_h = torch.empty(til.shape, dtype=til.dtype, pin_memory=True)
                avg_h.copy_(avg, non_blocking=True)
                til_h.copy_(til, non_blocking=True)
                avg.record_stream(down)
                til.record_stream(down)
                d2h = torch.cuda.Event()
                d2h.record(down)
            inflight.append((d2h, avg_h, til_h))
            if len(inflight) >= depth:
                yield finish(inflight.pop(0))
        while inflight:
            yield finish(inflight.pop(0))


def gather_rows(rows, n_total: int, world: int, group=None):
    """NCCL (or gloo) all_gather of the per-rank probability rows — the only data that crosses
    NVLink on this path. Shards are contiguous and balanced, so padding every rank to the largest
    shard and trimming after the gather restores window order."""
    import torch
    import torch.distributed as dist

    per = (n_total + world - 1) // world
    pad = torch.zeros(per, rows.shape[1], dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    buf = torch.empty(world * per, rows.shape[1], dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = [buf[r * per: r * per + (hi - lo)] for r, (lo, hi) in enumerate(shard_ranges(n_total, world))]
    return torch.cat(parts, dim=0)
