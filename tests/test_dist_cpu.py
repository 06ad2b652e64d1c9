"""CPU, world_size 2 over gloo: the multi-rank logic of the talk pipeline (window sharding and
the all_gather of probability rows — the path's only collective) is rank-count invariant."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wav2vecsegmenter_b200 import pipeline as pl


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rows_for(lo, hi, width):
    idx = torch.arange(lo, hi, dtype=torch.float32)[:, None]
    return idx * 1000 + torch.arange(width, dtype=torch.float32)[None, :]


def _worker(rank, world, port, n_total, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pl.shard_ranges(n_total, world)[rank]
    rows = _rows_for(lo, hi, width)
    full = pl.gather_rows(rows, n_total, world)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for n in (0, 1, 7, 14, 225, 1800):
        for w in (1, 2, 3, 4, 8):
            r = pl.shard_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_gather_rows_world2_equals_single_rank():
    n_total, width = 23, 17
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, width, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = _rows_for(0, n_total, width).numpy()
    for r in range(2):
        np.testing.assert_array_equal(got[r], expect)


def test_plan_is_deterministic_and_talk_major():
    """every rank derives the same global window list (no communication needed to agree on it)"""
    waves = [np.zeros(700_000, np.float32), np.ones(330_001, np.float32)]

    class _E:  # plan() needs no engine
        pass

    r = pl.TalkRunner(_E(), batch_size=3, inference_times=2)
    a, na = r.plan(waves)
    b, nb = r.plan(waves)
    assert [(w.talk, w.tiling, w.start, w.end, w.norm_len, w.out_len) for w in a] == \
           [(w.talk, w.tiling, w.start, w.end, w.norm_len, w.out_len) for w in b] and na == nb
    keys = [(w.talk, w.tiling, w.start) for w in a]
    assert keys == sorted(keys)
    assert all(w.norm_len >= w.n_samples for w in a)   # silence is detected on the device
