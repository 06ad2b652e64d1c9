"""CPU: the product's host algorithms (lib/segment.py, pipeline window plan) against the golden
vectors produced by the reference's own functions (bit-exact: integer / index work and fp64)."""
import importlib
import json
import sys

import numpy as np
import pytest
import yaml

from util import load_gold


@pytest.fixture(scope="module")
def seg():
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]
    return importlib.import_module("lib.segment")


def test_pdac_strm_pthr_match_reference(seg):
    g = load_gold("algos")
    for c in range(int(g["n_cases"])):
        p = g[f"p_{c}"]
        for tag, fn in (("dac", seg.pdac), ("strm", seg.strm), ("pthr", seg.pthr)):
            kw = yaml.safe_load(str(g[f"{tag}_kw_{c}"]))
            if tag == "pthr" and kw["moving_average_window"] > 0:
                continue  # moving average runs on the GPU: covered by tests/test_talk_gpu.py
            segs = fn(p, **kw)
            got = np.array([[s.start, s.end] for s in segs], dtype=np.float64).reshape(-1, 2)
            np.testing.assert_array_equal(got, g[f"{tag}_bounds_{c}"], err_msg=f"{tag} case {c}")
            text = yaml.dump(seg.update_yaml_content([], segs, "a.wav"), default_flow_style=True)
            assert text == str(g[f"{tag}_yaml_{c}"]), f"{tag} case {c}"


@pytest.mark.parametrize("name", ["speech_talk", "speech_talk_x2", "speech_talk_large", "speech_talk_mh_x2"])
def test_decisive_talk_fixture_host_side(seg, name):
    """the decisive-probability fixtures (oracle/make_golden.py:gold_talk_decisive): pDAC / pSTRM on
    the reference's own probabilities reproduce the reference's boundaries and yaml bit-exactly, and
    the fixture is what it claims to be (>= 99 % of frames far from the thresholds)."""
    g = load_gold(name)
    p = g["probs_avg"]
    assert (np.abs(p - 0.5) > 0.4).mean() > 0.99
    assert ((p > 0.5) == g["labels"]).mean() > 0.99
    kws = {"dac": dict(max_segment_length=16, min_segment_length=0.2, threshold=0.5),
           "strm": dict(max_segment_length=18, min_segment_length=0.2, min_pause_length=0.2, threshold=0.5)}
    for tag, fn in (("dac", seg.pdac), ("strm", seg.strm)):
        segs = fn(p, **kws[tag])
        got = np.array([[s.start, s.end] for s in segs], dtype=np.float64).reshape(-1, 2)
        np.testing.assert_array_equal(got, g[f"{tag}_bounds"])
        text = yaml.dump(seg.update_yaml_content([], segs, "talk.wav"), default_flow_style=True)
        assert text == str(g[f"{tag}_yaml"])


def test_segment_edge_cases(seg):
    assert seg.pdac(np.zeros(10), 16, 0.2, 0.5)[0].duration == 0.0
    assert seg.strm(np.zeros(0)) == []
    assert seg.pthr(np.zeros(0)) == []
    assert seg.pthr(np.ones(5), threshold=0.5)[0].start == 0
    s = seg.Segment(10, 1009)
    assert s.offset == round(10 / 49.95, 6) and s.duration == 20.0


def test_window_plan_matches_reference():
    from wav2vecsegmenter_b200 import pipeline as pl

    g = load_gold("plan")
    for d, it, i, n_out, starts, ends, sf, ef in json.loads(str(g["plans"])):
        s, e = pl.tiling_bounds(d, 20, it, i)
        assert (s, e) == (starts, ends)
        assert pl.samples_to_frames(d) == n_out
        wins = pl.plan_tiling(d, 20, it, i, 14)
        assert [w.start_f for w in wins] == sf and [w.end_f for w in wins] == ef
        for w in wins:
            assert 0 <= w.out_len <= w.end_f - w.start_f
            assert w.norm_len >= w.n_samples


def test_num_frames_matches_c_abi():
    from wav2vecsegmenter_b200 import _native, pipeline as pl

    lib = _native.load()
    for n in [0, 399, 400, 401, 719, 720, 16000, 113234, 320000, 352000, 1234567]:
        assert lib.w2vseg_num_frames(n) == pl.num_frames(n)
        assert lib.w2vseg_frame_stride(max(n, 1)) >= pl.num_frames(n) + 1


def test_scatter_plan_covers_like_reference():
    """frames left uncovered (NaN in the reference) appear exactly where `ends -= 1` bites"""
    from oracle import host_oracle as ho
    from wav2vecsegmenter_b200 import pipeline as pl

    d = 1_073_234
    for it in (1, 2):
        for i in range(it):
            wins = pl.plan_tiling(d, 20, it, i, 3)
            n = pl.samples_to_frames(d)
            st, ct, nan_idx = pl.scatter_plan(wins, n)
            talk = np.full(n, np.nan)
            for b0 in range(0, len(wins), 3):
                grp = wins[b0:b0 + 3]
                probs = np.ones((len(grp), 1200))
                ho.scatter_batch(talk, probs, [w.start_f for w in grp], [w.end_f for w in grp],
                                 [True] * len(grp), grp[0].ends_shift)
            assert np.array_equal(np.flatnonzero(np.isnan(talk)), nan_idx)
