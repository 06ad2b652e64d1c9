"""Dev-set scoring path: labelled fixed-length datasets and evaluate() against
golden vectors produced by the unmodified reference (oracle/make_golden.py: tiny_eval, tiny_eval_x1)."""
import json
import sys

import numpy as np
import pytest

from oracle.make_golden import EVAL_TALKS, write_eval_corpus
from util import load_gold


def _dropin():
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.") or k in ("constants", "datautils")]:
        del sys.modules[k]
    import lib.dataset as ds
    import lib.evaluate as ev

    return ds, ev


@pytest.mark.parametrize("name", ["tiny_eval", "tiny_eval_x1"])
def test_labelled_windows_and_targets_match_reference(tmp_path, name):
    """window plan, `included` strings, target tensors and frame ranges: bit-exact"""
    ds, _ = _dropin()
    g = load_gold(name)
    it, bs = int(g["inference_times"]), int(g["batch_size"])
    tl, sl = write_eval_corpus(tmp_path)
    gen = ds.FixedDataloaderGenerator(str(tl), str(sl), 20, bs, 0, it)
    assert gen.get_talk_ids() == [t[0] for t in EVAL_TALKS]
    for tid, *_ in EVAL_TALKS:
        for i in range(it):
            dl = gen.generate(tid, i)
            df = gen.dataset.fixed_segments_df
            assert list(df.start) == list(g[f"{tid}_{i}_starts"]) and list(df.end) == list(g[f"{tid}_{i}_ends"])
            assert list(df.included) == json.loads(str(g[f"{tid}_{i}_included"]))
            items = [gen.dataset[k] for k in range(len(gen.dataset))]
            assert np.array_equal(np.concatenate([t[1].numpy() for t in items]), g[f"{tid}_{i}_targets"])
            assert [len(t[1]) for t in items] == list(g[f"{tid}_{i}_target_lens"])
            assert [[t[2], t[3]] for t in items] == g[f"{tid}_{i}_frames"].tolist()
            batch = next(iter(dl))     # collated like the reference: padded targets, masks
            assert batch["target"].shape[1] == max(t[3] - t[2] for t in items[:bs])
        assert gen.dataset.duration_outframes == int(g[f"{tid}_duration_outframes"])


@pytest.mark.parametrize("name", ["tiny_eval", "tiny_eval_x1"])
def test_evaluate_aggregation_matches_reference(tmp_path, name, monkeypatch):
    """evaluate() on the reference's own per-tiling probabilities / targets / losses reproduces its
    metrics exactly (incl. the double division by inference_times and the last-talk eval_loss)"""
    ds, ev = _dropin()
    g = load_gold(name)
    it, bs = int(g["inference_times"]), int(g["batch_size"])
    tl, sl = write_eval_corpus(tmp_path)
    gen = ds.FixedDataloaderGenerator(str(tl), str(sl), 20, bs, 0, it)
    calls = []

    def fake_infer(model, dataloader, *a, **k):
        tid = dataloader.dataset.fixed_segments_df.talk_id.iloc[0]
        i = sum(1 for c in calls if c == tid)
        calls.append(tid)
        return (g[f"{tid}_{i}_probs"].copy(), None, g[f"{tid}_{i}_talk_targets"].copy(), float(g[f"{tid}_{i}_loss"]))

    monkeypatch.setattr(ev, "infer", fake_infer)
    res = ev.evaluate(gen, None, None, False, "bce", None, loss_fn=object())
    ref = json.loads(str(g["metrics"]))
    assert {k: float(v) for k, v in res.items()} == ref
