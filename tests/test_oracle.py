"""CPU: pins the oracle (oracle/*.py) to the golden vectors the UNMODIFIED reference produced
(oracle/make_golden.py). The oracle is only trusted by the GPU parity tests because this passes."""
import json

import numpy as np
import pytest
import torch
import yaml

from oracle import host_oracle as ho
from oracle import sfc_oracle
from wav2vecsegmenter_b200 import synth

from util import load_gold, make_batch, spec_of


@pytest.mark.parametrize("name", ["tiny_batch", "middle_window", "tiny_gn_batch", "tiny_gn_nobias_batch",
                                  "tiny_postln_batch"])
def test_forward_oracle_matches_reference(name):
    g = load_gold(name)
    spec = spec_of(g)
    lens = [int(x) for x in g["lens"]]
    sd = synth.random_state_dict(spec, int(g["seed"]))
    raw = make_batch(lens, int(g["audio_seed"]))
    norm = sfc_oracle.normalize_rows(raw, [True] * len(lens))
    np.testing.assert_allclose(norm[:, :64].numpy(), g["audio_norm_head"], rtol=1e-5, atol=1e-6)
    out_len = [int(np.round((n + 1e-6) * 49.95 / 16000)) for n in lens]
    out_mask = torch.zeros(len(lens), max(out_len), dtype=torch.bool)
    for i, n in enumerate(out_len):
        out_mask[i, :n] = True
    torch.set_num_threads(8)
    with torch.no_grad():
        hidden = sfc_oracle.encoder(sd, norm, lens, spec.keep_layers, post_ln=spec.post_ln)
        assert hidden.shape[1] == int(g["hidden_T"])
        ref_h = g["hidden"]
        rel = np.abs(hidden[:, g["hidden_frames"]].numpy() - ref_h).max() / np.abs(ref_h).max()
        assert rel < 2e-4, rel
        probs, logits, mask, _ = sfc_oracle.batch_probs(sd, norm, lens, out_mask, spec.keep_layers, spec.head_heads,
                                                        post_ln=spec.post_ln)
    assert (mask.numpy() == g["out_mask"]).all()
    assert np.abs(probs.numpy() - g["probs"]).max() < 2e-4
    assert np.abs(logits.numpy() - g["logits"]).max() < 2e-3


def test_window_plan_matches_reference():
    g = load_gold("plan")
    recs = json.loads(str(g["plans"]))
    assert len(recs) >= 100
    for d, it, i, n_out, starts, ends, sf, ef in recs:
        s, e = ho.window_plan(d, 20, it, i)
        assert (s, e) == (starts, ends), (d, it, i)
        assert ho.to_outframes(d) == n_out
        fr = [ho.window_frames(a, b) for a, b in zip(s, e)]
        assert [f[0] for f in fr] == sf and [f[1] for f in fr] == ef


def test_host_algorithms_match_reference():
    g = load_gold("algos")
    fns = {"dac": ho.pdac, "strm": ho.strm, "pthr": ho.pthr}
    for c in range(int(g["n_cases"])):
        p = g[f"p_{c}"]
        np.testing.assert_array_equal(ho.moving_average(p, int(g[f"maw_{c}"])), g[f"ma_{c}"])
        for tag, fn in fns.items():
            kw = yaml.safe_load(str(g[f"{tag}_kw_{c}"]))
            segs = fn(p, **kw)
            got = np.array([[s.start, s.end] for s in segs], dtype=np.float64).reshape(-1, 2)
            np.testing.assert_array_equal(got, g[f"{tag}_bounds_{c}"], err_msg=f"{tag} case {c}")
            text = yaml.dump(ho.yaml_records(segs, "a.wav"), default_flow_style=True)
            assert text == str(g[f"{tag}_yaml_{c}"]), f"{tag} case {c}"
