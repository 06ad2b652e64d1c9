"""End-to-end parity of the CUDA SFC path (through the C ABI) against
  (1) golden vectors produced by the UNMODIFIED reference (tests/golden/, oracle/make_golden.py), and
  (2) the CPU oracle (oracle/sfc_oracle.py) on seeded inputs the goldens do not cover.

Tolerance (BASELINE.json north_star): per-frame probabilities within max-abs 2e-2 of the
reference's fp32 forward (the CUDA path computes GEMM/conv/attention operands in bf16 with fp32
accumulation, fp32 residual stream / LayerNorm statistics / head).
"""
import numpy as np
import pytest
import torch

from wav2vecsegmenter_b200 import synth

from util import load_gold, make_batch, out_lens_ref, spec_of

pytestmark = pytest.mark.gpu
PROB_TOL = 2e-2

_engines = {}


def engine_for(spec, seed):
    from wav2vecsegmenter_b200.engine import SFCEngine

    key = (spec, seed)
    if key not in _engines:
        _engines.clear()  # one model resident at a time
        e = SFCEngine(spec)
        e.load_state_dict(synth.random_state_dict(spec, seed))
        _engines[key] = e
    return _engines[key]


@pytest.mark.parametrize("name", ["tiny_batch", "middle_window", "middle_half_batch", "large_batch",
                                  "tiny_gn_batch", "tiny_gn_nobias_batch", "tiny_postln_batch"])
def test_probs_match_reference_golden(name):
    g = load_gold(name)
    spec = spec_of(g)
    lens = [int(x) for x in g["lens"]]
    eng = engine_for(spec, int(g["seed"]))
    audio = make_batch(lens, int(g["audio_seed"])).cuda()
    lmax = max(lens)
    out_mask = torch.from_numpy(g["out_mask"])
    out_len = out_mask.sum(1).tolist()
    # fused path: normalisation over the padded batch row (norm_len = lmax for every window)
    logits, probs = eng.sfc_forward(audio, lens, [lmax] * len(lens), out_len, lmax)
    torch.cuda.synchronize()
    T = out_mask.shape[1]
    p = probs[:, :T].cpu().numpy()
    l = logits[:, :T].cpu().numpy()
    assert np.isfinite(p).all()
    err = np.abs(p - g["probs"]).max()
    assert err <= PROB_TOL, f"{name}: max-abs prob err {err}"
    assert (p[~g["out_mask"]] == 0).all() and (l[~g["out_mask"]] == 0).all()
    # frames the mask excludes beyond T stay zero too
    assert (probs[:, T:].cpu().numpy() == 0).all()
    # decisions at the 0.5 threshold (pDAC/pSTRM): every frame whose margin in the reference exceeds the
    # tolerance must fall on the same side; overall >= 99 % (random-init probabilities hover around 0.5:
    # measured 99.5 % for large_batch, 99.9-100 % for the others, profiles/parity_r02.md)
    same = (p > 0.5) == (g["probs"] > 0.5)
    decisive = np.abs(g["probs"] - 0.5) > PROB_TOL
    assert same[decisive & g["out_mask"]].all()
    agree = same[g["out_mask"]].mean()
    assert agree >= 0.99, f"{name}: only {agree:.4f} of frames on the same side of 0.5"
    if name == "large_batch":   # the headline model: bias-corrected bf16 weights leave half the tolerance as margin
        assert err <= 1e-2, f"{name}: max-abs prob err {err}"


@pytest.mark.parametrize("name", ["tiny_batch", "middle_half_batch", "tiny_gn_batch", "tiny_postln_batch"])
def test_hidden_and_two_call_path_match_golden(name):
    """model.wav2vec_model(...) then model.seg_model(...) as two calls (lib/evaluate.py:59,72)"""
    g = load_gold(name)
    spec = spec_of(g)
    lens = [int(x) for x in g["lens"]]
    eng = engine_for(spec, int(g["seed"]))
    audio = make_batch(lens, int(g["audio_seed"])).cuda()
    lmax = max(lens)
    hidden, enc_len = eng.encode(audio, lens, [lmax] * len(lens), lmax)
    T = int(g["hidden_T"])
    assert hidden.shape[1] >= T
    assert enc_len.tolist() == [eng.num_frames(n) for n in lens]
    frames = g["hidden_frames"]
    h = hidden[:, :T][:, frames].cpu().numpy()
    ref = g["hidden"]
    rel = np.abs(h - ref).max() / np.abs(ref).max()
    assert rel < 2e-2, f"hidden rel err {rel}"
    out_mask = torch.from_numpy(g["out_mask"])
    Tm = out_mask.shape[1]
    logits, probs = eng.head(hidden[:, :Tm], out_mask.sum(1).tolist())
    torch.cuda.synchronize()
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= PROB_TOL


@pytest.mark.parametrize("name", ["tiny_batch", "tiny_gn_batch"])
def test_prenormalised_audio_path(name):
    """audio already normalised by CollateFn (norm_len = 0) == on-device normalisation. For the GroupNorm
    extractor the statistics run over the padded row, whose padding CollateFn has normalised too."""
    g = load_gold(name)
    spec = spec_of(g)
    lens = [int(x) for x in g["lens"]]
    eng = engine_for(spec, int(g["seed"]))
    raw = make_batch(lens, int(g["audio_seed"]))
    norm = (raw - raw.mean(1, keepdim=True)) / raw.std(1, keepdim=True)
    np.testing.assert_allclose(norm[:, :64].numpy(), g["audio_norm_head"], rtol=1e-5, atol=1e-6)
    out_len = torch.from_numpy(g["out_mask"]).sum(1).tolist()
    lmax = max(lens)
    _, p0 = eng.sfc_forward(raw.cuda(), lens, [lmax] * 3, out_len, lmax)
    _, p1 = eng.sfc_forward(norm.cuda(), lens, [0] * 3, out_len, lmax)
    # fp32 rounding differences in the normalised samples are amplified by bf16 rounding of the
    # first conv layer: same noise floor as any bf16 run-to-run perturbation, well inside 2e-2
    assert (p0 - p1).abs().max().item() < 1e-2
    T = g["probs"].shape[1]
    assert np.abs(p1[:, :T].cpu().numpy() - g["probs"]).max() <= PROB_TOL


@pytest.mark.parametrize("lens", [[48000], [35000, 64000, 16000 * 3 + 17, 400 * 40]])
def test_probs_match_oracle_on_seeded_inputs(lens):
    from oracle import sfc_oracle

    spec = synth.TINY
    sd = synth.random_state_dict(spec, 3)
    eng = engine_for(spec, 3)
    raw = make_batch(lens, 77)
    lmax = max(lens)
    out_len = out_lens_ref(lens)
    T_hidden = sfc_oracle.conv_out_frames(lmax)
    out_mask = torch.zeros(len(lens), max(out_len), dtype=torch.bool)
    for i, n in enumerate(out_len):
        out_mask[i, :n] = True
    norm = sfc_oracle.normalize_rows(raw, [True] * len(lens))
    with torch.no_grad():
        ref_p, ref_l, ref_mask, _ = sfc_oracle.batch_probs(sd, norm, lens, out_mask, spec.keep_layers, spec.head_heads)
    ol = ref_mask.sum(1).tolist()
    _, probs = eng.sfc_forward(raw.cuda(), lens, [lmax] * len(lens), ol, lmax)
    p = probs[:, : ref_mask.shape[1]].cpu()
    assert (p - ref_p).abs().max().item() <= PROB_TOL
    assert T_hidden <= eng.frame_stride(lmax) - 1


@pytest.mark.parametrize("spec_name,seed", [("TINY", 11), ("TINY", 12), ("TINY_GN", 13), ("TINY_POSTLN", 14)])
def test_probs_match_oracle_on_edge_lengths(spec_name, seed):
    """ragged batches whose lengths sit on the edges the frame arithmetic has: the shortest encodable window
    (400 samples = 1 frame), one sample either side of a frame boundary of the conv stack (400 + 320 k - 1, + 0,
    + 1), lengths where the reference's round((n + 1e-6) * 49.95 / 16000) and the conv frame count differ by one
    (the +-1 fix-up of lib/evaluate.py:63-70), and an all-zero (silent) row, against the oracle on the same inputs"""
    from oracle import sfc_oracle

    spec = getattr(synth, spec_name)
    rng = np.random.RandomState(seed)
    k = int(rng.randint(20, 120))
    lens = [400 + 320 * k - 1, 400 + 320 * k, 400 + 320 * k + 1, 400, int(rng.randint(500, 30000)),
            int(rng.randint(30000, 46000))]
    rng.shuffle(lens)
    sd = synth.random_state_dict(spec, seed)
    eng = engine_for(spec, seed)
    raw = make_batch(lens, 100 + seed)
    silent = int(rng.randint(0, len(lens)))
    raw[silent] = 0.0
    lmax = max(lens)
    out_len = out_lens_ref(lens)
    out_mask = torch.zeros(len(lens), max(out_len), dtype=torch.bool)
    for i, n in enumerate(out_len):
        out_mask[i, :n] = True
    included = [bool(raw[i].abs().sum() > 0) for i in range(len(lens))]
    norm = sfc_oracle.normalize_rows(raw, included)
    with torch.no_grad():
        ref_p, _, ref_mask, _ = sfc_oracle.batch_probs(sd, norm, lens, out_mask, spec.keep_layers, spec.head_heads,
                                                       post_ln=spec.post_ln)
    ol = ref_mask.sum(1).tolist()
    _, probs = eng.sfc_forward(raw.cuda(), lens, [lmax] * len(lens), ol, lmax)
    p = probs[:, : ref_mask.shape[1]].cpu()
    assert torch.isfinite(p).all()
    live = [i for i in range(len(lens)) if i != silent]      # the silent row is reported as zeros downstream
    err = (p[live] - ref_p[live]).abs().max().item()         # (lib/evaluate.py:100-111), its probabilities are unused
    assert err <= PROB_TOL, (lens, err)
    for i in live:                                           # nothing leaks past a window's own frames
        assert (p[i, ol[i]:] == 0).all()


def test_outlier_channels_in_the_residual_stream():
    """trained XLS-R checkpoints carry a few residual-stream channels that are orders of magnitude larger than the
    rest (massive activations). Emulated on the tiny model: three output biases of the first FFN set to +-80 / 150 and
    one feature-projection bias to 40, so the fp32 residual stream, both LayerNorms after it, the adapters and the
    head see channels ~100x the typical magnitude. The probabilities must stay within the tolerance of the oracle."""
    from oracle import sfc_oracle
    from wav2vecsegmenter_b200.engine import SFCEngine

    spec = synth.TINY
    sd = synth.random_state_dict(spec, 21)
    w2v = "wav2vec_model.model."
    b = sd[w2v + "encoder.layers.0.feed_forward.output_dense.bias"]
    b[3], b[500], b[777] = 80.0, -80.0, 150.0
    sd[w2v + "feature_projection.projection.bias"][123] = 40.0
    lens = [48000, 31000, 40007]
    raw = make_batch(lens, 78)
    lmax = max(lens)
    out_len = out_lens_ref(lens)
    out_mask = torch.zeros(len(lens), max(out_len), dtype=torch.bool)
    for i, n in enumerate(out_len):
        out_mask[i, :n] = True
    norm = sfc_oracle.normalize_rows(raw, [True] * len(lens))
    with torch.no_grad():
        hid = sfc_oracle.encoder(sd, norm, lens, spec.keep_layers)
        ref_p, _, ref_mask, _ = sfc_oracle.batch_probs(sd, norm, lens, out_mask, spec.keep_layers, spec.head_heads)
    assert hid.abs().max().item() > 100 * hid.abs().median().item()     # the outliers are really there
    eng = SFCEngine(spec)
    eng.load_state_dict(sd)
    ol = ref_mask.sum(1).tolist()
    _, probs = eng.sfc_forward(raw.cuda(), lens, [lmax] * len(lens), ol, lmax)
    p = probs[:, : ref_mask.shape[1]].cpu()
    err = (p - ref_p).abs().max().item()
    print(f"OUTLIER max |hidden| {hid.abs().max().item():.1f}, median {hid.abs().median().item():.3f}; max prob err {err:.4f}")
    eng.close()
    assert torch.isfinite(p).all() and err <= PROB_TOL, err


def test_window_independent_of_batch_composition():
    """a window's valid-frame probabilities do not depend on what else is in the device batch
    (the property that lets windows shard across GPUs), given the same norm_len"""
    spec = synth.TINY
    eng = engine_for(spec, 3)
    lens = [64000, 40000]
    raw = make_batch(lens, 5)
    ol = out_lens_ref(lens)
    _, p_pair = eng.sfc_forward(raw.cuda(), lens, [64000, 64000], ol, 64000)
    _, p_single = eng.sfc_forward(raw[1:, :40000].contiguous().cuda(), [40000], [64000], ol[1:], 40000)
    a = p_pair[1, : ol[1]].cpu()
    b = p_single[0, : ol[1]].cpu()
    assert (a - b).abs().max().item() < 1e-5


def test_full_size_batch_properties():
    """BASELINE.json configs[1] at full size (large 24/24 + adapters, 14 x 20 s): size-independent
    properties instead of a CPU comparison — the first two windows equal the reference golden of the
    same model (`large_batch`), every window equals its own single-window run (what lets windows be
    batched and sharded freely), the run is deterministic, probabilities are finite and in (0, 1)."""
    g = load_gold("large_batch")
    spec = spec_of(g)
    eng = engine_for(spec, int(g["seed"]))
    B, L = 14, 320000
    lens = [L] * B
    raw = torch.zeros(B, L)
    raw[:2] = make_batch([int(x) for x in g["lens"]], int(g["audio_seed"]))
    raw[2:] = make_batch([L] * (B - 2), 4242)
    raw[9, 200_000:] = 0            # a window that is silent in its second half
    ol = out_lens_ref(lens)
    dev = raw.cuda()
    _, p = eng.sfc_forward(dev, lens, lens, ol, L)
    p = p.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(p).all() and (p[:, : ol[0]] > 0).all() and (p[:, : ol[0]] < 1).all()
    ref = torch.from_numpy(g["probs"])
    assert (p[:2, : ref.shape[1]].cpu() - ref).abs().max().item() <= PROB_TOL
    _, p2 = eng.sfc_forward(dev, lens, lens, ol, L)
    assert torch.equal(p, p2)                                   # deterministic, bit for bit
    for i in (0, 9, 13):
        _, pi = eng.sfc_forward(dev[i: i + 1].contiguous(), [L], [L], ol[i: i + 1], L)
        assert (pi[0, : ol[i]] - p[i, : ol[i]]).abs().max().item() < 1e-5, i


def test_cuda_graph_forward_matches_direct():
    """the forward is capturable as it is (no allocation, no host synchronisation): a CUDA-graph replay gives
    bit-identical probabilities, also after the inputs changed in place"""
    import time

    spec = synth.TINY
    eng = engine_for(spec, 3)
    B, L = 2, 64000
    g = eng.graphed(B, L)
    for seed in (5, 6):
        raw = make_batch([L, 50000], seed).cuda()
        lens = torch.tensor([L, 50000], dtype=torch.int32, device="cuda")
        ol = torch.tensor([eng.num_frames(L), eng.num_frames(50000)], dtype=torch.int32, device="cuda")
        g.audio.copy_(raw); g.sample_len.copy_(lens); g.out_len.copy_(ol)
        p_graph = g.replay().clone()
        _, p_direct = eng.sfc_forward(raw, lens, g.norm_len, ol, L)
        torch.cuda.synchronize()
        assert torch.equal(p_graph, p_direct)
    t0 = time.perf_counter()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(20):
        eng.sfc_forward(g.audio, g.sample_len, g.norm_len, g.out_len, L, g.logits, g.probs)
    torch.cuda.synchronize()
    t_direct = (time.perf_counter() - t0) / 20
    print(f"GRAPH tiny model, batch 2: graph replay {t_graph * 1e3:.3f} ms vs direct {t_direct * 1e3:.3f} ms per forward")
