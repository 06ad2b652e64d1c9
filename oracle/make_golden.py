"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference via oracle/ref_shims.py) on CPU in the build container. The reference cannot
travel to the GPU box, so these vectors are what pins both the oracle (tests/test_oracle.py) and
the CUDA path (tests/test_parity_gpu.py) to the reference's actual behaviour.

    python oracle/make_golden.py [--only NAME]

Every fixture stores only seeds + reference OUTPUTS; weights and audio are regenerated from the
seeds with wav2vecsegmenter_b200.synth (pure torch-CPU RNG, identical on every machine).
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
import time
import wave
from pathlib import Path

import numpy as np
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

from wav2vecsegmenter_b200 import synth  # noqa: E402


def write_wav(path, x: torch.Tensor):
    pcm = torch.round(x * 32767.0).clamp(-32768, 32767).to(torch.int16).numpy()
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(pcm.tobytes())


def build_reference_model(ns, spec: synth.ModelSpec, seed: int):
    from oracle import ref_shims

    finetune = spec.adapter_layers > 0
    name = "facebook/wav2vec2-xls-r-300m"
    if spec.post_ln:
        assert spec.feat_norm == "group" and not spec.conv_bias
        name = ref_shims.POST_LN_NAME
    elif spec.feat_norm == "group":
        name = ref_shims.GROUP_NORM_NAME if spec.conv_bias else ref_shims.GROUP_NORM_NOBIAS_NAME
    m = ns.models.SHAS(name, spec.keep_layers, finetune,
                       spec.adapter_layers if finetune else 99, False, False, True,
                       spec.head_layers, spec.head_heads, 0.1)
    sd = synth.random_state_dict(spec, seed)
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    m.eval()
    return m, sd


def gold_batch(ns, name, spec, seed, lens, audio_seed):
    """one collated batch through model.wav2vec_model / model.seg_model exactly like
    lib/evaluate.py:58-91"""
    m, _ = build_reference_model(ns, spec, seed)
    waves = [synth.synthetic_audio(n, audio_seed + i) for i, n in enumerate(lens)]
    starts = [0] * len(lens)
    ends = [int(np.round((n + 1e-6) * 49.95 / 16000)) for n in lens]
    batch = ns.datautils.CollateFn(0)([(w, None, s, e) for w, s, e in zip(waves, starts, ends)])
    t0 = time.time()
    with torch.no_grad():
        _, hidden = m.wav2vec_model(batch["audio"], batch["in_mask"])
        out_mask = batch["out_mask"]
        size1, size2 = hidden.shape[1], out_mask.shape[1]
        hid = hidden
        if size1 != size2:
            if size1 < size2:
                out_mask = out_mask[:, :-1]
            else:
                hid = hidden[:, :-1, :]
        logits = m.seg_model(hid, out_mask)
        probs = torch.sigmoid(logits)
        probs[~out_mask] = 0
        logits[~out_mask] = 0
    dt = time.time() - t0
    frames = sorted(set(list(range(0, hidden.shape[1], 9)) + list(range(max(0, hidden.shape[1] - 4), hidden.shape[1]))))
    np.savez_compressed(
        GOLD / f"{name}.npz",
        spec=np.array([spec.keep_layers, spec.adapter_layers, spec.head_layers, spec.head_heads,
                       int(spec.feat_norm == "group"), int(spec.conv_bias), int(spec.post_ln)]),
        seed=seed, audio_seed=audio_seed, lens=np.array(lens),
        hidden_frames=np.array(frames), hidden=hidden[:, frames, :].numpy().astype(np.float32),
        hidden_T=hidden.shape[1],
        out_mask=out_mask.numpy(), probs=probs.numpy(), logits=logits.numpy(),
        audio_norm_head=batch["audio"][:, :64].numpy(),
        ref_seconds=dt,
    )
    print(f"{name}: hidden {tuple(hidden.shape)} probs {tuple(probs.shape)} ref {dt:.2f}s "
          f"prob range [{probs.min():.3f}, {probs.max():.3f}]")


def gold_talk(ns, name, spec, seed, n_samples, audio_seed, inference_times, batch_size):
    """the per-talk body of generate() (segment.py:71-124): dataset -> DataLoader -> infer ->
    tiling average -> pdac / strm / pthr -> yaml"""
    from torch.utils.data import DataLoader

    m, _ = build_reference_model(ns, spec, seed)
    x = synth.synthetic_audio(n_samples, audio_seed)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        wav = Path(td) / "talk.wav"
        write_wav(wav, x)
        ds = ns.dataset.FixedSegmentationDatasetNoTarget(wav, 20, inference_times)
        acc = None
        for i in range(inference_times):
            ds.fixed_length_segmentation(i)
            out[f"starts_{i}"] = np.array(ds.starts)
            out[f"ends_{i}"] = np.array(ds.ends)
            dl = DataLoader(ds, batch_size=batch_size, num_workers=0, shuffle=False, drop_last=False,
                            collate_fn=ns.datautils.CollateFn(0))
            probs, logits, _, _ = ns.evaluate.infer(m, dl, torch.device("cpu"), False, "bce", None)
            out[f"probs_{i}"] = probs.copy()
            acc = probs.copy() if acc is None else acc + probs
        acc /= inference_times
    out["probs_avg"] = acc
    algos = {
        "dac": (ns.segment.pdac, dict(max_segment_length=16, min_segment_length=0.2, threshold=0.5)),
        "strm": (ns.segment.strm, dict(max_segment_length=18, min_segment_length=0.2, min_pause_length=0.2, threshold=0.5)),
        "pthr": (ns.segment.pthr, dict(max_segment_length=28, min_segment_length=0.2, max_lerp_range=4,
                                       min_lerp_range=0.4, threshold=0.1, moving_average_window=0.1)),
    }
    for tag, (fn, kw) in algos.items():
        segs = fn(acc, **kw)
        out[f"{tag}_bounds"] = np.array([[s.start, s.end] for s in segs], dtype=np.float64).reshape(-1, 2)
        content = ns.segment.update_yaml_content([], segs, "talk.wav")
        out[f"{tag}_yaml"] = np.array(yaml.dump(content, default_flow_style=True))
    np.savez_compressed(
        GOLD / f"{name}.npz",
        spec=np.array([spec.keep_layers, spec.adapter_layers, spec.head_layers, spec.head_heads]),
        seed=seed, audio_seed=audio_seed, n_samples=n_samples, inference_times=inference_times,
        batch_size=batch_size, duration_outframes=int(ds.duration_outframes), **out)
    print(f"{name}: {len(acc)} frames, nan-free={not np.isnan(acc).any()}, "
          + ", ".join(f"{t}:{len(out[t + '_bounds'])}" for t in algos))


def gold_talk_decisive(ns, name, spec, seed, n_samples, audio_seed, inference_times, batch_size,
                       fit_batches=None, probs_every=1):
    """see below; fit_batches: fit the output layer on the features of the first N reference batches
    only (long talks); probs_every: store every k-th frame of the probability vectors (fp32) instead of
    all of them (long talks: the fixture pins the yaml / boundaries, the subsample is for diagnosis)."""
    return _gold_talk_decisive(ns, name, spec, seed, n_samples, audio_seed, inference_times, batch_size,
                               fit_batches, probs_every)


def _gold_talk_decisive(ns, name, spec, seed, n_samples, audio_seed, inference_times, batch_size,
                        fit_batches, probs_every):
    """Boundary-parity fixture with DECISIVE probabilities (what a trained checkpoint produces; the
    random-init tracks of gold_talk hover around the thresholds). Audio = synth.speech_like_audio
    (noise bursts / near-silent pauses); the model is the seeded random one, except that its final
    Linear(1024 -> 1) (lib/models.py:304,319) is ridge-fitted on the REFERENCE's own features (hook on
    seg_model.layer_norm) to +-3.5 logits for speech / pause frames. The fitted row and bias are
    stored in the fixture, so the CUDA test rebuilds exactly this checkpoint. Everything after the
    fit is the unmodified reference pipeline, as in gold_talk."""
    from torch.utils.data import DataLoader

    m, sd = build_reference_model(ns, spec, seed)
    x, lab = synth.speech_like_audio(n_samples, audio_seed)
    feats = []
    hook = m.seg_model.layer_norm.register_forward_hook(
        lambda mod, i, o: feats.append(o.detach().clone()) if fit_batches is None or len(feats) < fit_batches else None)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        wav = Path(td) / "talk.wav"
        write_wav(wav, x)
        ds = ns.dataset.FixedSegmentationDatasetNoTarget(wav, 20, inference_times)

        def loader():
            return DataLoader(ds, batch_size=batch_size, num_workers=0, shuffle=False, drop_last=False,
                              collate_fn=ns.datautils.CollateFn(0))

        # pass 1 (tiling 0, random output layer): collect the features and fit the output layer
        ds.fixed_length_segmentation(0)
        probs0, _, _, _ = ns.evaluate.infer(m, loader(), torch.device("cpu"), False, "bce", None)
        hook.remove()
        n_frames = len(probs0)
        lab = lab[:n_frames]
        Z = np.zeros((n_frames if fit_batches is None else min(n_frames, fit_batches * batch_size * 1000),
                      spec.hidden), dtype=np.float32 if fit_batches else np.float64)
        have = np.zeros(n_frames, bool)
        fr = lambda v: int(np.round((v + 1e-6) * 49.95 / 16000))
        j = 0
        for zb in feats:
            for i in range(zb.shape[0]):
                s0, e0 = fr(ds.starts[j]), fr(ds.ends[j])
                cnt = min(e0 - s0, zb.shape[1], len(Z) - s0)
                if cnt <= 0:
                    j += 1
                    continue
                Z[s0:s0 + cnt] = zb[i, :cnt].numpy()
                have[s0:s0 + cnt] = True
                j += 1
        edge = np.zeros(n_frames, bool)
        for k in np.flatnonzero(np.diff(lab.astype(int)) != 0):
            edge[max(0, k - 3):k + 5] = True
        sel = (have & ~edge)[:len(Z)]
        lab_fit = lab[:len(Z)]
        A = np.concatenate([Z[sel].astype(np.float64), np.ones((int(sel.sum()), 1))], 1)
        G = A.T @ A + 100.0 * np.eye(spec.hidden + 1)
        G[-1, -1] -= 100.0
        wb = np.linalg.solve(G, A.T @ np.where(lab_fit, 3.5, -3.5)[sel])
        out_w = torch.tensor(wb[:-1], dtype=torch.float32)[None, :]
        out_b = torch.tensor(wb[-1:], dtype=torch.float32)
        with torch.no_grad():
            m.seg_model.output_layer.weight.copy_(out_w)
            m.seg_model.output_layer.bias.copy_(out_b)
        # pass 2: the reference pipeline with the calibrated checkpoint
        acc = None
        for i in range(inference_times):
            ds.fixed_length_segmentation(i)
            probs, _, _, _ = ns.evaluate.infer(m, loader(), torch.device("cpu"), False, "bce", None)
            out[f"probs_{i}"] = probs.copy() if probs_every == 1 else probs[::probs_every].astype(np.float32)
            acc = probs.copy() if acc is None else acc + probs
        acc /= inference_times
    out["probs_avg"] = acc if probs_every == 1 else acc[::probs_every].astype(np.float32)
    algos = {
        "dac": (ns.segment.pdac, dict(max_segment_length=16, min_segment_length=0.2, threshold=0.5)),
        "strm": (ns.segment.strm, dict(max_segment_length=18, min_segment_length=0.2, min_pause_length=0.2, threshold=0.5)),
        "pthr": (ns.segment.pthr, dict(max_segment_length=28, min_segment_length=0.2, max_lerp_range=4,
                                       min_lerp_range=0.4, threshold=0.1, moving_average_window=0.1)),
    }
    for tag, (fn, kw) in algos.items():
        segs = fn(acc, **kw)
        out[f"{tag}_bounds"] = np.array([[s.start, s.end] for s in segs], dtype=np.float64).reshape(-1, 2)
        content = ns.segment.update_yaml_content([], segs, "talk.wav")
        out[f"{tag}_yaml"] = np.array(yaml.dump(content, default_flow_style=True))
    np.savez_compressed(
        GOLD / f"{name}.npz",
        spec=np.array([spec.keep_layers, spec.adapter_layers, spec.head_layers, spec.head_heads]),
        seed=seed, audio_seed=audio_seed, n_samples=n_samples, inference_times=inference_times,
        batch_size=batch_size, duration_outframes=int(ds.duration_outframes), probs_every=probs_every,
        n_frames=len(acc), out_w=out_w.numpy(), out_b=out_b.numpy(),
        labels=np.packbits(lab) if probs_every > 1 else lab, **out)
    dec = (np.abs(acc - 0.5) > 0.4).mean()
    print(f"{name}: {len(acc)} frames, {dec:.3f} of them with p < 0.1 or p > 0.9, label agreement "
          f"{((acc > 0.5) == lab).mean():.4f}, " + ", ".join(f"{t}:{len(out[t + '_bounds'])}" for t in algos))


def _prob_tracks(rng, n):
    """probability-like test signals: smooth random walk through [0,1] with exact-zero runs"""
    steps = rng.normal(0, 0.08, n).cumsum()
    p = 1 / (1 + np.exp(-(np.sin(np.arange(n) / rng.uniform(20, 200)) * 2 + steps % 3 - 1.5)))
    k = rng.integers(0, 6)
    for _ in range(k):
        a = rng.integers(0, n)
        p[a: a + rng.integers(1, 400)] = 0.0
    if rng.random() < 0.3:
        p = np.round(p, 2)  # many exact ties
    return p


def gold_algos(ns, name, n_cases=48):
    """the segmentation algorithms and moving_average on their own (lib/segment.py)"""
    rng = np.random.default_rng(1234)
    out = {}
    cases = []
    for c in range(n_cases):
        n = int(rng.choice([1, 2, 7, 60, 999, 3351, 12000, 30000]))
        p = _prob_tracks(rng, n)
        if c == 0:
            p = np.zeros(500)          # nothing above threshold
        if c == 1:
            p = np.ones(4000) * 0.9    # never splittable
        kw_dac = dict(max_segment_length=float(rng.choice([4, 10, 16])), min_segment_length=0.2,
                      threshold=float(rng.choice([0.3, 0.5])))
        kw_strm = dict(max_segment_length=float(rng.choice([6, 18])), min_segment_length=0.2,
                       min_pause_length=float(rng.choice([0.1, 0.2])), threshold=0.5)
        kw_pthr = dict(max_segment_length=float(rng.choice([10, 28])), min_segment_length=0.2,
                       max_lerp_range=float(rng.choice([0, 4])), min_lerp_range=float(rng.choice([0, 0.4])),
                       threshold=float(rng.choice([0.1, 0.5])),
                       moving_average_window=float(rng.choice([0, 0.1, 0.5])))
        out[f"p_{c}"] = p
        w = int(rng.choice([1, 2, 5, 25, 300]))
        out[f"ma_{c}"] = ns.segment.moving_average(p, w)
        out[f"maw_{c}"] = w
        for tag, fn, kw in (("dac", ns.segment.pdac, kw_dac), ("strm", ns.segment.strm, kw_strm),
                            ("pthr", ns.segment.pthr, kw_pthr)):
            segs = fn(p, **kw)
            out[f"{tag}_bounds_{c}"] = np.array([[s.start, s.end] for s in segs], dtype=np.float64).reshape(-1, 2)
            out[f"{tag}_kw_{c}"] = np.array(yaml.dump(kw))
            out[f"{tag}_yaml_{c}"] = np.array(
                yaml.dump(ns.segment.update_yaml_content([], segs, "a.wav"), default_flow_style=True))
        cases.append(c)
    np.savez_compressed(GOLD / f"{name}.npz", n_cases=n_cases, **out)
    print(f"{name}: {n_cases} cases")


def gold_plan(ns, name):
    """window plans + frame indices of FixedSegmentationDatasetNoTarget for many durations"""
    out = {}
    rng = np.random.default_rng(7)
    durs = [32000, 40000, 320000, 320001, 351999, 352000, 352001, 640000, 1073234, 115_200_000] + \
        [int(x) for x in rng.integers(33000, 3_000_000, 20)]
    recs = []
    with tempfile.TemporaryDirectory() as td:
        for d in durs:
            wav = Path(td) / f"{d}.wav"
            with wave.open(str(wav), "wb") as w:
                w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
                w.writeframes(b"\0\0" * d)
            for it in (1, 2, 3, 4):
                ds = ns.dataset.FixedSegmentationDatasetNoTarget(wav, 20, it)
                for i in range(it):
                    ds.fixed_length_segmentation(i)
                    sf = [int(ds._inframes_to_outframes(s + 1e-6)) for s in ds.starts]
                    ef = [int(ds._inframes_to_outframes(e + 1e-6)) for e in ds.ends]
                    recs.append([d, it, i, int(ds.duration_outframes), list(map(int, ds.starts)),
                                 list(map(int, ds.ends)), sf, ef])
            wav.unlink()
    out["plans"] = np.array(json.dumps(recs))
    np.savez_compressed(GOLD / f"{name}.npz", **out)
    print(f"{name}: {len(recs)} plans")


# ---- dev-set scoring path (lib/evaluate.py:130-214 + FixedSegmentationDataset, lib/dataset.py:335-498)
EVAL_TALKS = [  # (id, n_samples, audio seed, true segments [start, end) in samples)
    ("talk_a", 1_073_234, 60, [(8_000, 150_000), (150_000, 310_000), (333_333, 640_001), (655_000, 700_000),
                               (905_000, 1_073_234)]),
    ("talk_b", 500_017, 61, [(0, 90_000), (120_000, 320_000), (320_500, 480_000)]),
]


EVAL_POS_WEIGHT = 0.4


def write_eval_corpus(root: Path):
    """wav files + the two tsv lists the reference's data preparation produces (only the columns the
    scoring path reads: talks id / path / total_frames, segments talk_id / start / end)"""
    talks, segs = ["\tid\tpath\ttotal_frames"], ["\ttalk_id\tstart\tend"]
    k = 0
    for i, (tid, n, seed, true) in enumerate(EVAL_TALKS):
        write_wav(root / f"{tid}.wav", synth.synthetic_audio(n, seed))
        talks.append(f"{i}\t{tid}\t{root / (tid + '.wav')}\t{n}")
        for a, b in true:
            segs.append(f"{k}\t{tid}\t{a}\t{b}")
            k += 1
    (root / "talks.tsv").write_text("\n".join(talks) + "\n")
    (root / "segments.tsv").write_text("\n".join(segs) + "\n")
    return root / "talks.tsv", root / "segments.tsv"


def gold_eval(ns, name, spec, seed, inference_times, batch_size):
    """FixedDataloaderGenerator -> infer (with targets) -> evaluate, all from the unmodified reference"""
    m, _ = build_reference_model(ns, spec, seed)
    loss_fn = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(EVAL_POS_WEIGHT), reduction="none")
    out = {}
    with tempfile.TemporaryDirectory() as td:
        talk_list, seg_list = write_eval_corpus(Path(td))
        gen = ns.dataset.FixedDataloaderGenerator(str(talk_list), str(seg_list), 20, batch_size, 0, inference_times)
        for tid, *_ in EVAL_TALKS:
            for i in range(inference_times):
                dl = gen.generate(tid, i)
                df = gen.dataset.fixed_segments_df
                out[f"{tid}_{i}_starts"] = df.start.to_numpy().astype(np.int64)
                out[f"{tid}_{i}_ends"] = df.end.to_numpy().astype(np.int64)
                out[f"{tid}_{i}_included"] = np.array(json.dumps(list(df.included)))
                items = [gen.dataset[k] for k in range(len(gen.dataset))]
                out[f"{tid}_{i}_targets"] = np.concatenate([t[1].numpy() for t in items])
                out[f"{tid}_{i}_target_lens"] = np.array([len(t[1]) for t in items])
                out[f"{tid}_{i}_frames"] = np.array([[t[2], t[3]] for t in items], dtype=np.int64)
                p, l, t, loss = ns.evaluate.infer(m, dl, torch.device("cpu"), False, "bce", None, loss_fn)
                out[f"{tid}_{i}_probs"] = p
                out[f"{tid}_{i}_talk_targets"] = t
                out[f"{tid}_{i}_loss"] = float(loss)
            out[f"{tid}_duration_outframes"] = int(gen.dataset.duration_outframes)
        # loss_fn as train.py builds it from conf/task/shas.yaml (:25-30, train.py:355-374); without one
        # the reference's evaluate() dies on an unbound `eval_loss` (lib/evaluate.py:211)
        res = ns.evaluate.evaluate(gen, m, torch.device("cpu"), False, "bce", None, loss_fn)
    out["metrics"] = np.array(json.dumps({k: float(v) for k, v in res.items()}))
    out["pos_weight"] = EVAL_POS_WEIGHT
    np.savez_compressed(
        GOLD / f"{name}.npz",
        spec=np.array([spec.keep_layers, spec.adapter_layers, spec.head_layers, spec.head_heads]),
        seed=seed, inference_times=inference_times, batch_size=batch_size, **out)
    print(f"{name}: metrics {res}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    from oracle import ref_shims

    ns = ref_shims.install()
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    jobs = {
        "plan": lambda: gold_plan(ns, "plan"),
        "algos": lambda: gold_algos(ns, "algos"),
        # TINY (2 layers, last one with adapter): ragged batch incl. a window whose out_mask is one
        # frame longer than the encoder mask and a short one
        "tiny_batch": lambda: gold_batch(ns, "tiny_batch", synth.TINY, 0, [64000, 113234, 48000], 10),
        # GroupNorm feature extractor (HF feat_extract_norm="group"): ragged batches with / without conv bias;
        # the GroupNorm statistics run over the PADDED row, so the padding of the batch matters
        "tiny_gn_batch": lambda: gold_batch(ns, "tiny_gn_batch", synth.TINY_GN, 0, [64000, 113234, 48000], 70),
        "tiny_gn_nobias_batch": lambda: gold_batch(ns, "tiny_gn_nobias_batch", synth.TINY_GN_NOBIAS, 0, [160000, 90001], 71),
        # wav2vec2-large-960h-like architecture: GroupNorm extractor, no conv bias, post-LN encoder layers (3 kept)
        "tiny_postln_batch": lambda: gold_batch(ns, "tiny_postln_batch", synth.TINY_POSTLN, 0, [96000, 57011], 72),
        # BASELINE.json configs[0]: middle (0/16), frozen encoder, single 20 s window
        "middle_window": lambda: gold_batch(ns, "middle_window", synth.MIDDLE, 0, [320000], 20),
        # middle+half (8/16): adapters in layers 8..15, ragged pair
        "middle_half_batch": lambda: gold_batch(ns, "middle_half_batch", synth.MIDDLE_HALF, 0, [320000, 200000], 30),
        # large (24/24) + 24 adapters: the headline model, 2 x 20 s
        "large_batch": lambda: gold_batch(ns, "large_batch", synth.LARGE_ALL, 0, [320000, 320000], 40),
        # whole-talk path with overlapped tilings (configs[3] shape): 67 s + odd samples
        "tiny_talk": lambda: gold_talk(ns, "tiny_talk", synth.TINY, 0, 1_073_234, 50, 2, 3),
        "tiny_talk_x1": lambda: gold_talk(ns, "tiny_talk_x1", synth.TINY, 0, 753_234, 51, 1, 14),
        # decisive (trained-like) probabilities: the boundary-identity fixtures
        "speech_talk": lambda: gold_talk_decisive(ns, "speech_talk", synth.TINY, 0, 16000 * 400 + 1234, 60, 1, 14),
        "speech_talk_x2": lambda: gold_talk_decisive(ns, "speech_talk_x2", synth.TINY, 0, 16000 * 200 + 777, 61, 2, 3),
        # the configurations BASELINE.json names, with decisive probabilities: large (24/24) + 24 adapters on a
        # 420 s talk (one tiling), middle+half (8/16) with two overlapped tilings (configs[3] as written)
        "speech_talk_large": lambda: gold_talk_decisive(ns, "speech_talk_large", synth.LARGE_ALL, 0,
                                                        16000 * 420 + 345, 62, 1, 14),
        "speech_talk_mh_x2": lambda: gold_talk_decisive(ns, "speech_talk_mh_x2", synth.MIDDLE_HALF, 0,
                                                        16000 * 300 + 1234, 63, 2, 14),
        # configs[0]'s model, middle (0/16) with the frozen encoder (no adapters), on a 260 s talk
        "speech_talk_middle": lambda: gold_talk_decisive(ns, "speech_talk_middle", synth.MIDDLE, 0,
                                                         16000 * 260 + 4321, 65, 1, 14),
        # configs[4]: 2 h single stream (360 windows, 359 640 frames): the reference's full yaml for
        # dac / strm / pthr(+moving average); probabilities stored as an every-8th-frame fp32 sample
        "speech_talk_2h": lambda: gold_talk_decisive(ns, "speech_talk_2h", synth.TINY, 0, 16000 * 7200, 64, 1, 14,
                                                     fit_batches=4, probs_every=8),
        # dev-set scoring: two talks with labelled segments, two tilings
        "tiny_eval": lambda: gold_eval(ns, "tiny_eval", synth.TINY, 0, 2, 3),
        "tiny_eval_x1": lambda: gold_eval(ns, "tiny_eval_x1", synth.TINY, 0, 1, 14),
    }
    for k, fn in jobs.items():
        if args.only is None or args.only == k:
            fn()


if __name__ == "__main__":
    main()
