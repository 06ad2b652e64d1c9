"""segment.py / inference.py end to end through their command lines on the GPU (reference
segment.py:159-177, inference.py:156-189), including the frozen-encoder load order
(`.to(device)` BEFORE `seg_model.load_state_dict`, reference segment.py:41-51)."""
import sys
import wave
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from wav2vecsegmenter_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _fresh(*names):
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.") or k in names]:
        del sys.modules[k]


def _write_wav(path, n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g) * (0.02 + 0.3 * torch.rand(n // 8000 + 1, generator=g).repeat_interleave(8000)[:n])
    pcm = (x.clamp(-1, 1) * 32767).to(torch.int16).numpy()
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(pcm.tobytes())


def _model_cfg(spec, finetune):
    return {"_target_": "lib.models.SHAS", "wav2vec_model_name": "facebook/wav2vec2-xls-r-300m",
            "wav2vec_keep_layers": spec.keep_layers, "finetune_wav2vec": finetune,
            "wav2vec_ft_layers": spec.adapter_layers if finetune else 99, "finetune_w2v_feat_enc": False,
            "finetune_w2v_ffn": False, "ffn_adapter": True, "n_transformer_enc_layers": 1,
            "n_transformer_enc_heads": 8, "init_dropout": 0.1}


def test_segment_main_writes_into_the_run_dir(tmp_path, monkeypatch):
    _fresh("segment")
    import segment as cli

    spec = synth.TINY
    _write_wav(tmp_path / "a.wav", 500_123, 3)
    torch.save({"state_dict": synth.random_state_dict(spec, 0)}, tmp_path / "ckpt.pt")
    (tmp_path / "train.yaml").write_text(yaml.dump({"exp_name": "t", "task": {
        "autoregression": False, "vocab": None, "model": _model_cfg(spec, True), "loss": {"tag": "bce"}}}))
    (tmp_path / "orig.yaml").write_text(yaml.dump([{"wav": "a.wav", "offset": 0.0, "duration": 1.0}]))
    out = tmp_path / "out"
    cli.main([f"ckpt_path={tmp_path / 'ckpt.pt'}", f"config_path={tmp_path / 'train.yaml'}", f"output_dir={out}",
              "algorithm=dac", "algorithm.max_segment_length=10", "infer_data=toy", f"infer_data.wav_dir={tmp_path}",
              f"infer_data.orig_seg_yaml={tmp_path / 'orig.yaml'}"])
    run = out / (f"algorithm.max_segment_length=10,algorithm=dac,infer_data.orig_seg_yaml={tmp_path / 'orig.yaml'},"
                 f"infer_data.wav_dir={tmp_path},infer_data=toy")
    text = (run / "custom_segments.yaml").read_text()
    assert text.startswith("[{duration:")
    recs = yaml.safe_load(text)
    assert recs and all(r["wav"] == "a.wav" and r["duration"] <= 10.0 + 1e-6 for r in recs)
    assert (run / ".hydra" / "config.yaml").exists()


def test_inference_main_frozen_encoder(tmp_path, monkeypatch):
    """reference inference.py flow with finetune_wav2vec=False: the checkpoint holds ONLY the head
    (train.py:596-604), the encoder is the pretrained one (random-init stand-in here), and the model
    is moved to the device before the head is loaded."""
    monkeypatch.setenv("W2VSEG_RANDOM_INIT", "1")
    monkeypatch.setenv("W2VSEG_SEED", "5")
    _fresh("segment", "inference")
    import inference as cli
    from wav2vecsegmenter_b200.pipeline import TalkRunner
    from wav2vecsegmenter_b200.engine import SFCEngine
    from lib.dataset import read_wav
    from lib.segment import pdac

    spec = synth.ModelSpec(keep_layers=2, adapter_layers=0)
    exp = tmp_path / "outputs" / "run1"
    (exp / ".hydra").mkdir(parents=True)
    (exp / "middle" / "ckpts").mkdir(parents=True)
    (exp / ".hydra" / "config.yaml").write_text(yaml.dump({"exp_name": "middle", "task": {
        "autoregression": False, "vocab": None, "model": _model_cfg(spec, False), "loss": {"tag": "bce"}}}))
    head = synth.random_head_state_dict(spec, 9, 2.0, prefix="")
    torch.save({"state_dict": head}, exp / "middle" / "ckpts" / "best.pt")
    wavs = tmp_path / "wav"
    wavs.mkdir()
    _write_wav(wavs / "b.wav", 420_000, 2)
    _write_wav(wavs / "a.wav", 333_333, 1)
    cli.main([f"outputs={exp}", "ckpt=best.pt", "log_wandb=False", "infer_data=toy", f"infer_data.wav_dir={wavs}",
              "inference_times=2"])
    run = exp / "infer_outputs" / f"ckpt=best.pt,infer_data.wav_dir={wavs},infer_data=toy,inference_times=2,log_wandb=False"
    recs = yaml.safe_load((run / "custom_segments.yaml").read_text())
    assert [r["wav"] for r in recs] == sorted(r["wav"] for r in recs) and {r["wav"] for r in recs} == {"a.wav", "b.wav"}

    # the same records from the engine driven directly with the same weights (default algorithm: dac)
    sd = {k: v for k, v in synth.random_state_dict(spec, seed=5).items() if k.startswith("wav2vec_model.model.")}
    sd.update({"seg_model." + k: v for k, v in head.items()})
    eng = SFCEngine(spec, "cuda:0")
    eng.load_state_dict(sd)
    expect = []
    for name in ("a.wav", "b.wav"):
        w, _ = read_wav(wavs / name)
        res = TalkRunner(eng, batch_size=14, inference_times=2).run([w])[0]
        expect += [s for s in pdac(res.probs, 16, 0.2, 0.5)]
    assert len(expect) == len(recs)
    for s, r in zip(expect, recs):
        assert abs(r["offset"] - round(s.offset, 6)) < 1e-9 and abs(r["duration"] - round(s.duration, 6)) < 1e-9


def test_frozen_encoder_missing_head_is_loud():
    """no CPU fallback and no silent partial model: a forward without the head weights fails"""
    import os

    os.environ["W2VSEG_RANDOM_INIT"] = "1"
    try:
        _fresh()
        from lib.models import SHAS
        from wav2vecsegmenter_b200._native import W2VSegError

        m = SHAS("x", 2, False, 99, False, False, True, 1, 8, 0.1).to("cuda:0")   # must not raise
        with pytest.raises(W2VSegError):
            m.engine
    finally:
        del os.environ["W2VSEG_RANDOM_INIT"]
