"""BASELINE.json configs[2] as a test: the whole-talk path sharded over 2 GPUs (windows sharded, one NCCL
all_gather of probability rows) reproduces the single-GPU result bit for bit. Spawns 2 ranks with torchrun;
skipped on a box with fewer than 2 GPUs (the driver's 1-GPU run). scripts/check_multigpu.py is the worker
(the 10 h / 8-GPU measurement of profiles/check_10h_n8_*.json uses the same script with --model large)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharding_is_bit_identical():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "scripts" / "check_multigpu.py"), "--talks", "5", "--seconds", "170"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["world"] == 2 and rec["bit_identical_to_single_gpu"] is True
    assert rec["rank0_only_results_identical"] is True


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_segment_cli_two_ranks_same_yaml(tmp_path):
    """segment.py under torchrun (windows of all talks sharded over 2 ranks, rank 0 segments and writes) produces
    the same custom_segments.yaml, byte for byte, as the single-GPU run"""
    import wave as wavmod

    import yaml

    sys.path.insert(0, str(ROOT))
    from wav2vecsegmenter_b200 import synth

    spec = synth.TINY
    names = []
    for i, n in enumerate((500_123, 333_000, 41_000)):
        x, _ = synth.speech_like_audio(n, 40 + i)
        pcm = torch.round(x * 32767.0).clamp(-32768, 32767).to(torch.int16).numpy()
        with wavmod.open(str(tmp_path / f"t{i}.wav"), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(pcm.tobytes())
        names.append(f"t{i}.wav")
    torch.save({"state_dict": synth.random_state_dict(spec, 0)}, tmp_path / "ckpt.pt")
    (tmp_path / "train.yaml").write_text(yaml.dump({"exp_name": "t", "task": {
        "autoregression": False, "vocab": None, "loss": {"tag": "bce"}, "model": {
            "_target_": "lib.models.SHAS", "wav2vec_model_name": "x", "wav2vec_keep_layers": spec.keep_layers,
            "finetune_wav2vec": True, "wav2vec_ft_layers": spec.adapter_layers, "finetune_w2v_feat_enc": False,
            "finetune_w2v_ffn": False, "ffn_adapter": True, "n_transformer_enc_layers": 1,
            "n_transformer_enc_heads": 8, "init_dropout": 0.1}}}))
    (tmp_path / "orig.yaml").write_text(yaml.dump([{"wav": n, "offset": 0.0, "duration": 1.0} for n in names]))
    common = [f"ckpt_path={tmp_path / 'ckpt.pt'}", f"config_path={tmp_path / 'train.yaml'}", "algorithm=pthr",
              "infer_data=toy", f"infer_data.wav_dir={tmp_path}", f"infer_data.orig_seg_yaml={tmp_path / 'orig.yaml'}",
              "inference_times=2"]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    outs = []
    for tag, launcher in (("one", [sys.executable]),
                          ("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29534"])):
        out = tmp_path / tag
        r = subprocess.run(launcher + [str(ROOT / "segment.py"), f"output_dir={out}"] + common, capture_output=True,
                           text=True, env=env, timeout=600, cwd=str(ROOT))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        files = list(out.rglob("custom_segments.yaml"))
        assert len(files) == 1, files
        outs.append(files[0].read_text())
    assert outs[0] == outs[1] and outs[0].startswith("[{duration:")
    assert {rec["wav"] for rec in yaml.safe_load(outs[0])} == set(names)
