"""CLI / config boundary of segment.py and inference.py (reference segment.py:159-177,
inference.py:156-189, conf/*.yaml): the conf tree is the reference's, and the reference's own
command lines (README.md:73-79,105-160) compose to the same values and output locations."""
import os
import sys
from pathlib import Path

import pytest
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from wav2vecsegmenter_b200 import config as cfglib  # noqa: E402

CONF = ROOT / "conf"


def test_conf_tree_has_the_reference_files_and_keys():
    seg = yaml.safe_load((CONF / "segment.yaml").read_text())
    assert seg["defaults"] == ["_self_", {"algorithm": "pthr"}, {"infer_data": "mustc_ende_tst-COMMON"}]
    assert [k for k in seg if k not in ("defaults", "hydra")] == [
        "work_dir", "ckpt_path", "config_path", "output_dir", "cust_seg_yaml", "batch_size",
        "inference_segment_length", "inference_times"]
    inf = yaml.safe_load((CONF / "inference.yaml").read_text())
    assert inf["defaults"] == ["_self_", {"algorithm": "dac"}, {"infer_data": "mustc_ende_tst-COMMON"}]
    for k in ("outputs", "base_cfg", "ckpt", "log_wandb", "project_name", "cust_seg_yaml", "st_metrics"):
        assert k in inf
    assert inf["outputs"] == "???" and inf["ckpt"] == "???" and inf["base_cfg"] == "${outputs}/.hydra"
    assert sorted(p.stem for p in (CONF / "algorithm").glob("*.yaml")) == ["dac", "dac_logits", "pthr", "strm"]
    assert yaml.safe_load((CONF / "task" / "shas.yaml").read_text())["model"]["_target_"] == "lib.models.SHAS"
    ref = Path("/root/reference/conf")
    if ref.exists():   # build container only: byte identity with the reference's files
        for f in CONF.rglob("*.yaml"):
            assert f.read_bytes() == (ref / f.relative_to(CONF)).read_bytes(), f


def test_segment_command_line(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    train_cfg = tmp_path / "run" / ".hydra" / "config.yaml"
    train_cfg.parent.mkdir(parents=True)
    shas = yaml.safe_load((CONF / "task" / "shas.yaml").read_text())
    shas["model"].update(finetune_wav2vec=True, wav2vec_keep_layers=24, wav2vec_ft_layers=24)   # README.md:73-79
    train_cfg.write_text(yaml.safe_dump({"exp_name": "lna_l24_ft24", "batch_size": 4, "task": shas}))
    out = tmp_path / "seg_out"
    cfg = cfglib.compose(CONF, "segment", [f"ckpt_path={tmp_path}/run/ckpts/best.pt", f"config_path={train_cfg}",
                                            f"output_dir={out}", "algorithm=strm", "algorithm.threshold=0.4",
                                            "inference_times=2"])
    assert cfg.algorithm.tag == "strm" and cfg.algorithm.threshold == 0.4 and cfg.algorithm.min_pause_length == 0.2
    assert cfg.batch_size == 14 and cfg.inference_times == 2 and cfg.inference_segment_length == 20
    assert cfg.work_dir == str(tmp_path)
    assert cfg.infer_data.wav_dir == f"{tmp_path}/data/corpus/MuST-C/v2.0_IWSLT2022/en-de/data/tst-COMMON/wav"
    merged = cfglib.merge(cfglib.load(cfg.config_path), cfg)        # reference segment.py:161-163
    assert merged.task.model.wav2vec_keep_layers == 24 and merged.task.model.finetune_wav2vec is True
    assert merged.batch_size == 14                                   # the segment config wins over the saved one
    assert merged.exp_name == "lna_l24_ft24"
    # run dir = ${output_dir}/${hydra.job.override_dirname}: sorted overrides minus exclude_keys
    d = cfglib.run_dir(cfg)
    assert d == out / "algorithm.threshold=0.4,algorithm=strm,inference_times=2"
    assert yaml.safe_load((d / ".hydra" / "overrides.yaml").read_text())[3] == "algorithm=strm"
    assert cfglib.to_object(cfg.algorithm) == {"tag": "strm", "max_segment_length": 18, "min_segment_length": 0.2,
                                               "min_pause_length": 0.2, "threshold": 0.4}


def test_inference_command_line(tmp_path, monkeypatch):
    """README.md:112-121 (inference_st_pipe.py shares conf keys with inference.py)"""
    monkeypatch.chdir(tmp_path)
    exp = tmp_path / "outputs" / "large+all"
    (exp / ".hydra").mkdir(parents=True)
    shas = yaml.safe_load((CONF / "task" / "shas.yaml").read_text())
    (exp / ".hydra" / "config.yaml").write_text(yaml.safe_dump({"exp_name": "lna_l24_ft24", "task": shas}))
    args = [f"outputs={exp}", "ckpt=epoch-15_best_eval_f1.pt", "log_wandb=False", "infer_data=mustc_ende_dev",
            "batch_size=14", "algorithm=dac", "algorithm.max_segment_length=16", "algorithm.threshold=0.5"]
    cfg = cfglib.compose(CONF, "inference", args)
    assert cfg.base_cfg == f"{exp}/.hydra" and cfg.log_wandb is False
    assert cfg.algorithm.tag == "dac" and cfg.algorithm.max_segment_length == 16
    assert cfg.infer_data.orig_seg_yaml.endswith("/dev/txt/dev.yaml")
    assert cfg.fairseq_root == f"{tmp_path}/tools/fairseq"           # relative interpolation ${.work_dir}
    assert cfg.st_metrics == ["bleu", "bertscore"] and cfg.group is None
    import inference

    merged = cfglib.merge(cfglib.load(Path(cfg.base_cfg) / "config.yaml"), cfg)
    assert inference.checkpoint_path(merged) == f"{exp}/lna_l24_ft24/ckpts/epoch-15_best_eval_f1.pt"
    d = cfglib.run_dir(cfg)
    assert d == exp / "infer_outputs" / ("algorithm.max_segment_length=16,algorithm.threshold=0.5,algorithm=dac,"
                                         "ckpt=epoch-15_best_eval_f1.pt,infer_data=mustc_ende_dev,log_wandb=False")


def test_default_algorithms_and_mandatory_values(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    seg = cfglib.compose(CONF, "segment", [])
    assert seg.algorithm.tag == "pthr" and seg.algorithm.moving_average_window == 0.1
    with pytest.raises(cfglib.MissingMandatoryValue):
        seg.ckpt_path                                    # `???` raises on access, not at load
    inf = cfglib.compose(CONF, "inference", ["algorithm=dac_logits"])
    assert inf.algorithm.tag == "dac_logits"             # composes; lib.segment raises at use
    with pytest.raises(cfglib.MissingMandatoryValue):
        inf.base_cfg                                     # ${outputs} is still ???
    with pytest.raises(SystemExit):
        cfglib.compose(CONF, "inference", ["algorithm=nope"])
    from lib.segment import pdac_with_logits

    with pytest.raises(NotImplementedError):
        pdac_with_logits(None, None, None, 18, 0.2)


def test_env_resolver(tmp_path, monkeypatch):
    (tmp_path / "c.yaml").write_text("a: ${oc.env:W2V_TEST_VAR,fallback}\nb:\n  c: ${..a}/x\n  d: ${.c}/y\n")
    assert cfglib.compose(tmp_path, "c", []).b.d == "fallback/x/y"
    monkeypatch.setenv("W2V_TEST_VAR", "v")
    assert cfglib.compose(tmp_path, "c", ["+e.f=3"]).a == "v"
    assert cfglib.compose(tmp_path, "c", ["+e.f=3"]).e.f == 3
