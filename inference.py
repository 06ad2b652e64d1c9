#!/usr/bin/env python
"""Segment every *.wav under infer_data.wav_dir with a checkpoint of a training run (drop-in for the
reference's inference.py: same config keys, same output location).

    python inference.py outputs=<training run dir> ckpt=<file under ckpts/> [log_wandb=False] \\
        [infer_data=mustc_ende_dev] [batch_size=14] [algorithm=dac algorithm.max_segment_length=16 ...]

Kept from the reference (inference.py:26-189, conf/inference.yaml): the keys `outputs`, `ckpt`,
`base_cfg` (default `${outputs}/.hydra`, whose `config.yaml` — the training run's saved config with
`task.model` and `exp_name` — is merged underneath), the checkpoint path
`<outputs>/<exp_name>/ckpts/<ckpt>` (inference.py:46-49), the default algorithm `dac`, and the run
directory `${outputs}/infer_outputs/${hydra.job.override_dirname}` into which
`custom_segments.yaml` is written (inference.py:164-165,188-189). The ST evaluation that follows in
inference_st_pipe.py is out of scope (external fairseq model); its keys are accepted and ignored.
Config composition: Hydra when importable (`@hydra.main(config_path="conf",
config_name="inference")`), else wav2vecsegmenter_b200.config.
"""
from __future__ import annotations

import logging
import os
import sys
from pathlib import Path

import yaml

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import segment as seg  # noqa: E402
from wav2vecsegmenter_b200 import config as cfglib  # noqa: E402

logger = logging.getLogger("inference")


def checkpoint_path(config) -> str:
    """reference inference.py:46-49"""
    return "/".join([str(config.outputs), str(config.exp_name), "ckpts", str(config.ckpt)])


def _run(config, results_path: Path) -> None:
    """reference inference.py:157-189 after config composition"""
    if config.base_cfg is not None:
        if isinstance(config, cfglib.Cfg):
            config = cfglib.merge(cfglib.load(Path(config.base_cfg) / "config.yaml"), config)
        else:
            from omegaconf import OmegaConf  # Hydra path

            config = OmegaConf.merge(OmegaConf.load(Path(config.base_cfg) / "config.yaml"), config)
    rank0 = int(os.environ.get("RANK", "0")) == 0
    run = None
    if config.log_wandb and rank0:
        try:
            import wandb

            run = wandb.init(project=config.project_name, group=config.group,
                             name="/".join([str(config.exp_name), results_path.name]), tags=config.tags,
                             notes=config.notes, dir=str(results_path))
        except ImportError:
            logger.warning("log_wandb=True but wandb is not installed: continuing without logging")
    wavs = sorted(Path(config.infer_data.wav_dir).glob("*.wav"))      # inference.py:69
    yaml_content = seg.generate(config, wav_paths=wavs, ckpt_path=checkpoint_path(config))
    if run is not None:
        import wandb

        wandb.log({"n_segments": len(yaml_content)}, step=0)
        wandb.finish()
    if rank0:
        with open(results_path / config.cust_seg_yaml, "w") as f:
            yaml.dump(yaml_content, f, default_flow_style=True)
        logger.info("Saved to [%s].", results_path / config.cust_seg_yaml)


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    try:
        import hydra  # noqa: F401
    except ImportError:
        hydra = None
    if hydra is not None and argv is None:
        hydra.main(config_path="conf", config_name="inference")(lambda config: _run(config, Path(os.getcwd())))()
        return
    config = cfglib.compose(ROOT / "conf", "inference", list(sys.argv[1:] if argv is None else argv))
    _run(config, cfglib.run_dir(config))


if __name__ == "__main__":
    main()
