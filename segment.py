#!/usr/bin/env python
"""Segment every wav of a corpus with a trained SFC model and write custom_segments.yaml
(drop-in for the reference's segment.py: same config keys, same output).

    python segment.py ckpt_path=... config_path=... output_dir=... [algorithm=dac|strm|pthr] \
        [infer_data=...] [inference_times=N] [batch_size=14]

Config: conf/segment.yaml (byte-identical to the reference's) composed by Hydra when it is
importable (`@hydra.main(config_path="conf", config_name="segment")`, reference segment.py:159),
otherwise by wav2vecsegmenter_b200.config (same defaults list / overrides / interpolation / run dir
`${output_dir}/${hydra.job.override_dirname}`). Like the reference, `custom_segments.yaml` lands in
the job's run directory (reference segment.py:175-176 writes relative to Hydra's cwd).

Mechanism (new): each wav is decoded once, all windows of all tilings of a talk go through the
CUDA SFC forward in device batches (wav2vecsegmenter_b200.pipeline.TalkRunner), the talk vector
is assembled / NaN-filled / averaged on the GPU, and only the final per-frame probabilities come
back to the host for the segmentation algorithm. With torchrun, windows are sharded across ranks
and rank 0 writes the yaml.
"""
from __future__ import annotations

import itertools
import logging
import os
import sys
from pathlib import Path

import torch
import yaml

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from lib.dataset import read_wav  # noqa: E402
from lib.segment import pdac, pdac_with_logits, pthr, strm, update_yaml_content  # noqa: E402
from wav2vecsegmenter_b200 import config as cfglib  # noqa: E402
from wav2vecsegmenter_b200.pipeline import TalkRunner  # noqa: E402

logger = logging.getLogger("segment")


def load_model(config, device, ckpt_path=None):
    """build SHAS from config.task.model and load the checkpoint (reference segment.py:41-52)"""
    model = cfglib.instantiate(config.task.model).to(device)
    checkpoint = torch.load(ckpt_path if ckpt_path is not None else config.ckpt_path, map_location="cpu")
    if config.task.model.finetune_wav2vec:
        model.load_state_dict(checkpoint["state_dict"])
    else:
        model.seg_model.load_state_dict(checkpoint["state_dict"])
    model.eval()
    return model


def run_algorithm(config, probs, logits=None, vocab=None):
    algo_conf = cfglib.to_object(config.algorithm)
    algorithm = algo_conf.pop("tag")
    if algorithm == "dac":
        return pdac(probs, **algo_conf)
    if algorithm == "dac_logits":
        return pdac_with_logits(probs, logits, vocab, **algo_conf)
    if algorithm == "strm":
        return strm(probs, **algo_conf)
    if algorithm == "pthr":
        return pthr(probs, **algo_conf)
    raise ValueError(f"unknown algorithm tag '{algorithm}'")


def wav_names(config):
    with open(config.infer_data.orig_seg_yaml, "r") as f:
        seg_yaml = yaml.load(f, Loader=yaml.SafeLoader)
    return [name for name, _ in itertools.groupby(seg_yaml, key=lambda x: x["wav"])]


def generate(config, wav_paths=None, ckpt_path=None) -> list:
    if torch.cuda.device_count() == 0:
        raise RuntimeError("segment.py (B200 build) needs a CUDA device: there is no CPU path")
    if config.task.get("vocab"):
        raise NotImplementedError("vocabulary (ce/ssl) variants are outside the accelerated SFC path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    group = None
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    device = torch.device("cuda", local)
    model = load_model(config, device, ckpt_path)
    runner = TalkRunner(model.engine, batch_size=config.batch_size,
                        segment_sec=config.inference_segment_length,
                        inference_times=config.inference_times, dist_group=group)
    if wav_paths is None:
        wav_paths = [Path(config.infer_data.wav_dir) / n for n in wav_names(config)]
    yaml_content = []

    def waves():
        for wav_path in wav_paths:
            wave, sr = read_wav(wav_path)
            assert sr == 16000, "Audio needs to have sample rate of 16000"
            yield wave

    rank0 = int(os.environ.get("RANK", "0")) == 0
    if world == 1:
        # one GPU: pipelined over talks — decode + H2D of the next wav overlap the forward of the current one
        for wav_path, result in zip(wav_paths, runner.run_stream(waves())):
            segments = run_algorithm(config, result.probs)
            yaml_content = update_yaml_content(yaml_content, segments, Path(wav_path).name)
    else:
        # several GPUs: shard the windows of MANY talks at once (per-talk sharding would leave every rank with a
        # handful of windows and one gather per talk): groups of talks of up to W2VSEG_GROUP_SECONDS of audio
        # (default 2 h = 460 MB of host samples), one all_gather per group, rank 0 assembles, segments and writes
        budget = float(os.environ.get("W2VSEG_GROUP_SECONDS", "7200")) * 16000
        group_paths, group_waves, total = [], [], 0

        def flush():
            nonlocal yaml_content, group_paths, group_waves, total
            if group_waves:
                results = runner.run(group_waves, results_on=0)
                if rank0:
                    for wp, res in zip(group_paths, results):
                        segments = run_algorithm(config, res.probs)
                        yaml_content = update_yaml_content(yaml_content, segments, Path(wp).name)
            group_paths, group_waves, total = [], [], 0

        for wav_path, wave in zip(wav_paths, waves()):
            if group_waves and total + len(wave) > budget:
                flush()
            group_paths.append(wav_path)
            group_waves.append(wave)
            total += len(wave)
        flush()
    del model
    torch.cuda.empty_cache()
    return yaml_content


def _run(config, results_dir: Path) -> None:
    """reference segment.py:160-177 after config composition"""
    if config.config_path is not None:
        prev_cfg = cfglib.load(config.config_path) if isinstance(config, cfglib.Cfg) else None
        if prev_cfg is None:
            from omegaconf import OmegaConf  # Hydra path

            config = OmegaConf.merge(OmegaConf.load(config.config_path), config)
        else:
            config = cfglib.merge(prev_cfg, config)
    logger.info("Output directory : [%s]", config.output_dir)
    yaml_content = generate(config)
    logger.info("Number of segments: %d", len(yaml_content))
    if int(os.environ.get("RANK", "0")) == 0:
        target = results_dir / config.cust_seg_yaml
        with open(target, "w") as f:
            yaml.dump(yaml_content, f, default_flow_style=True)
        logger.info("Saved to [%s].", target)


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    try:
        import hydra  # noqa: F401
    except ImportError:
        hydra = None
    if hydra is not None and argv is None:
        # Hydra changes into hydra.run.dir before calling the task function (as for the reference)
        hydra.main(config_path="conf", config_name="segment")(lambda config: _run(config, Path(os.getcwd())))()
        return
    config = cfglib.compose(ROOT / "conf", "segment", list(sys.argv[1:] if argv is None else argv))
    _run(config, cfglib.run_dir(config))


if __name__ == "__main__":
    main()
