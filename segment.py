#!/usr/bin/env python
"""Segment every wav of a corpus with a trained SFC model and write custom_segments.yaml
(drop-in for the reference's segment.py: same config keys, same output).

    python segment.py ckpt_path=... config_path=... output_dir=... [algorithm=dac|strm|pthr] \
        [infer_data=...] [inference_times=N] [batch_size=14]

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


def load_model(config, device):
    """build SHAS from config.task.model and load the checkpoint (reference segment.py:41-52)"""
    model = cfglib.instantiate(dict(config.task.model)).to(device)
    checkpoint = torch.load(config.ckpt_path, map_location="cpu")
    if config.task.model.finetune_wav2vec:
        model.load_state_dict(checkpoint["state_dict"])
    else:
        model.seg_model.load_state_dict(checkpoint["state_dict"])
    model.eval()
    return model


def run_algorithm(config, probs, logits=None, vocab=None):
    algo_conf = dict(config.algorithm)
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


def generate(config, wav_paths=None) -> list:
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
    model = load_model(config, device)
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

    # pipelined over talks: decode + H2D of the next wav overlap the forward of the current one
    for wav_path, result in zip(wav_paths, runner.run_stream(waves())):
        segments = run_algorithm(config, result.probs)
        yaml_content = update_yaml_content(yaml_content, segments, Path(wav_path).name)
    del model
    torch.cuda.empty_cache()
    return yaml_content


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    config = cfglib.compose(ROOT / "conf", "segment", list(sys.argv[1:] if argv is None else argv))
    out_dir = Path(config.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    logger.info("Output directory : [%s]", out_dir)
    yaml_content = generate(config)
    logger.info("Number of segments: %d", len(yaml_content))
    if int(os.environ.get("RANK", "0")) == 0:
        target = out_dir / config.cust_seg_yaml
        with open(target, "w") as f:
            yaml.dump(yaml_content, f, default_flow_style=True)
        logger.info("Saved to [%s].", target)


if __name__ == "__main__":
    main()
