#!/usr/bin/env python
"""Like segment.py but for an experiment directory: segments every *.wav under
infer_data.wav_dir with outputs/<exp_name>/ckpts/<ckpt> (drop-in for the reference's
inference.py:26-128; the ST evaluation that follows in inference_st_pipe.py is out of scope)."""
from __future__ import annotations

import logging
import sys
from pathlib import Path

import yaml

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import segment as seg  # noqa: E402
from wav2vecsegmenter_b200 import config as cfglib  # noqa: E402


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    args = list(sys.argv[1:] if argv is None else argv)
    config = cfglib.compose(ROOT / "conf", "inference", args)
    if config.get("exp_name") and config.get("ckpt"):
        config["ckpt_path"] = str(Path(config.get("outputs_dir", "outputs")) / config.exp_name / "ckpts" / config.ckpt)
    wavs = sorted(Path(config.infer_data.wav_dir).glob("*.wav"))
    content = seg.generate(config, wav_paths=wavs)
    out_dir = Path(config.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    with open(out_dir / config.cust_seg_yaml, "w") as f:
        yaml.dump(content, f, default_flow_style=True)


if __name__ == "__main__":
    main()
