"""BASELINE.json configs[4]: long-form single-stream audio (default 2 h = 360 windows of 20 s) through the
public host API, wall clock: TalkRunner.run (pageable numpy samples in, per-frame probabilities out) plus
the pTHR moving average and the pDAC / pSTRM segmentation on the host.
W2VSEG_UPLOAD_CHUNK=1000000000 reproduces the single up-front host->device copy for comparison."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "lib"))
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402
from wav2vecsegmenter_b200.pipeline import TalkRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hours", type=float, default=2.0)
ap.add_argument("--model", default="large", choices=["tiny", "large"])
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

spec = synth.LARGE_ALL if args.model == "large" else synth.TINY
eng = SFCEngine(spec)
eng.load_state_dict(synth.random_state_dict(spec, 0))
n = int(args.hours * 3600 * 16000)
rng = np.random.default_rng(0)
wave = (rng.standard_normal(n, dtype=np.float32) * 0.1).astype(np.float32)
runner = TalkRunner(eng, batch_size=14, segment_sec=20, inference_times=1)
runner.run([wave[: 16000 * 600]])      # warm-up (allocator, kernels)
torch.cuda.synchronize()
times = []
for _ in range(args.reps):
    t0 = time.perf_counter()
    res = runner.run([wave])[0]
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
import lib.segment as seg  # noqa: E402

t0 = time.perf_counter()
seg.pdac(res.probs, 16, 0.2, 0.5)
t_dac = time.perf_counter() - t0
best = min(times)
print(json.dumps({"config": f"long-form {args.hours} h single stream, {args.model}", "frames": int(len(res.probs)),
                  "upload_chunk_samples": int(os.environ.get("W2VSEG_UPLOAD_CHUNK", 14 * 320000)),
                  "wall_s": [round(t, 4) for t in times], "audio_s_per_s": round(n / 16000 / best, 1),
                  "pdac_host_s": round(t_dac, 3)}))
