"""torchrun target: the whole-talk path sharded over WORLD_SIZE GPUs (NCCL all_gather of
probability rows) must reproduce the single-GPU result bit for bit (BASELINE.json configs[2]).

    torchrun --nproc-per-node 2 scripts/check_multigpu.py [--talks 6 --seconds 900]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import synth  # noqa: E402
from wav2vecsegmenter_b200.engine import SFCEngine  # noqa: E402
from wav2vecsegmenter_b200.pipeline import TalkRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--talks", type=int, default=4)
ap.add_argument("--seconds", type=float, default=300.0)
ap.add_argument("--model", default="tiny", choices=["tiny", "large"])
args = ap.parse_args()

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0

spec = synth.TINY if args.model == "tiny" else synth.LARGE_ALL
eng = SFCEngine(spec, torch.device("cuda", local))
eng.load_state_dict(synth.random_state_dict(spec, 0))
rng = np.random.default_rng(0)
waves = [synth.synthetic_audio(int(args.seconds * 16000 + rng.integers(32000, 320000)), 900 + i).numpy()
         for i in range(args.talks)]

single = TalkRunner(eng, batch_size=14, inference_times=2).run(waves)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
sharded = TalkRunner(eng, batch_size=14, inference_times=2, dist_group=dist.group.WORLD if world > 1 else None).run(waves)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
same = all(np.array_equal(a.probs, b.probs) for a, b in zip(single, sharded))
# results_on=0: only rank 0 assembles the talks (what segment.py uses); the other ranks get None
r0 = TalkRunner(eng, batch_size=14, inference_times=2, dist_group=dist.group.WORLD if world > 1 else None).run(waves, results_on=0)
same_r0 = (r0 is None) if rank != 0 else all(np.array_equal(a.probs, b.probs) for a, b in zip(single, r0))
if rank == 0:
    print(json.dumps({"world": world, "talks": args.talks, "audio_seconds": float(sum(len(w) for w in waves) / 16000),
                      "frames": int(sum(len(r.probs) for r in single)), "bit_identical_to_single_gpu": bool(same), "rank0_only_results_identical": bool(same_r0),
                      "sharded_wall_s": round(dt, 3)}))
assert same and same_r0
if world > 1:
    dist.destroy_process_group()
