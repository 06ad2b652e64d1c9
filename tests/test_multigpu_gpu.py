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
