"""Builds libw2vseg.so (sm_100a only) in-tree with nvcc. No JIT cache, no torch extension machinery:
the shared library is a plain C-ABI artefact that travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = CSRC / "libw2vseg.so"
SOURCES = ["common.cu", "gemm_tc.cu", "gemm_tc2.cu", "posconv_tc.cu", "conv0_tc.cu", "kernels.cu", "attention.cu", "attention_bwd.cu", "attention_tc.cu", "attention_tc64.cu", "train_kernels.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return cand if Path(cand).exists() else "nvcc"


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "w2vseg.h"]
    objs = []
    jobs = []
    for src in SOURCES:
        s = CSRC / src
        o = CSRC / (Path(src).stem + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(LIB_PATH, objs):
        run([_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static"])
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
