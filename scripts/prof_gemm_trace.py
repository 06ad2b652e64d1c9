"""experiment (needs the -DW2VSEG_TRACE build, W2VSEG_LIB=...): per-tile timeline of CTA 0 of the pair GEMM.
slots: 0 epilogue warp starts waiting for the accumulator, 1 accumulator ready, 2 epilogue of the tile done;
3 MMA warp starts waiting for a free accumulator, 4 got it, 5 all MMAs of the tile issued;
6 first TMEM block of the tile in registers, 7 first block computed and staged (before its global stores)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
raw = C.CDLL(str(n.LIB_PATH))
M, K = 14000, 1024
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
import os
cases = ((4608, 1024, 0, 0), (4608, 1024, 1, 0), (3072, 1024, 0, 0))
if os.environ.get("F32"):
    cases = ((1024, 1024, 0, 1), (1024, 4608, 0, 1))   # in-place residual GEMMs (attn_out, ffn_down)
for N, K, act, f32 in cases:
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / 32).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.randn(M, N, device="cuda") if f32 else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        n.check(lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, n.ptr(bias), act, n.ptr(out) if f32 else None, n.ptr(out), f32, 512, n.current_stream_ptr()))
    torch.cuda.synchronize()
    buf = np.zeros(768, dtype=np.int64)
    raw.w2vseg_debug_trace(buf.ctypes.data_as(C.c_void_p), 768)
    t = buf.reshape(12, 64)
    nt = int((t[2] > 0).sum())
    t0 = t[3, 0]
    print(f"N={N} K={K} act={act} f32-inplace={f32}: {nt} tiles on CTA 0; cycles relative to the MMA warp's start")
    print(" tile | mma: wait_acc got_acc issued | epi: wait ready done | epi busy  epi idle  mma-issue span | ready->ld0 math0 stage0 | ld1 math1 stage1+exit")
    for i in range(min(nt, 14)):
        r = [int(t[s, i] - t0) for s in (3, 4, 5, 0, 1, 2)]
        print(f"  {i:3d} | {r[0]:8d} {r[1]:8d} {r[2]:8d} | {r[3]:8d} {r[4]:8d} {r[5]:8d} | {r[5]-r[4]:8d} {r[4]-r[3]:8d} {r[2]-r[1]:8d} | {int(t[6,i]-t[1,i]):6d} {int(t[9,i]-t[6,i]):6d} {int(t[7,i]-t[9,i]):6d} | {int(t[8,i]-t[7,i]):6d} {int(t[10,i]-t[8,i]):6d} {int(t[2,i]-t[10,i]):6d}")
