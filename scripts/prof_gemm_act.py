"""experiment: what the activation in the epilogue costs the FFN-up GEMM (M=14000, K=1024, pair kernel)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
M, K = 14000, 1024
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
for N in (4608, 3072):
    W = (torch.randn(N, K, device="cuda", generator=g) / 32).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    import os
    acts = ((0, "none"), (2, "relu"), (1, "gelu"), (0, "none"))
    if os.environ.get("W2VSEG_LIB"):   # experimental build (-DW2VSEG_EPI_EXPERIMENT): sensitivity variants
        acts += ((3, "MUFU only"), (4, "8 FP ops, no MUFU"), (5, "3 FP ops + MUFU"), (1, "gelu"))
    for act, nm in acts:
        for _ in range(3):
            n.check(lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, n.ptr(bias), act, None, n.ptr(out), 0, 512,
                                    n.current_stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.w2vseg_gemm(n.ptr(A), n.ptr(W), M, N, K, n.ptr(bias), act, None, n.ptr(out), 0, 512,
                            n.current_stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"N={N} act={nm}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
