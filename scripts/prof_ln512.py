"""LayerNorm(512) + GELU of the conv stack at the layer-1 size (14 x 31 999 rows): GB/s of one launch"""
import torch
from wav2vecsegmenter_b200 import _native as n

lib = n.load()
for rows in (14 * 31999, 14 * 15999, 14 * 3999):
    x = (torch.randn(rows, 512, device="cuda") * 2).bfloat16()
    gamma = torch.randn(512, device="cuda"); beta = torch.randn(512, device="cuda")
    out = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n.check(lib.w2vseg_layernorm(n.ptr(x), 0, rows, 512, n.ptr(gamma), n.ptr(beta), 1e-5, 1, n.ptr(out), n.current_stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"rows {rows}: {t * 1e3:.1f} us, {rows * 512 * 4 / t / 1e6:.0f} GB/s")
