"""attention time vs number of key tiles and vs batch size (is it bound by the K/V loads?)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wav2vecsegmenter_b200 import _native as n  # noqa: E402

lib = n.load()
R, heads, dh = 1000, 16, 64
D = heads * dh
for B in (14, 4, 28):
    qkv = torch.randn(B * R, 3 * D, device="cuda").bfloat16()
    ctx = torch.empty(B * R, D, device="cuda", dtype=torch.bfloat16)
    for klen in (128, 999):
        kv = torch.full((B,), klen, dtype=torch.int32, device="cuda")
        for _ in range(3):
            lib.w2vseg_attention(n.ptr(qkv), B, R, heads, dh, n.ptr(kv), dh ** -0.5, n.ptr(ctx), n.current_stream_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.w2vseg_attention(n.ptr(qkv), B, R, heads, dh, n.ptr(kv), dh ** -0.5, n.ptr(ctx), n.current_stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"B={B:3d} keys={klen:4d}: {us:7.1f} us  ({us / B:6.2f} us per window, qkv {B * R * 3 * D * 2 / 1e6:.0f} MB)")
