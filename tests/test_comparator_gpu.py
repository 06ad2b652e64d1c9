"""Library-kernel comparator on the same GPU (SURVEY §8d "also useful"): the reference's computation — the oracle's
plain torch ops, i.e. what `model.to("cuda")` gives a user of the reference: cuBLAS / cuDNN / eager elementwise
kernels — on the headline workload (large 24/24 + adapters, 14 x 20 s), in fp32 (the reference's precision), with
TF32 matmuls, and in bf16, next to this library's fused forward. The oracle is only the thing being compared WITH
here; the numbers are printed (`-s`) and archived in profiles/comparator_r02.md."""
import pytest
import torch

from oracle import sfc_oracle
from wav2vecsegmenter_b200 import synth
from wav2vecsegmenter_b200.engine import SFCEngine

pytestmark = pytest.mark.gpu
B, L, T = 14, 320000, 999


def _time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_fused_forward_beats_library_eager_on_the_same_gpu():
    spec = synth.LARGE_ALL
    sd = synth.random_state_dict(spec, seed=0)
    audio = torch.stack([synth.synthetic_audio(L, 900 + i) for i in range(B)])
    audio = sfc_oracle.normalize_rows(audio, [True] * B).cuda()
    out_mask = torch.ones(B, T, dtype=torch.bool, device="cuda")
    lens = [L] * B
    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.no_grad():
            sd32 = {k: v.cuda() for k, v in sd.items()}
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            probs32 = sfc_oracle.batch_probs(sd32, audio, lens, out_mask, spec.keep_layers, spec.head_heads)[0]
            res["eager fp32"] = _time(lambda: sfc_oracle.batch_probs(sd32, audio, lens, out_mask, spec.keep_layers,
                                                                     spec.head_heads), 1, 3)
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
            res["eager tf32"] = _time(lambda: sfc_oracle.batch_probs(sd32, audio, lens, out_mask, spec.keep_layers,
                                                                     spec.head_heads), 1, 3)
            del sd32
            sd16 = {k: v.cuda().bfloat16() for k, v in sd.items()}
            a16 = audio.bfloat16()
            probs16 = sfc_oracle.batch_probs(sd16, a16, lens, out_mask, spec.keep_layers, spec.head_heads)[0].float()
            res["eager bf16"] = _time(lambda: sfc_oracle.batch_probs(sd16, a16, lens, out_mask, spec.keep_layers,
                                                                     spec.head_heads))
            del sd16
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()
    eng = SFCEngine(spec)
    eng.load_state_dict(sd)
    sl = torch.full((B,), L, dtype=torch.int32, device="cuda")
    ol = torch.full((B,), T, dtype=torch.int32, device="cuda")
    raw = torch.stack([synth.synthetic_audio(L, 900 + i) for i in range(B)]).cuda()
    _, probs = eng.sfc_forward(raw, sl, sl, ol, L)
    res["this library (bf16 operands, fp32 accumulate / residual)"] = _time(lambda: eng.sfc_forward(raw, sl, sl, ol, L), 3, 10)
    err_ours = (probs[:, :T] - probs32).abs().max().item()
    err_bf16 = (probs16 - probs32).abs().max().item()
    for k, ms in res.items():
        print(f"COMPARATOR {k}: {ms:.2f} ms per 14 x 20 s step = {B * 20 / ms * 1e3:.0f} audio-s/s")
    print(f"COMPARATOR max |p - p_fp32|: this library {err_ours:.4f}, eager bf16 {err_bf16:.4f}")
    eng.close()
    ours = res["this library (bf16 operands, fp32 accumulate / residual)"]
    assert ours < res["eager bf16"] and ours < res["eager tf32"] and ours < res["eager fp32"]
    assert err_ours < 2e-2
