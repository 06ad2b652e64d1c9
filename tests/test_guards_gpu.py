"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU pool: profiles/sanitizer_r02.md).
Every buffer the C ABI writes — the caller-provided workspace, the logits / probability outputs, the row
matrix — is carved out of a larger allocation with 1 MiB guard bands filled with a byte pattern; after the
forward the bands must be untouched and a second, identically seeded run must reproduce the outputs bit for
bit (a race or a read of uninitialised workspace shows up as run-to-run differences)."""
import numpy as np
import pytest
import torch

from wav2vecsegmenter_b200 import _native as nat
from wav2vecsegmenter_b200 import synth
from wav2vecsegmenter_b200.engine import SFCEngine

pytestmark = pytest.mark.gpu
GUARD = 1 << 20
PAT = 0xA5


class Guarded:
    """`nbytes` usable bytes (1 KiB aligned) between two guard bands"""

    def __init__(self, nbytes, dev="cuda:0"):
        self.n = (int(nbytes) + 1023) // 1024 * 1024
        self.buf = torch.full((self.n + 2 * GUARD,), PAT, dtype=torch.uint8, device=dev)

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD

    def view(self, dtype, *shape):
        return self.buf[GUARD: GUARD + self.n].view(dtype)[: int(np.prod(shape))].view(*shape)

    def intact(self):
        return bool((self.buf[:GUARD] == PAT).all() and (self.buf[GUARD + self.n:] == PAT).all())


@pytest.mark.parametrize("lens", [[48000, 33234, 40000], [320000, 1234, 400], [7000]])
def test_forward_writes_stay_inside_their_buffers(lens):
    spec = synth.TINY
    eng = SFCEngine(spec)
    eng.load_state_dict(synth.random_state_dict(spec, 0))
    lib = eng.lib
    B, lmax = len(lens), max(lens)
    R = eng.frame_stride(lmax)
    audio = torch.zeros(B, lmax)
    for i, n in enumerate(lens):
        audio[i, :n] = synth.synthetic_audio(n, 200 + i)
    audio = audio.cuda()
    sl = torch.tensor(lens, dtype=torch.int32, device="cuda")
    nl = torch.full((B,), lmax, dtype=torch.int32, device="cuda")
    ol = torch.tensor([min(R, eng.num_frames(n) + 1) for n in lens], dtype=torch.int32, device="cuda")
    need = int(lib.w2vseg_workspace_bytes(eng._h, B, lmax))
    outs = []
    for rep in range(2):
        ws, lg, pr, rows = Guarded(need), Guarded(B * R * 4), Guarded(B * R * 4), Guarded(B * (R + 3) * 4)
        st = torch.cuda.current_stream().cuda_stream
        nat.check(lib.w2vseg_sfc_forward(eng._h, audio.data_ptr(), audio.stride(0), sl.data_ptr(), nl.data_ptr(),
                                         ol.data_ptr(), B, lmax, lg.ptr, pr.ptr, None, ws.ptr, ws.n, st), "sfc_forward")
        nat.check(lib.w2vseg_sfc_forward_rows(eng._h, audio.data_ptr(), audio.stride(0), sl.data_ptr(), nl.data_ptr(),
                                              ol.data_ptr(), B, lmax, rows.ptr, R + 3, R + 2, R + 1, ws.ptr, ws.n, st),
                  "sfc_forward_rows")
        torch.cuda.synchronize()
        assert ws.intact(), "workspace overrun"
        assert lg.intact() and pr.intact() and rows.intact(), "output overrun"
        p = pr.view(torch.float32, B, R).cpu()
        r = rows.view(torch.float32, B, R + 3).cpu()
        assert torch.isfinite(p).all()
        assert torch.equal(r[:, :R], p), "row-matrix entry point differs from sfc_forward"
        assert (r[:, R] == 0).all() and (r[:, R + 1] == 1).all()              # zero tail, `included` flag
        assert (r.view(torch.int32)[:, R + 2] == int.from_bytes(bytes([PAT] * 4), "little", signed=True)).all()  # beyond row_cols: untouched
        outs.append((p, lg.view(torch.float32, B, R).cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "run-to-run difference"
    eng.close()


def test_reductions_stay_inside_their_buffers():
    lib = nat.load()
    n, rows_n, stride = 5000, 6, 1001
    rows = torch.rand(rows_n, stride, device="cuda")
    rows[:, stride - 1] = 1.0
    start = torch.tensor([0, 999, 1998, 2997, 3996, 4900], dtype=torch.int32, device="cuda")
    count = torch.tensor([999, 999, 999, 999, 904, 100], dtype=torch.int32, device="cuda")
    talk, out = Guarded(n * 8), Guarded(n * 8)
    st = torch.cuda.current_stream().cuda_stream
    nat.check(lib.w2vseg_scatter_rows(rows.data_ptr(), stride, start.data_ptr(), count.data_ptr(), rows_n, talk.ptr, n,
                                      stride - 1, st), "scatter")
    nat.check(lib.w2vseg_moving_average(talk.ptr, n, 5, out.ptr, st), "moving_average")
    nat.check(lib.w2vseg_overlap_average(talk.ptr, 1, n, out.ptr, st), "overlap_average")
    torch.cuda.synchronize()
    assert talk.intact() and out.intact()


def test_two_handles_same_weights_bit_identical():
    """weight packing, folding and the bias-correction calibration are deterministic: two handles (as on
    two ranks of a multi-GPU job) loaded with the same checkpoint give bit-identical probabilities"""
    spec = synth.TINY
    sd = synth.random_state_dict(spec, 3)
    lens = [64000, 50000]
    audio = torch.zeros(2, 64000)
    for i, n in enumerate(lens):
        audio[i, :n] = synth.synthetic_audio(n, 300 + i)
    audio = audio.cuda()
    outs = []
    for _ in range(2):
        eng = SFCEngine(spec)
        eng.load_state_dict(sd)
        _, p = eng.sfc_forward(audio, lens, [64000, 64000], [eng.num_frames(n) for n in lens], 64000)
        outs.append(p.cpu())
        eng.close()
    assert torch.equal(outs[0], outs[1])


def test_clock_probe_reports_a_plausible_sm_clock():
    lib = nat.load()
    out = torch.zeros(8, device="cuda")
    nat.check(lib.w2vseg_clock_probe(out.data_ptr(), 8, 30, nat.current_stream_ptr()), "clock_probe")
    torch.cuda.synchronize()
    assert ((out > 300) & (out < 3000)).all(), out
