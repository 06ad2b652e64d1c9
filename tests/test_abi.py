"""CPU: the C-ABI library loads and exports every symbol include/w2vseg.h declares; argument
errors are reported through return codes + w2vseg_last_error (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

from wav2vecsegmenter_b200 import _native

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "w2vseg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2vseg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _native.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in w2vseg.h but not exported"
        assert s in _native.SIGNATURES, f"{s} has no ctypes prototype"
    assert set(_native.SIGNATURES) == set(syms)
    assert lib.w2vseg_abi_version() == _native.ABI_VERSION == 4


def test_geometry_is_pure_host_arithmetic():
    lib = _native.load()
    assert lib.w2vseg_num_frames(320000) == 999
    assert lib.w2vseg_num_frames(399) == 0 and lib.w2vseg_num_frames(400) == 1
    assert lib.w2vseg_frame_stride(320000) == 1000
    assert lib.w2vseg_frame_stride(352000) == 1100


def test_errors_are_codes_not_exceptions():
    lib = _native.load()
    rc = lib.w2vseg_moving_average(None, 10, 5, None, None)
    assert rc == -1 and b"moving_average" in lib.w2vseg_last_error()
    h = ctypes.c_void_p()
    cfg = _native.Config(n_layers=2, n_adapter_layers=0, hidden=768, heads=12, ffn=3072, adapter_dim=512,
                         adapter_scale=4.0, conv_dim=512, pos_kernel=128, pos_groups=16, head_layers=1,
                         head_heads=8, head_ffn=2048, ln_eps=1e-5)
    assert lib.w2vseg_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"hidden" in lib.w2vseg_last_error()


def test_product_path_does_not_import_oracle():
    """the oracle is test infrastructure: nothing shipped may import it"""
    bad = []
    files = list((ROOT / "wav2vecsegmenter_b200").rglob("*.py")) + list((ROOT / "lib").rglob("*.py")) + \
        [ROOT / "segment.py", ROOT / "inference.py"]
    for f in files:
        if re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), flags=re.M):
            bad.append(str(f))
    assert not bad, bad
