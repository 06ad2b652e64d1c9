"""ctypes binding of libw2vseg.so (the C ABI declared in include/w2vseg.h).

The product path has no CPU fallback: if the library is missing `load()` raises, and every compute
entry point fails loudly on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# W2VSEG_LIB points the loader at another BUILD of the same library (kernel A/B experiments under
# experiments/); it is never a fallback: the file must exist and export the whole ABI.
LIB_PATH = Path(os.environ.get("W2VSEG_LIB") or Path(__file__).resolve().parent / "csrc" / "libw2vseg.so")

_lib = None
ABI_VERSION = 4   # W2VSEG_ABI_VERSION of include/w2vseg.h this binding was written against


class W2VSegError(RuntimeError):
    pass


class Config(C.Structure):
    """mirror of w2vseg_config (include/w2vseg.h)"""

    _fields_ = [
        ("n_layers", C.c_int32),
        ("n_adapter_layers", C.c_int32),
        ("hidden", C.c_int32),
        ("heads", C.c_int32),
        ("ffn", C.c_int32),
        ("adapter_dim", C.c_int32),
        ("adapter_scale", C.c_float),
        ("conv_dim", C.c_int32),
        ("pos_kernel", C.c_int32),
        ("pos_groups", C.c_int32),
        ("head_layers", C.c_int32),
        ("head_heads", C.c_int32),
        ("head_ffn", C.c_int32),
        ("ln_eps", C.c_float),
        ("feat_group_norm", C.c_int32),
        ("conv_bias", C.c_int32),
        ("post_layer_norm", C.c_int32),
    ]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol include/w2vseg.h declares
SIGNATURES = {
    "w2vseg_abi_version": (_I32, []),
    "w2vseg_clock_probe": (_I32, [_P, _I32, _I32, _P]),
    "w2vseg_last_error": (C.c_char_p, []),
    "w2vseg_launch_count": (_I64, []),
    "w2vseg_device_ok": (_I32, []),
    "w2vseg_profile_enable": (_I32, [_I32]),
    "w2vseg_profile_collect": (_I64, [C.c_char_p, _SZ]),
    "w2vseg_num_frames": (_I32, [_I64]),
    "w2vseg_frame_stride": (_I32, [_I64]),
    "w2vseg_create": (_I32, [C.POINTER(Config), C.POINTER(_P)]),
    "w2vseg_destroy": (None, [_P]),
    "w2vseg_set_weight": (_I32, [_P, C.c_char_p, _P, _I64, _P]),
    "w2vseg_finalize_weights": (_I32, [_P, _P]),
    "w2vseg_workspace_bytes": (_SZ, [_P, _I32, _I64]),
    "w2vseg_encode": (_I32, [_P, _P, _I64, _P, _P, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "w2vseg_head": (_I32, [_P, _P, _I64, _I32, _P, _I32, _P, _P, _P, _SZ, _P]),
    "w2vseg_attention_train": (_I32, [_P, _I32, _I32, _I32, _I32, _P, C.c_float, _P, _P, C.c_float, C.c_uint32, _P]),
    "w2vseg_attention_bwd": (_I32, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, C.c_float, _P, C.c_float, C.c_uint32,
                             _P]),
    "w2vseg_head_grad_floats": (_I64, [_P]),
    "w2vseg_head_grad_offset": (_I64, [_P, C.c_char_p, C.POINTER(_I64)]),
    "w2vseg_head_train_workspace_bytes": (_SZ, [_P, _I32, _I32]),
    "w2vseg_head_train_step": (_I32, [_P, _P, _I64, _I32, _P, _P, C.c_float, _I32, _P, _P, _P, _SZ, C.c_float, C.c_float,
                               C.c_uint32, _P, _SZ, _P]),
    "w2vseg_calibrate": (_I32, [_P, _P, _I64, _P, _P, _P, _I32, _I64, _P, _SZ, _P]),
    "w2vseg_correct_bias": (_I32, [_P, C.c_char_p, _P, _I64, _P]),
    "w2vseg_sfc_forward": (_I32, [_P, _P, _I64, _P, _P, _P, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "w2vseg_sfc_forward_rows": (_I32, [_P, _P, _I64, _P, _P, _P, _I32, _I64, _P, _I64, _I32, _I32, _P, _SZ, _P]),
    "w2vseg_scatter_rows": (_I32, [_P, _I64, _P, _P, _I32, _P, _I64, _I32, _P]),
    "w2vseg_nanfill": (_I32, [_P, _I64, _P, _I32, _P]),
    "w2vseg_overlap_average": (_I32, [_P, _I32, _I64, _P, _P]),
    "w2vseg_moving_average": (_I32, [_P, _I64, _I32, _P, _P]),
    "w2vseg_gemm": (_I32, [_P, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _I32, _I32, _P]),
    "w2vseg_conv_gemm": (_I32, [_P, _I64, _I32, _I32, _I32, _P, _I32, _P, _P, _P]),
    "w2vseg_posconv": (_I32, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _I32, _P]),
    "w2vseg_conv0": (_I32, [_P, _I64, _P, _P, _P, _P, _P, _P, C.c_float, _P, _I32, _I32, _I32, _P, _SZ, _P]),
    "w2vseg_layernorm": (_I32, [_P, _I32, _I64, _I32, _P, _P, C.c_float, _I32, _P, _P]),
    "w2vseg_attention": (_I32, [_P, _I32, _I32, _I32, _I32, _P, C.c_float, _P, _P]),
    "w2vseg_attention_mma": (_I32, [_P, _I32, _I32, _I32, _I32, _P, C.c_float, _P, _P]),
}


def load():
    """dlopen libw2vseg.so and declare the prototypes. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise W2VSegError(
            f"{LIB_PATH} is missing: build it with `python -m wav2vecsegmenter_b200.build` "
            "(there is no CPU or PyTorch fallback for the SFC path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.w2vseg_abi_version() != ABI_VERSION:
        raise W2VSegError("libw2vseg.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().w2vseg_last_error().decode("utf-8", "replace")
        raise W2VSegError(f"{what or 'libw2vseg call'} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """device/host pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
