"""ctypes binding of libfocr_sm100.so (the C-ABI declared in include/focr.h).

The library is the product: there is no Python/PyTorch fallback for any op.  Importing this module
on a machine where the .so has not been built raises immediately.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libfocr_sm100.so"


class FocrError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise FocrError(
            f"{_LIB_PATH} is missing: run `python -m fudanocr_b200.build` (needs nvcc). "
            "There is no CPU / PyTorch fallback for the focr kernels.")
    return C.CDLL(str(_LIB_PATH))


lib = _load()
lib.focr_last_error.restype = C.c_char_p
lib.focr_version.restype = C.c_int

_vp, _i, _l, _sz, _fp = C.c_void_p, C.c_int, C.c_long, C.c_size_t, C.c_void_p

_SIGS = {
    "focr_sync_check": (C.c_int, [_vp]),
    "focr_conv2d_workspace_bytes": (_sz, [_i, _i, _i]),
    "focr_conv2d_fwd": (C.c_int, [_vp, _fp, _fp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_conv2d_dgrad": (C.c_int, [_vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_linear_workspace_bytes": (_sz, [_i, _i]),
    "focr_linear_fwd": (C.c_int, [_vp, _fp, _fp, _vp, _vp, _l, _i, _i, _i, _vp, _sz, _vp]),
    "focr_linear_dgrad": (C.c_int, [_vp, _fp, _vp, _l, _i, _i, _vp, _sz, _vp]),
}


def _bind():
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue
        fn.restype = res
        fn.argtypes = args


_bind()


def check(rc: int, what: str = "focr call") -> None:
    if rc != 0:
        raise FocrError(f"{what} failed ({rc}): {lib.focr_last_error().decode()}")


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
