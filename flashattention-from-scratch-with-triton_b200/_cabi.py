"""ctypes binding of libfa_sm100.so (include/fa_sm100.h).

There is no fallback: if the library is missing or does not export a declared symbol the import
of the operator fails loudly — the product never routes through the CPU oracle or another backend.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FA_SM100_LIB selects another build of the SAME library (A/B tuning variants); there is still no other backend
LIB_PATH = os.environ.get("FA_SM100_LIB") or os.path.join(_HERE, "libfa_sm100.so")

# every symbol include/fa_sm100.h declares: name -> (restype, argtypes)
_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
SYMBOLS = {
    "fa_sm100_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "fa_sm100_bwd": (_i, [_vp] * 10 + [_i] * 7 + [_f, _vp]),
    "fa_sm100_fwd_strided": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "fa_sm100_bwd_strided": (_i, [_vp] * 10 + [_i] * 8 + [_f, _vp, _vp, _i]),
    "fa_sm100_fwd_ranges": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "fa_sm100_bwd_ranges": (_i, [_vp] * 10 + [_i] * 8 + [_f, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "fa_sm100_fwd_opt": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "fa_sm100_bwd_opt": (_i, [_vp] * 10 + [_i] * 8 + [_f, _vp, _vp, _vp, _i]),
    "fa_sm100_bwd_fused": (_i, [_vp] * 11 + [_i] * 8 + [_f, _vp, _vp, _i]),
    "fa_sm100_bwd_fused_opt": (_i, [_vp] * 11 + [_i] * 8 + [_f, _vp, _vp, _vp, _i]),
    "fa_sm100_bwd_fused_workspace": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "fa_sm100_bwd_parts": (_i, [_vp] * 10 + [_i] * 7 + [_f, _vp, _i]),
    "fa_sm100_delta": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fa_sm100_merge": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fa_sm100_supported": (_i, [_i, _i, _i, _i]),
    "fa_last_error": (ctypes.c_char_p, []),
    "fa_sm100_version": (_i, []),
    "fa_sm100_launch_count": (ctypes.c_ulonglong, []),
    "fa_sm100_set_shared_sms": (ctypes.c_int, [ctypes.c_int]),
    "fa_sm100_last_hang": (_i, [ctypes.POINTER(ctypes.c_uint * 4)]),
}

FA_DTYPE_FP16, FA_DTYPE_BF16 = 0, 1


class Options(ctypes.Structure):
    """struct fa_sm100_options (include/fa_sm100.h): range masks and dropout."""
    _fields_ = [("row_lo", _vp), ("row_hi", _vp), ("col_lo", _vp), ("col_hi", _vp),
                ("dropout_p", _f), ("dropout_seed", ctypes.c_ulonglong)]

_lib = None


class FaSm100Error(RuntimeError):
    """A C-ABI call returned non-zero (negative: FA_ERR_*; positive: cudaError_t)."""

    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load the library once; raise if absent.  `build_if_missing` runs nvcc (build.py) first."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        _build.build_lib()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with "
            "`python flashattention-from-scratch-with-triton_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU or Triton fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(fn_name: str, rc: int) -> None:
    if rc != 0:
        raise FaSm100Error(fn_name, rc, load().fa_last_error().decode(errors="replace"))


def last_hang():
    out = (ctypes.c_uint * 4)()
    if load().fa_sm100_last_hang(ctypes.byref(out)):
        return dict(tag=out[0], block=out[1], thread=out[2], parity=out[3])
    return None
