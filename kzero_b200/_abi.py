"""ctypes binding of libkzb200.so (include/kzb200.h).  Fails loudly if the library is missing."""
from __future__ import annotations

import ctypes
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libkzb200.so"
_lib = None


class NetInfo(ctypes.Structure):
    _fields_ = [("input_channels", ctypes.c_int32), ("board_h", ctypes.c_int32), ("board_w", ctypes.c_int32),
                ("policy_len", ctypes.c_int32), ("channels", ctypes.c_int32), ("depth", ctypes.c_int32),
                ("max_batch", ctypes.c_int32), ("precision", ctypes.c_int32), ("device", ctypes.c_int32),
                ("conv_mode", ctypes.c_int32), ("flops_per_position", ctypes.c_double)]


# every symbol include/kzb200.h declares: name -> (restype, argtypes)
_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
SYMBOLS = {
    "kzb_device_count": (_i, []),
    "kzb_last_error": (ctypes.c_char_p, []),
    "kzb_net_create_from_onnx": (_i, [_i, _vp, _sz, _i, _i, ctypes.POINTER(_vp)]),
    "kzb_net_bind_mapper": (_i, [_vp, _i, _i, _i, _i, _i]),
    "kzb_net_get_info": (_i, [_vp, ctypes.POINTER(NetInfo)]),
    "kzb_onnx_inspect": (_i, [_vp, _sz, ctypes.POINTER(NetInfo)]),
    "kzb_net_destroy": (None, [_vp]),
    "kzb_eval_planes": (_i, [_vp, _vp, _i, _vp, _vp]),
    "kzb_eval_packed": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "kzb_encode_planes": (_i, [_vp, _vp, _vp, _i, _vp]),
    "kzb_stage_packed": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "kzb_time_staged": (_i, [_vp, _i, _i, _vp]),
    "kzb_profile_staged": (_i, [_vp, _i, _vp, _sz, _vp, _i, ctypes.POINTER(_i)]),
    "kzb_launches_per_eval": (_i, [_vp]),
}


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} is missing: build it with `python -m kzero_b200.build` "
                               "(there is no CPU fallback)")
        L = ctypes.CDLL(str(_LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class KzbError(RuntimeError):
    """A non-zero return from the C ABI; the reference panics in the same situations."""


def check(rc: int) -> None:
    if rc != 0:
        raise KzbError(lib().kzb_last_error().decode(errors="replace"))
