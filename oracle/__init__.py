"""oracle/ -- CPU restatement of the reference's self-play inference hot path.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py, never by the product package `kzero_b200`.

Pieces:
  kz_oracle.c   plain-C loops (plane expansion, conv, gemm, decode) -> oracle/_build/libkz_oracle.so
  onnx_min.py   minimal ONNX wire-format reader
  graph_exec.py op-by-op f32 interpreter of the ONNX graph (restates what the nn-graph CPU executor,
                kn-graph 0.7.3 `cpu_eval_graph_exec`, call site rust/kz-core/src/network/cpu.rs:50, does)

Parity pinning: see the header of kz_oracle.c and DESIGN.md section "Oracle".
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_build" / "libkz_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile kz_oracle.c with gcc (OpenMP on).  Building the checker is not using it."""
    src = _DIR / "kz_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        _LIB_PATH.parent.mkdir(exist_ok=True)
        cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden", "-o", str(_LIB_PATH), str(src),
               "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        i64 = ctypes.c_int64
        vp = ctypes.c_void_p
        L.kzo_max_threads.restype = ctypes.c_int
        L.kzo_set_threads.argtypes = [ctypes.c_int]
        L.kzo_expand_planes.argtypes = [vp, vp, i64, i64, i64, i64, vp]
        L.kzo_conv2d.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ctypes.c_int]
        L.kzo_gemm.argtypes = [vp, vp, vp, vp, i64, i64, i64, ctypes.c_int, ctypes.c_float, ctypes.c_float]
        L.kzo_decode_output.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp]
        L.kzo_decode_output.restype = i64
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_threads(n: int) -> None:
    lib().kzo_set_threads(int(n))


def max_threads() -> int:
    return int(lib().kzo_max_threads())


def expand_planes(bits: np.ndarray, scalars: np.ndarray, bool_shape, scalar_count: int) -> np.ndarray:
    """(bits [n, ceil(Cb*A/8)] u8, scalars [n, Cs] f32) -> planes [n, Cs+Cb, H, W] f32.

    Restates InputMapper::encode_input_full, rust/kz-core/src/mapping/mod.rs:40-63."""
    cb, h, w = bool_shape
    n = bits.shape[0]
    area = h * w
    bits = np.ascontiguousarray(bits, dtype=np.uint8).reshape(n, -1)
    assert bits.shape[1] == (cb * area + 7) // 8, (bits.shape, cb, area)
    scalars = np.ascontiguousarray(scalars, dtype=np.float32).reshape(n, scalar_count)
    out = np.empty((n, scalar_count + cb, h, w), dtype=np.float32)
    lib().kzo_expand_planes(_p(bits), _p(scalars), n, cb * area, scalar_count, area, _p(out))
    return out


def conv2d(x: np.ndarray, w: np.ndarray, b, pad: int, relu: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    n, ci, h, wd = x.shape
    co, ci2, k, k2 = w.shape
    assert ci == ci2 and k == k2
    ho, wo = h + 2 * pad - k + 1, wd + 2 * pad - k + 1
    y = np.empty((n, co, ho, wo), dtype=np.float32)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
    lib().kzo_conv2d(_p(x), _p(w), None if bb is None else _p(bb), _p(y), n, ci, h, wd, co, k, pad, int(relu))
    return y


def gemm(x: np.ndarray, w: np.ndarray, c, trans_b: bool, alpha: float = 1.0, beta: float = 1.0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    m, k = x.shape
    n = w.shape[0] if trans_b else w.shape[1]
    assert (w.shape[1] if trans_b else w.shape[0]) == k
    y = np.empty((m, n), dtype=np.float32)
    cc = None if c is None else np.ascontiguousarray(np.broadcast_to(c, (n,)), dtype=np.float32)
    lib().kzo_gemm(_p(x), _p(w), None if cc is None else _p(cc), _p(y), m, k, n, int(trans_b), alpha, beta)
    return y


def decode_output(scalars: np.ndarray, policy_logits: np.ndarray, mv_idx: np.ndarray, mv_off: np.ndarray):
    """Restates decode_output, rust/kz-core/src/network/common.rs:16-100.

    -> (values [B,5] = tanh(value), wdl softmax, moves_left ; policy probs CSR-aligned with mv_idx)."""
    scalars = np.ascontiguousarray(scalars, dtype=np.float32)
    b = scalars.shape[0]
    policy_logits = np.ascontiguousarray(policy_logits, dtype=np.float32).reshape(b, -1)
    mv_idx = np.ascontiguousarray(mv_idx, dtype=np.uint32)
    mv_off = np.ascontiguousarray(mv_off, dtype=np.uint32)
    assert mv_off.shape[0] == b + 1
    assert mv_idx.size == 0 or int(mv_idx.max()) < policy_logits.shape[1], "policy index out of range"
    out_v = np.empty((b, 5), dtype=np.float32)
    out_p = np.empty((int(mv_off[-1]),), dtype=np.float32)
    rc = lib().kzo_decode_output(_p(scalars), _p(policy_logits), b, policy_logits.shape[1], _p(mv_idx), _p(mv_off),
                                 _p(out_v), _p(out_p))
    if rc != 0:
        raise FloatingPointError(f"softmax sum not strictly positive for board {-rc - 1} (reference would panic)")
    return out_v, out_p
