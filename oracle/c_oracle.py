"""ctypes loader for oracle/libcova_oracle.so (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcova_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cova_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcova_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u8p, i32p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_int32, ctypes.c_int64))
        L.oracle_ccl.restype = ctypes.c_int
        L.oracle_ccl.argtypes = [u8p, ctypes.c_int, ctypes.c_int, i32p, i32p, i32p]
        L.oracle_bboxcc.restype = ctypes.c_long
        L.oracle_bboxcc.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int32, u8p, ctypes.c_size_t, i32p]
        L.oracle_bboxcc_batch.restype = ctypes.c_long
        L.oracle_bboxcc_batch.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int32, u8p,
                                          ctypes.c_size_t, i64p, i32p]
        L.oracle_metapreprocess_stream.restype = ctypes.c_long
        L.oracle_metapreprocess_stream.argtypes = [u8p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                                   ctypes.c_uint32, ctypes.c_uint32, u8p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _scratch(h, w):
    nb = ((h + 1) // 2) * ((w + 1) // 2)
    return np.empty(3 * h * w + 16 + h * w + 5 * (nb + 1), dtype=np.int32)


def ccl(mask: np.ndarray):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    nb = ((h + 1) // 2) * ((w + 1) // 2)
    labels = np.empty((h, w), np.int32)
    stats = np.empty((nb + 1, 5), np.int32)
    scratch = _scratch(h, w)
    n = lib().oracle_ccl(_p(mask, ctypes.c_uint8), h, w, _p(labels, ctypes.c_int32), _p(stats, ctypes.c_int32),
                         _p(scratch, ctypes.c_int32))
    return n, labels, stats[:n].copy()


def bboxcc(mask: np.ndarray, area_thresh: int) -> bytes:
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    cap = 8 + 24 * ((h + 1) // 2) * ((w + 1) // 2)
    out = np.empty(cap, np.uint8)
    scratch = _scratch(h, w)
    r = lib().oracle_bboxcc(_p(mask, ctypes.c_uint8), h, w, int(area_thresh), _p(out, ctypes.c_uint8), cap,
                            _p(scratch, ctypes.c_int32))
    assert r > 0
    return out[:r].tobytes()


def bboxcc_batch(masks: np.ndarray, area_thresh: int):
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    n, h, w = masks.shape
    cap = 8 + 24 * ((h + 1) // 2) * ((w + 1) // 2)
    out = np.empty((n, cap), np.uint8)
    lens = np.empty(n, np.int64)
    scratch = _scratch(h, w)
    lib().oracle_bboxcc_batch(_p(masks, ctypes.c_uint8), n, h, w, int(area_thresh), _p(out, ctypes.c_uint8), cap,
                              _p(lens, ctypes.c_int64), _p(scratch, ctypes.c_int32))
    return [out[i, :lens[i]].tobytes() for i in range(n)]


def metapreprocess_stream(frames: np.ndarray, timestep: int, gamma: int = 1) -> np.ndarray:
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    F, H, W, C = frames.shape
    out = np.empty((max(F, 1), timestep * H, W, 4), np.uint8)
    n = lib().oracle_metapreprocess_stream(_p(frames, ctypes.c_uint8), F, W * 16, H * 16, timestep, gamma,
                                           _p(out, ctypes.c_uint8))
    return out[:n]
