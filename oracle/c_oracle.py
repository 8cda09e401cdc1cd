"""ctypes loader for the plain-C oracle (oracle/hamming_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hamming_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p, i32p = ctypes.c_void_p, ctypes.c_void_p
        L.orc_hamming_matrix.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, i32p]
        L.orc_hamming_matrix.restype = None
        L.orc_knn.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int64, i32p, i32p]
        L.orc_knn.restype = None
        L.orc_cross_check.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, u8p, ctypes.c_int64, i32p, i32p, i32p]
        L.orc_cross_check.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _prep(q, t, mask):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    assert q.ndim == 2 and t.ndim == 2 and q.shape[1] == 32 and t.shape[1] == 32
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert mask.shape == (q.shape[0], t.shape[0])
    return q, t, mask


def knn(q, t, k=1, mask=None):
    q, t, mask = _prep(q, t, mask)
    idx = np.full((q.shape[0], k), -1, dtype=np.int32)
    dist = np.full((q.shape[0], k), -1, dtype=np.int32)
    if q.shape[0] and t.shape[0]:
        lib().orc_knn(_p(q), q.shape[0], _p(t), t.shape[0], k, _p(mask), t.shape[0], _p(idx), _p(dist))
    return idx, dist


def cross_check(q, t, mask=None):
    q, t, mask = _prep(q, t, mask)
    n = q.shape[0]
    mq = np.empty(n, np.int32); mt = np.empty(n, np.int32); md = np.empty(n, np.int32)
    c = lib().orc_cross_check(_p(q), n, _p(t), t.shape[0], _p(mask), t.shape[0], _p(mq), _p(mt), _p(md)) if n and t.shape[0] else 0
    return mq[:c].copy(), mt[:c].copy(), md[:c].copy()
