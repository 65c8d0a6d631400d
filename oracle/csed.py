"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/sed_exact.c (checker + CPU baseline)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_sed.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sed_exact.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/liboracle_sed.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_sed_exact.restype = ctypes.c_double
        _lib.oracle_sed_exact.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def sed_exact_many(e, xa, ya, xb, yb):
    """Bit-faithful sed.py values for one E and arrays of K-normalised correspondences."""
    e = np.ascontiguousarray(e, dtype=np.float64).reshape(9)
    xa, ya, xb, yb = (np.ascontiguousarray(v, dtype=np.float64) for v in (xa, ya, xb, yb))
    out = np.empty(len(xa), dtype=np.float64)
    lib().oracle_sed_exact_many(_p(e), _p(xa), _p(ya), _p(xb), _p(yb), ctypes.c_int64(len(xa)), _p(out))
    return out


def score_batch(E, xa, ya, xb, yb, thr, table=None, valid=None, nthreads=1):
    """(count_extra, S1, S2) per hypothesis — ransac.py:66-82 restated (see sed_exact.c)."""
    E = np.ascontiguousarray(E, dtype=np.float64).reshape(-1, 9)
    h = E.shape[0]
    xa, ya, xb, yb = (np.ascontiguousarray(v, dtype=np.float64) for v in (xa, ya, xb, yb))
    cnt = np.empty(h, dtype=np.int32)
    s1 = np.empty(h, dtype=np.float64)
    s2 = np.empty(h, dtype=np.float64)
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.int32).reshape(h, 8)
    if valid is not None:
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
    lib().oracle_score_batch(
        _p(E), _p(valid) if valid is not None else None, _p(xa), _p(ya), _p(xb), _p(yb),
        ctypes.c_int64(len(xa)), ctypes.c_int64(h), _p(table) if table is not None else None,
        ctypes.c_double(thr), _p(cnt), _p(s1), _p(s2), ctypes.c_int(nthreads))
    return cnt, s1, s2
