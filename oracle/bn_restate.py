"""bottleneck.move_mean / move_var stand-in (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED).

bottleneck is a third-party dependency of the reference (adapted/detect/mvs.py:15,93-96,103-106) that is
absent from this image.  The arithmetic lives in oracle/bn_restate.c; this wrapper only dispatches on
dtype like the real library (float32 in -> float32 out); the hot path never feeds anything else.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._clib import lib

__version__ = "1.3.7"


def _run(fn_name: str, a, window: int):
    a = np.ascontiguousarray(a)
    if a.ndim != 1:
        raise ValueError("oracle bottleneck stand-in: 1-D input only")
    if window < 1 or window > a.size:
        raise ValueError("Moving window (=%d) must between 1 and %d, inclusive" % (window, a.size))
    if a.dtype != np.float32:
        raise TypeError("oracle bottleneck stand-in only restates the float32 kernels")
    y = np.empty_like(a)
    fp = ctypes.POINTER(ctypes.c_float)
    getattr(lib(), fn_name)(a.ctypes.data_as(fp), a.size, int(window), y.ctypes.data_as(fp))
    return y


def move_mean(a, window, min_count=None, axis=-1):
    assert min_count is None
    return _run("adb_oracle_move_mean_f32", a, window)


def move_var(a, window, min_count=None, axis=-1, ddof=0):
    assert min_count is None and ddof == 0
    return _run("adb_oracle_move_var_f32", a, window)
