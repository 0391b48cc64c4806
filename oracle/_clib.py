"""ctypes loader for the C part of the oracle (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("llr_gains.c", "bn_restate.c")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", *srcs, "-o", _SO, "-lm"]
        )
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, dp, fp = ctypes.c_int64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
        L.adb_oracle_pairwise_sum_f64.argtypes = [dp, i64]
        L.adb_oracle_pairwise_sum_f64.restype = ctypes.c_double
        L.adb_oracle_cumsum.argtypes = [dp, i64, dp, dp]
        L.adb_oracle_cumsum.restype = None
        L.adb_oracle_llr_gains.argtypes = [dp, dp, i64, i64, i64, i64, i64, i64, ctypes.c_int, i64, i64, i64, i64, dp]
        L.adb_oracle_llr_gains.restype = ctypes.c_int
        L.adb_oracle_move_mean_f32.argtypes = [fp, i64, i64, fp]
        L.adb_oracle_move_mean_f32.restype = None
        L.adb_oracle_move_var_f32.argtypes = [fp, i64, i64, fp]
        L.adb_oracle_move_var_f32.restype = None
        _lib = L
    return _lib
