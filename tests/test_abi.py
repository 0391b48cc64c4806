"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/adapted_b200.h declares,
struct layouts agree, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from adapted_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "adapted_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/adapted_b200.h but not exported"


def test_struct_layouts():
    L = _lib.load()
    assert L.adb_abi_version() == 1
    assert L.adb_record_size() == _lib.RECORD_DTYPE.itemsize == 512
    assert L.adb_config_size() == ctypes.sizeof(_lib.AdbConfig)


def test_no_cpu_fallback():
    L = _lib.load()
    if L.adb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.AdbError, match="no CPU fallback"):
        _lib.Context(0)
    from adapted_b200.config import get_chemistry_specific_config
    from adapted_b200.detect import combined_detect_llr2

    spc = get_chemistry_specific_config("rna002")
    with pytest.raises(_lib.AdbError):
        combined_detect_llr2(np.zeros((2, spc.sig_preload_size), np.float32), np.array([5, 5], np.int32), spc)


def test_product_does_not_import_oracle():
    """the product path must never route through the oracle"""
    pkg = os.path.join(ROOT, "adapted_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
