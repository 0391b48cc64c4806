"""Build libadapted_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libadapted_b200.so")
SOURCES = ["adb_api.cu", "adb_csv.cpp"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # the arithmetic contract: no silent FMA contraction; explicit fmaf() only
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".cpp"))]
    deps.append(os.path.join(HERE, "..", "..", "include", "adapted_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *FLAGS, *[os.path.join(HERE, s) for s in SOURCES], "-o", SO]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libadapted_b200.so")
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
