"""Result tables: host-side mirror of the reference's writer side.

``save_detected_boundaries`` stands in for adapted/output.py:26-51 and ``BoundaryTableWriter`` for the two saver
threads of adapted/file_proc.py:312-457 (4000 reads per file, ``detected_boundaries_<i>.csv`` /
``failed_reads_<i>.csv``, ``fail_reason`` only in the fail files, batch indices continuing after a previous run,
file_proc.py:97-140).  The text comes from the native formatter ``adb_format_csv`` (adapted_b200/csrc/adb_csv.cpp),
which works on the fixed-layout records of the CUDA library -- no DetectResults objects, no pandas.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


def _id_array(read_ids: Sequence):
    """NUL-terminated ids in one fixed-width numpy buffer + the array of pointers into it (char **)."""
    ids = np.asarray(read_ids, dtype=object)
    enc = np.array([str(i).encode() for i in ids], dtype="S") if ids.size else np.zeros(0, "S1")
    enc = enc.astype(f"S{enc.dtype.itemsize + 1}")  # one more byte: numpy pads with NULs
    ptrs = enc.ctypes.data + np.arange(enc.size, dtype=np.uint64) * np.uint64(enc.dtype.itemsize)
    return np.ascontiguousarray(ptrs, dtype=np.uint64), enc


def format_detected_boundaries(recs: np.ndarray, read_ids: Sequence, primary_method: int,
                               save_fail_reasons: bool = False, sel: Optional[Iterable[int]] = None,
                               llr_detect_log: Optional[str] = None, open_pore_overflow: Optional[dict] = None) -> bytes:
    """CSV text of the reads ``sel`` (default: all) of ``recs``, as the reference's pandas path writes it.
    ``open_pore_overflow``: {record index: full open-pore list} for records beyond ADB_MAX_OPEN_PORES
    (``detect.open_pore_overflow``); such a record without its list raises OverflowError -- never truncated."""
    recs = np.ascontiguousarray(recs, dtype=_lib.RECORD_DTYPE)
    if len(read_ids) != recs.size:
        raise ValueError("one read id per record")
    L = _lib.load()
    ids, keep = _id_array(read_ids)
    if sel is None:
        sel_arr, n_sel, sel_ptr = None, recs.size, None
    else:
        sel_arr = np.ascontiguousarray(np.fromiter(sel, dtype=np.int32))
        if sel_arr.size and (sel_arr.min() < 0 or sel_arr.max() >= recs.size):
            raise IndexError("selection outside the record array")
        n_sel, sel_ptr = int(sel_arr.size), sel_arr.ctypes.data
    log = None if llr_detect_log is None else llr_detect_log.encode()
    from .detect import overflow_tables

    tabs = overflow_tables(open_pore_overflow, recs.size)
    op = (None, None, None) if tabs is None else tuple(t.ctypes.data for t in tabs)
    cap = 1024 + 400 * n_sel
    while True:
        buf = np.empty(cap, dtype=np.uint8)
        n = L.adb_format_csv_ex(recs.ctypes.data, sel_ptr, n_sel, ids.ctypes.data, int(primary_method), log,
                                int(bool(save_fail_reasons)), op[0], op[1], op[2], buf.ctypes.data, cap)
        if n == -6:
            raise OverflowError("a record's open-pore list is longer than ADB_MAX_OPEN_PORES and its overflow list was "
                                "not supplied (adapted_b200.detect.open_pore_overflow)")
        if n < 0:
            raise _lib.AdbError(int(n), "adb_format_csv: invalid argument")
        if n <= cap:
            del keep
            return buf[:n].tobytes()
        cap = int(n)


def save_detected_boundaries(recs: np.ndarray, read_ids: Sequence, filename: str, primary_method: int,
                             save_fail_reasons: bool = False, sel: Optional[Iterable[int]] = None,
                             llr_detect_log: Optional[str] = None, open_pore_overflow: Optional[dict] = None) -> None:
    """adapted/output.py:26-51 for record arrays."""
    with open(filename, "wb") as f:
        f.write(format_detected_boundaries(recs, read_ids, primary_method, save_fail_reasons, sel, llr_detect_log,
                                           open_pore_overflow))


class BoundaryTableWriter:
    """Collects minibatch results and writes the pass / fail tables like the reference's saver threads
    (file_proc.py:312-351: full files of ``batch_size_output`` reads in arrival order, the remainder on close)."""

    def __init__(self, output_dir_boundaries: str, output_dir_fail: str, primary_method: int,
                 batch_size_output: int = 4000, bidx_pass: int = 0, bidx_fail: int = 0,
                 llr_detect_log: Optional[str] = None):
        self.dirs = {"pass": output_dir_boundaries, "fail": output_dir_fail}
        self.names = {"pass": "detected_boundaries", "fail": "failed_reads"}
        self.bidx = {"pass": int(bidx_pass), "fail": int(bidx_fail)}
        self.method = int(primary_method)
        self.batch = int(batch_size_output)
        self.log = llr_detect_log
        self._recs = {"pass": [], "fail": []}
        self._ids = {"pass": [], "fail": []}
        self._count = {"pass": 0, "fail": 0}
        self._over = {}  # read id -> full open-pore list of a record beyond ADB_MAX_OPEN_PORES
        self.files: List[str] = []
        for d in self.dirs.values():
            os.makedirs(d, exist_ok=True)

    @classmethod
    def continue_from(cls, path: str, primary_method: int, **kw) -> "BoundaryTableWriter":
        """Batch indices continue after the files of a previous run (file_proc.py:97-140)."""
        def next_idx(sub, prefix):
            d = os.path.join(path, sub)
            idx = [int(f.split("_")[-1].split(".")[0]) for f in os.listdir(d)
                   if f.startswith(prefix) and f.endswith(".csv")] if os.path.isdir(d) else []
            return max(idx, default=-1) + 1

        return cls(os.path.join(path, "boundaries"), os.path.join(path, "failed_reads"), primary_method,
                   bidx_pass=next_idx("boundaries", "detected_boundaries_"),
                   bidx_fail=next_idx("failed_reads", "failed_reads_"), **kw)

    def add(self, recs: np.ndarray, read_ids: Sequence, open_pore_overflow: Optional[dict] = None) -> None:
        """One minibatch: split by ``success`` (file_proc.py:246-266) and flush complete files.
        ``open_pore_overflow``: {index into recs: full open-pore list} (``detect.open_pore_overflow``)."""
        recs = np.asarray(recs, dtype=_lib.RECORD_DTYPE)
        ok = recs["success"] != 0
        ids = np.asarray(read_ids, dtype=object)
        for i, lst in (open_pore_overflow or {}).items():
            self._over[ids[i]] = np.asarray(lst, dtype=np.int32)
        for key, mask in (("fail", ~ok), ("pass", ok)):
            if mask.any():
                self._recs[key].append(recs[mask])
                self._ids[key].extend(ids[mask].tolist())
                self._count[key] += int(mask.sum())
            while self._count[key] >= self.batch:
                self._flush(key, self.batch)

    def _flush(self, key: str, n: int) -> None:
        allr = np.concatenate(self._recs[key]) if len(self._recs[key]) != 1 else self._recs[key][0]
        fn = os.path.join(self.dirs[key], f"{self.names[key]}_{self.bidx[key]}.csv")
        over = {i: self._over.pop(rid) for i, rid in enumerate(self._ids[key][:n]) if rid in self._over} if self._over else None
        save_detected_boundaries(allr[:n], self._ids[key][:n], fn, self.method, save_fail_reasons=(key == "fail"),
                                 llr_detect_log=self.log, open_pore_overflow=over)
        self.files.append(fn)
        self._recs[key] = [allr[n:]] if allr.size > n else []
        self._ids[key] = self._ids[key][n:]
        self._count[key] -= n
        self.bidx[key] += 1

    def close(self) -> None:
        for key in ("pass", "fail"):
            if self._count[key] > 0:
                self._flush(key, self._count[key])

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
