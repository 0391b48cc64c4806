"""Result containers: host-side mirror of the reference's output schema.

Field names and their order are the CSV contract (adapted/container_types.py:22-120);
they are generated from one table here.  :func:`records_to_results` turns the fixed-layout ``adb_record``
array written by the CUDA library into ``DetectResults`` objects with the reference's None / value
conventions and the exact ``fail_reason`` strings of adapted/detect/combined.py:396-580.
"""
from __future__ import annotations

from dataclasses import make_dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

_PART = ("start", "len", "mean", "std", "med", "mad")
FIELD_ORDER: List[str] = (
    ["success", "signal_len", "preloaded"]
    + [f"adapter_{k}" for k in ("start", "end", "len", "mean", "std", "med", "mad")]
    + [f"polya_{k}" for k in ("start", "end", "len", "mean", "std", "med", "mad", "truncated", "candidates")]
    + [f"rna_preloaded_{k}" for k in _PART]
    + [f"start_peak_{k}" for k in ("idx", "pa", "next_max_idx", "next_max_pa", "open_pore_idx", "open_pore_type")]
    + ["adapter_rna_median_shift", "llr_adapter_end", "llr_polya_end", "cnn_adapter_end", "cnn_polya_end",
       "start_peak_adapter_end", "start_peak_polya_end", "llr_trace", "llr_adapter_end_adjust",
       "llr_polya_end_adjust", "llr_trace_early_stop_pos", "mvs_llr_polya_end_adjust_ignored",
       "mvs_llr_polya_end_to_early_stop", "mvs_adapter_end"]
    + [f"mvs_detect_{k}" for k in ("mean_at_loc", "var_at_loc", "polya_med", "polya_local_range", "med_shift")]
    + ["real_adapter_mean_start", "real_adapter_mean_end", "real_adapter_local_range", "open_pores",
       "fail_reason", "llr_detect_log"]
)


def _to_dict(self) -> Dict[str, Any]:
    return dict(self.__dict__)


def _update(self, d: dict) -> None:
    self.__dict__.update(d)


DetectResults = make_dataclass(
    "DetectResults",
    [("success", bool)] + [(n, Optional[Any], field(default=None)) for n in FIELD_ORDER[1:]],
    namespace={"to_dict": _to_dict, "update": _update},
)
DetectResults.__doc__ = "Per-read detection result (same fields as the reference's DetectResults)."


def _summary(self) -> Dict[str, Any]:
    d = self.detect_results.to_dict() if self.detect_results else {}
    d.pop("fail_reason", None)
    return {"read_id": self.read_id, **d, "fail_reason": self.fail_reason}


ReadResult = make_dataclass(
    "ReadResult",
    [("read_id", Optional[str], field(default=None)), ("success", bool, field(default=True)),
     ("fail_reason", Optional[str], field(default=None)),
     ("detect_results", Optional[Any], field(default=None))],
    namespace={"to_summary_dict": _summary},
)

# valid bits (include/adapted_b200.h)
(V_ADAPTER, V_POLYA, V_RNA, V_MVS, V_REAL_MEANS, V_REAL_RANGE, V_OPEN, V_MEDSHIFT, V_CAND, V_SP, V_SP_OPEN, V_FIELDS,
 V_MVS_ADAPTER_END, V_TO_EARLY_STOP, V_POLYA_NONE) = (1 << i for i in range(15))

_FAIL_TEXT = {
    1: "No adapter detected (primary)",
    2: "adapter MAD check failed",
    3: "Open pore too close to boundary",
    4: "Real signal check failed",
    5: "No polya detected (primary)",
    6: "MVS polya check failed: not enough signal",
    8: "No adapter detected in range (mvs_detect)",
    9: "Median shift check failed",
    20: "pA_mean_range is not specified",
    21: "'NoneType' object is not iterable",
    22: "attempt to get argmin of an empty sequence",
    23: "slice indices must be integers or None or have an __index__ method",
    24: "MAD normalization failed: scale is 0",
}
_MVS_NAMES = ("mean ", "var ", "med ", "range ", "shift")
_SP_FLAGS = {1: "open pore in adapter", 2: "potential concatemer adapter-only read"}
_METHOD = {0: "llr", 1: "cnn", 2: "start_peak"}


def fail_reason_text(code: int, mask: int) -> Optional[str]:
    if code == 0:
        return None
    if code == 7:  # combined.py:497-515
        return "MVS polya check failed: " + "".join(n for i, n in enumerate(_MVS_NAMES) if mask >> i & 1).rstrip()
    return _FAIL_TEXT[code]


_TEMPLATE = {k: None for k in FIELD_ORDER}


def _new_result(success: bool):
    """DetectResults(success=...) without running the 56-argument dataclass __init__ (all other fields None)"""
    d = object.__new__(DetectResults)
    d.__dict__.update(_TEMPLATE)
    d.__dict__["success"] = success
    return d


def records_to_results(recs: np.ndarray, primary_method: int, llr_log: Optional[str],
                       open_pore_overflow: Optional[dict] = None) -> List[Any]:
    """adb_record[N] -> list[DetectResults].  Columns are pulled out of the structured array once (python lists / typed
    numpy columns); the per-read loop only builds the objects.  ``open_pore_overflow`` maps a record index to the full
    open-pore list of a read whose list does not fit the record (``detect.open_pore_overflow``); without it such a
    read raises instead of being truncated."""
    out = []
    method = _METHOD[primary_method]
    n = len(recs)
    if n == 0:
        return out
    col = {k: recs[k].tolist() for k in ("valid", "fail_code", "mvs_fail_mask", "success", "adapter_start", "adapter_end",
                                         "polya_end", "preloaded", "n_cand", "n_open_pores", "mvs_adapter_end", "sp_flag")}
    signal_len = recs["signal_len"].astype(np.int32)
    a_end64, p_end64 = recs["adapter_end"].astype(np.int64), recs["polya_end"].astype(np.int64)
    prim_ae, prim_pe = recs["primary_adapter_end"].astype(np.int64), recs["primary_polya_end"].astype(np.int64)
    stats = recs["stats"].tolist()
    mvs = recs["mvs"].tolist()
    real32 = recs["real"][:, :2].astype(np.float32)
    real_lr = recs["real"][:, 2].astype(np.float64)
    med_shift = recs["med_shift"].astype(np.float32)
    cand = recs["cand"].astype(np.int64)
    open_pores = recs["open_pores"].astype(np.int64)
    sp_idx, sp_next = recs["sp_idx"].astype(np.int64), recs["sp_next_idx"].astype(np.int64)
    sp_op = recs["sp_open_pore_idx"].astype(np.int64)
    sp_pa, sp_next_pa = recs["sp_pa"].astype(np.float32), recs["sp_next_pa"].astype(np.float32)
    cap_op = open_pores.shape[1]
    ae_key, pe_key = f"{method}_adapter_end", f"{method}_polya_end"
    parts = (("adapter", V_ADAPTER), ("polya", V_POLYA), ("rna_preloaded", V_RNA))
    for i in range(n):
        valid = col["valid"][i]
        reason = fail_reason_text(col["fail_code"][i], col["mvs_fail_mask"][i])
        if not valid & V_FIELDS:  # died on an exception: DetectResults(success=False, fail_reason=str(e))
            out.append(DetectResults(success=False, fail_reason=reason))
            continue
        a_start, a_end, p_end, size = col["adapter_start"][i], col["adapter_end"][i], col["polya_end"][i], col["preloaded"][i]
        d = _new_result(bool(col["success"][i]))
        dd = d.__dict__
        dd["signal_len"] = signal_len[i]
        dd["preloaded"] = size
        dd["adapter_end"] = a_end64[i]
        # mvs_detect_overwrite can leave polya_end = trace_early_stop_pos = None (combined.py:560-562)
        polya_none = bool(valid & V_POLYA_NONE)
        dd["polya_end"] = None if polya_none else p_end64[i]
        dd["fail_reason"] = reason
        dd["llr_detect_log"] = llr_log
        dd["mvs_llr_polya_end_adjust_ignored"] = False
        dd["mvs_llr_polya_end_to_early_stop"] = bool(valid & V_TO_EARLY_STOP)
        if valid & V_MVS_ADAPTER_END:
            dd["mvs_adapter_end"] = col["mvs_adapter_end"][i]
        # partitions (signal_partitions.py:65-96)
        bounds = ((a_start, a_end), (a_end, p_end), (p_end, size))
        st = stats[i]
        for idx, (name, bit) in enumerate(parts):
            s0, s1 = bounds[idx]
            dd[name + "_start"] = None if (polya_none and idx == 2) else s0
            if valid & bit:
                dd[name + "_len"] = s1 - s0
                q = st[idx]
                dd[name + "_mean"], dd[name + "_std"], dd[name + "_med"], dd[name + "_mad"] = q[0], q[1], q[2], q[3]
        if valid & V_CAND:
            dd["polya_candidates"] = cand[i, : col["n_cand"][i]].copy()
        dd[ae_key] = prim_ae[i]
        dd[pe_key] = prim_pe[i]
        if valid & V_MVS:
            m = mvs[i]
            (dd["mvs_detect_mean_at_loc"], dd["mvs_detect_var_at_loc"], dd["mvs_detect_polya_med"],
             dd["mvs_detect_polya_local_range"], dd["mvs_detect_med_shift"]) = m[0], m[1], m[2], m[3], m[4]
        if valid & V_REAL_MEANS:
            dd["real_adapter_mean_start"] = real32[i, 0]
            dd["real_adapter_mean_end"] = real32[i, 1]
        if valid & V_REAL_RANGE:
            dd["real_adapter_local_range"] = real_lr[i]
        if valid & V_OPEN:
            n_op = col["n_open_pores"][i]
            if n_op > cap_op:
                full = None if open_pore_overflow is None else open_pore_overflow.get(i)
                if full is None or len(full) != n_op:
                    raise OverflowError(
                        f"read has {n_op} open-pore runs; the record keeps {cap_op} (ADB_MAX_OPEN_PORES) and no "
                        "overflow list was supplied (adapted_b200.detect.open_pore_overflow)")
                dd["open_pores"] = np.asarray(full, dtype=np.int64).copy()
            else:
                dd["open_pores"] = open_pores[i, :n_op].copy()
        if valid & V_MEDSHIFT:
            dd["adapter_rna_median_shift"] = med_shift[i]
        if valid & V_SP:
            dd["start_peak_idx"] = sp_idx[i]
            dd["start_peak_pa"] = sp_pa[i]
            dd["start_peak_next_max_idx"] = sp_next[i]
            dd["start_peak_next_max_pa"] = sp_next_pa[i]
            if valid & V_SP_OPEN:
                dd["start_peak_open_pore_idx"] = sp_op[i]
                dd["start_peak_open_pore_type"] = _SP_FLAGS.get(col["sp_flag"][i])
                # combined.py:340-347: a flagged read fails; the type is appended only if it had failed already
                if col["fail_code"][i] != 0 and dd["start_peak_open_pore_type"] is not None:
                    dd["fail_reason"] = reason + "+" + dd["start_peak_open_pore_type"]
        out.append(d)
    return out
