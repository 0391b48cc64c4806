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


def records_to_results(recs: np.ndarray, primary_method: int, llr_log: Optional[str]) -> List[Any]:
    """adb_record[N] -> list[DetectResults]."""
    out = []
    method = _METHOD[primary_method]
    for r in recs:
        valid = int(r["valid"])
        reason = fail_reason_text(int(r["fail_code"]), int(r["mvs_fail_mask"]))
        if not valid & V_FIELDS:  # died on an exception: DetectResults(success=False, fail_reason=str(e))
            out.append(DetectResults(success=False, fail_reason=reason))
            continue
        a_start, a_end, p_end = int(r["adapter_start"]), int(r["adapter_end"]), int(r["polya_end"])
        size = int(r["preloaded"])
        d = DetectResults(success=bool(r["success"]))
        d.signal_len = np.int32(r["signal_len"])
        d.preloaded = size
        d.adapter_end = np.int64(a_end)
        # mvs_detect_overwrite can leave polya_end = trace_early_stop_pos = None (combined.py:560-562)
        polya_none = bool(valid & V_POLYA_NONE)
        d.polya_end = None if polya_none else np.int64(p_end)
        d.fail_reason = reason
        d.llr_detect_log = llr_log
        d.mvs_llr_polya_end_adjust_ignored = False
        d.mvs_llr_polya_end_to_early_stop = bool(valid & V_TO_EARLY_STOP)
        if valid & V_MVS_ADAPTER_END:
            d.mvs_adapter_end = int(r["mvs_adapter_end"])
        # partitions (signal_partitions.py:65-96)
        for idx, (name, s0, s1, bit) in enumerate((("adapter", a_start, a_end, V_ADAPTER),
                                                    ("polya", a_end, p_end, V_POLYA),
                                                    ("rna_preloaded", p_end, size, V_RNA))):
            setattr(d, f"{name}_start", None if (polya_none and name == "rna_preloaded") else s0)
            if valid & bit:
                setattr(d, f"{name}_len", s1 - s0)
                for q, key in enumerate(("mean", "std", "med", "mad")):
                    setattr(d, f"{name}_{key}", float(r["stats"][idx][q]))
        if valid & V_CAND:
            d.polya_candidates = np.asarray(r["cand"][: int(r["n_cand"])], dtype=np.int64)
        setattr(d, f"{method}_adapter_end", np.int64(r["primary_adapter_end"]))
        setattr(d, f"{method}_polya_end", np.int64(r["primary_polya_end"]))
        if valid & V_MVS:
            (d.mvs_detect_mean_at_loc, d.mvs_detect_var_at_loc, d.mvs_detect_polya_med,
             d.mvs_detect_polya_local_range, d.mvs_detect_med_shift) = (float(v) for v in r["mvs"])
        if valid & V_REAL_MEANS:
            d.real_adapter_mean_start = np.float32(r["real"][0])
            d.real_adapter_mean_end = np.float32(r["real"][1])
        if valid & V_REAL_RANGE:
            d.real_adapter_local_range = np.float64(r["real"][2])
        if valid & V_OPEN:
            n_op = int(r["n_open_pores"])
            if n_op > r["open_pores"].size:
                raise OverflowError(
                    f"read has {n_op} open-pore runs; the record keeps {r['open_pores'].size} "
                    "(ADB_MAX_OPEN_PORES) -- rebuild with a larger cap")
            d.open_pores = np.asarray(r["open_pores"][:n_op], dtype=np.int64)
        if valid & V_MEDSHIFT:
            d.adapter_rna_median_shift = np.float32(r["med_shift"])
        if valid & V_SP:
            d.start_peak_idx = np.int64(r["sp_idx"])
            d.start_peak_pa = np.float32(r["sp_pa"])
            d.start_peak_next_max_idx = np.int64(r["sp_next_idx"])
            d.start_peak_next_max_pa = np.float32(r["sp_next_pa"])
            if valid & V_SP_OPEN:
                d.start_peak_open_pore_idx = np.int64(r["sp_open_pore_idx"])
                d.start_peak_open_pore_type = _SP_FLAGS.get(int(r["sp_flag"]))
                # combined.py:340-347: a flagged read fails; the type is appended only if it had failed already
                if int(r["fail_code"]) != 0:
                    d.fail_reason = reason + "+" + d.start_peak_open_pore_type
        out.append(d)
    return out
