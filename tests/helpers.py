"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import math
from typing import Any, Dict, Iterable, List

import numpy as np

INT_FIELDS = (
    "signal_len preloaded adapter_start adapter_end adapter_len polya_start polya_end polya_len "
    "rna_preloaded_start rna_preloaded_len llr_adapter_end llr_polya_end cnn_adapter_end cnn_polya_end "
    "start_peak_adapter_end start_peak_polya_end start_peak_idx start_peak_next_max_idx "
    "start_peak_open_pore_idx mvs_adapter_end llr_adapter_end_adjust llr_polya_end_adjust "
    "llr_trace_early_stop_pos"
).split()
FLOAT_FIELDS = (
    "adapter_mean adapter_std adapter_med adapter_mad polya_mean polya_std polya_med polya_mad "
    "rna_preloaded_mean rna_preloaded_std rna_preloaded_med rna_preloaded_mad start_peak_pa "
    "start_peak_next_max_pa adapter_rna_median_shift mvs_detect_mean_at_loc mvs_detect_var_at_loc "
    "mvs_detect_polya_med mvs_detect_polya_local_range mvs_detect_med_shift real_adapter_mean_start "
    "real_adapter_mean_end real_adapter_local_range"
).split()
OTHER_FIELDS = (
    "success polya_truncated polya_candidates start_peak_open_pore_type llr_trace "
    "mvs_llr_polya_end_adjust_ignored mvs_llr_polya_end_to_early_stop open_pores fail_reason llr_detect_log"
).split()
ALL_FIELDS = INT_FIELDS + FLOAT_FIELDS + OTHER_FIELDS

FLOAT_RTOL = 1e-5  # north_star: float statistics within 1e-5 relative


def as_dict(r) -> Dict[str, Any]:
    return r if isinstance(r, dict) else r.to_dict()


def field_equal(name: str, a, b, exact_floats: bool = False) -> bool:
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, str) or isinstance(b, str):
        return a == b
    if name in FLOAT_FIELDS:
        a, b = float(a), float(b)
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        if exact_floats:
            return a == b
        return abs(a - b) <= FLOAT_RTOL * max(abs(a), abs(b)) or a == b
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.array_equal(a, b))


def diff_results(got: Iterable, want: Iterable, exact_floats: bool = False, limit: int = 20) -> List[str]:
    """Field-by-field differences between two result lists (dicts or DetectResults-like objects)."""
    out: List[str] = []
    got, want = list(got), list(want)
    if len(got) != len(want):
        return [f"length {len(got)} != {len(want)}"]
    for i, (g, w) in enumerate(zip(got, want)):
        g, w = as_dict(g), as_dict(w)
        for k in ALL_FIELDS:
            if not field_equal(k, g.get(k), w.get(k), exact_floats):
                out.append(f"read {i} field {k}: got {g.get(k)!r} want {w.get(k)!r}")
                if len(out) >= limit:
                    return out
    return out


def results_to_records(results, primary_method: int):
    """Inverse of adapted_b200.records.records_to_results: reference-style result dicts -> adb_record array.
    Lets the CPU tests feed the native table writer with exactly what the executed reference returned."""
    from adapted_b200 import _lib
    from adapted_b200 import records as R

    codes = {v: k for k, v in R._FAIL_TEXT.items()}
    sp_flags = {v: k for k, v in R._SP_FLAGS.items()}
    method = R._METHOD[primary_method]
    recs = np.zeros(len(results), dtype=_lib.RECORD_DTYPE)
    for rec, r in zip(recs, results):
        r = as_dict(r)
        reason = r.get("fail_reason")
        sp_type = r.get("start_peak_open_pore_type")
        base = reason
        if reason is not None and sp_type is not None and reason.endswith("+" + sp_type):
            base = reason[: -len(sp_type) - 1]
        if base is None:
            rec["fail_code"] = 0
        elif base.startswith("MVS polya check failed: ") and base not in codes:
            rec["fail_code"] = 7
            names = base[len("MVS polya check failed: "):].split()
            rec["mvs_fail_mask"] = sum(1 << i for i, n in enumerate(("mean", "var", "med", "range", "shift")) if n in names)
        else:
            rec["fail_code"] = codes[base]
        rec["success"] = int(bool(r["success"]))
        if r.get("signal_len") is None:  # DetectResults(success=False, fail_reason=str(e))
            continue
        valid = R.V_FIELDS
        rec["signal_len"], rec["preloaded"] = r["signal_len"], r["preloaded"]
        rec["adapter_start"], rec["adapter_end"] = r["adapter_start"], r["adapter_end"]
        if r["polya_end"] is None:
            valid |= R.V_POLYA_NONE
        else:
            rec["polya_end"] = r["polya_end"]
        for idx, (name, bit) in enumerate((("adapter", R.V_ADAPTER), ("polya", R.V_POLYA), ("rna_preloaded", R.V_RNA))):
            if r.get(f"{name}_len") is not None:
                valid |= bit
                for q, key in enumerate(("mean", "std", "med", "mad")):
                    v = r[f"{name}_{key}"]
                    rec["stats"][idx][q] = np.nan if v is None else v
        if r.get("polya_candidates") is not None:
            valid |= R.V_CAND
            c = np.asarray(r["polya_candidates"])
            rec["n_cand"] = c.size
            rec["cand"][: c.size] = c
        rec["primary_adapter_end"] = r[f"{method}_adapter_end"]
        rec["primary_polya_end"] = r[f"{method}_polya_end"]
        if r.get("mvs_detect_mean_at_loc") is not None:
            valid |= R.V_MVS
            rec["mvs"][:] = [r[f"mvs_detect_{k}"] for k in ("mean_at_loc", "var_at_loc", "polya_med", "polya_local_range", "med_shift")]
        if r.get("mvs_adapter_end") is not None:
            valid |= R.V_MVS_ADAPTER_END
            rec["mvs_adapter_end"] = r["mvs_adapter_end"]
        if r.get("mvs_llr_polya_end_to_early_stop"):
            valid |= R.V_TO_EARLY_STOP
        if r.get("real_adapter_mean_start") is not None:
            valid |= R.V_REAL_MEANS
            rec["real"][0], rec["real"][1] = r["real_adapter_mean_start"], r["real_adapter_mean_end"]
        if r.get("real_adapter_local_range") is not None:
            valid |= R.V_REAL_RANGE
            rec["real"][2] = r["real_adapter_local_range"]
        if r.get("open_pores") is not None:
            valid |= R.V_OPEN
            op = np.asarray(r["open_pores"])
            rec["n_open_pores"] = op.size
            rec["open_pores"][: min(op.size, rec["open_pores"].size)] = op[: rec["open_pores"].size]
        if r.get("adapter_rna_median_shift") is not None:
            valid |= R.V_MEDSHIFT
            rec["med_shift"] = r["adapter_rna_median_shift"]
        if r.get("start_peak_idx") is not None:
            valid |= R.V_SP
            rec["sp_idx"], rec["sp_pa"] = r["start_peak_idx"], r["start_peak_pa"]
            rec["sp_next_idx"], rec["sp_next_pa"] = r["start_peak_next_max_idx"], r["start_peak_next_max_pa"]
            if r.get("start_peak_open_pore_idx") is not None:
                valid |= R.V_SP_OPEN
                rec["sp_open_pore_idx"] = r["start_peak_open_pore_idx"]
                rec["sp_flag"] = sp_flags.get(sp_type, 0)
        rec["valid"] = valid
    return recs


def assert_csv_equivalent(got: str, want: str, moved_ok=None) -> int:
    """Two boundary tables cell by cell: integer, boolean, text and array cells identical; float cells within the 1e-5
    contract before rounding, i.e. at most one unit of the third decimal.  `moved_ok(header, got_row, want_row)` may
    accept a whole row that differs for a tolerated reason (CNN primaries moved by one step); returns how many rows
    were accepted that way."""
    import csv
    import io

    if got == want:
        return 0
    g = list(csv.reader(io.StringIO(got)))
    w = list(csv.reader(io.StringIO(want)))
    assert g[0] == w[0] and len(g) == len(w), (g[0] == w[0], len(g), len(w))
    hdr, moved = g[0], 0
    for gr, wr in zip(g[1:], w[1:]):
        bad = []
        for col, a, b in zip(hdr, gr, wr):
            if a == b:
                continue
            if col in FLOAT_FIELDS and a and b and abs(float(a) - float(b)) <= 0.0011 + 1e-5 * abs(float(b)):
                continue
            bad.append((col, a, b))
        if bad:
            if moved_ok is not None and moved_ok(hdr, gr, wr):
                moved += 1
            else:
                raise AssertionError(f"row {gr[0]}: {bad[:4]}")
    return moved


class _DictResult:
    """oracle result dict with the to_dict() of a DetectResults (for the pandas restatement of the writer)"""

    def __init__(self, d):
        self._d = d

    def to_dict(self):
        return dict(self._d)
