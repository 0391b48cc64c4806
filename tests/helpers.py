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
