"""Load tests/golden/*.json.gz (written by oracle/make_golden.py from the executed reference)."""
from __future__ import annotations

import gzip
import hashlib
import json
import os
from typing import Any, Dict

import numpy as np

from adapted_b200.config import config_from_dict
from adapted_b200.synth import make_reads

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SEAM_CASES = {
    "llr2": ["llr_rna002_basic", "llr_rna002_stress", "llr_rna002_lost_minibatch", "llr_rna002_overwrite",
             "llr_rna002_overwrite_stress"],
    "cnn": ["cnn_rna004_basic", "cnn_rna004_short", "cnn_rna004_stress", "cnn_rna004_overwrite",
            "cnn_rna004_overwrite_short"],
    "start_peak": ["start_peak_rna004_basic", "start_peak_rna004_poisoned"],
}


def _decode(v):
    if isinstance(v, dict) and "array" in v:
        return np.asarray(v["array"], dtype=v["dtype"])
    return v


def load_case(name: str) -> Dict[str, Any]:
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rb") as f:
        rec = json.loads(f.read().decode())
    if "results" in rec:
        rec["results"] = [{k: _decode(v) for k, v in r.items()} for r in rec["results"]]
    cfg = rec["config"]
    for sec in cfg.values():
        for k, v in sec.items():
            if isinstance(v, list):
                sec[k] = tuple(v)
    rec["spc"] = config_from_dict(cfg)
    batch = make_reads(rec["n"], rec["chemistry"], rec["m"], seed=rec["seed"], **rec["gen_kwargs"])
    sha = hashlib.sha256(batch.adc.tobytes()).hexdigest()
    assert sha == rec["adc_sha256"], "synthetic generator drifted: golden inputs cannot be regenerated"
    rec["batch"] = batch
    return rec


def load_cnn_weights() -> Dict[str, np.ndarray]:
    with np.load(os.path.join(GOLDEN, "cnn_weights_rna004_130bps_v0.2.4.npz")) as z:
        return {k: z[k] for k in z.files}


def load_stream_cases():
    """tests/golden/mvs_stream.json.gz: expected poly(A) starts of mean_var_shift_polyA_detect (executed reference)."""
    from adapted_b200.config import StreamingConfig

    with gzip.open(os.path.join(GOLDEN, "mvs_stream.json.gz"), "rb") as f:
        doc = json.loads(f.read().decode())
    out = []
    for c in doc["cases"]:
        p = StreamingConfig()
        for k, v in c["overrides"].items():
            setattr(p, k, tuple(v) if isinstance(v, list) else v)
        batch = make_reads(c["n"], c["chemistry"], c["m"], seed=c["seed"], **c["gen_kwargs"])
        assert hashlib.sha256(batch.adc.tobytes()).hexdigest() == c["adc_sha256"], "synthetic generator drifted"
        out.append(dict(name=c["name"], params=p, batch=batch, m=c["m"], want=np.asarray(c["polya_start"], dtype=np.int64)))
    return out
