"""Pin the oracle: the CPU restatement (oracle/detect_ref.py) must reproduce, bit for bit, what the
executed reference returned for the committed golden minibatches (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import detect_ref
from tests.golden_io import SEAM_CASES, load_case, load_cnn_weights
from tests.helpers import diff_results


def _run(seam, rec):
    x = rec["batch"].to_dense_pa()
    lens = rec["batch"].full_lens
    if seam == "llr2":
        return detect_ref.detect_llr2(x, lens, rec["spc"])
    if seam == "cnn":
        return detect_ref.detect_cnn(x, lens, load_cnn_weights(), rec["spc"])
    return detect_ref.detect_start_peak(x, lens, rec["spc"])


@pytest.mark.parametrize("seam,name", [(s, n) for s, names in SEAM_CASES.items() for n in names])
def test_oracle_reproduces_reference_golden(seam, name):
    rec = load_case(name)
    if "raises" in rec:
        with pytest.raises(Exception) as ei:
            _run(seam, rec)
        assert type(ei.value).__name__ == rec["raises"]["type"]
        assert str(ei.value) == rec["raises"]["message"]
        return
    got = _run(seam, rec)
    assert diff_results(got, rec["results"], exact_floats=True) == []


def test_c_llr_trace_golden():
    with np.load("tests/golden/c_llr_trace.npz") as z:
        for i in range(6):
            x = z[f"x{i}"]
            for tag in ("full", "aes", "pes", "tail"):
                a = [int(v) for v in z[f"a{i}_{tag}"]]
                g = detect_ref.llr_trace(x, *a, 0)
                assert np.array_equal(g, z[f"g{i}_{tag}"], equal_nan=True), (i, tag)
