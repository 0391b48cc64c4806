"""the oracle restatement of the streaming poly(A) detector (mean_var_shift_polyA_detect,
adapted/detect/mvs.py:341-426) reproduces what the executed reference returned (tests/golden/mvs_stream.json.gz)."""
import numpy as np
import pytest

from oracle import detect_ref
from tests.golden_io import load_stream_cases

CASES = load_stream_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_stream_detector_reproduces_reference(case):
    b = case["batch"]
    x = b.to_dense_pa()
    lens = np.minimum(b.full_lens, case["m"])
    stats = {}
    got = [detect_ref.mvs_stream_detect(x[i, : lens[i]], case["params"], stats) for i in range(b.n)]
    assert np.array_equal(got, case["want"])
    if case["name"] == "retry_loop":
        assert stats.get("rejected", 0) > 0, "the case is meant to exercise the offset loop"
