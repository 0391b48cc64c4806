"""Parity of the GPU LLR path (combined_detect_llr2 drop-in) against the oracle and the reference goldens."""
import numpy as np
import pytest

from adapted_b200.config import get_chemistry_specific_config
from adapted_b200.synth import make_reads
from oracle import detect_ref
from tests.golden_io import load_case
from tests.helpers import diff_results

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["llr_rna002_basic", "llr_rna002_stress"])
def test_llr2_golden(name):
    from adapted_b200.detect import combined_detect_llr2

    rec = load_case(name)
    x = rec["batch"].to_dense_pa()
    got = combined_detect_llr2(x, rec["batch"].full_lens, rec["spc"])
    assert diff_results(got, rec["results"]) == []


def test_llr2_lost_minibatch_raises_like_reference():
    from adapted_b200.detect import combined_detect_llr2

    rec = load_case("llr_rna002_lost_minibatch")
    x = rec["batch"].to_dense_pa()
    with pytest.raises(ValueError) as ei:
        combined_detect_llr2(x, rec["batch"].full_lens, rec["spc"])
    assert str(ei.value) == rec["raises"]["message"]


@pytest.mark.parametrize("seed,kw", [(401, {}), (402, {"stress": True})])
def test_llr2_dense_f32_matches_oracle(seed, kw):
    from adapted_b200.detect import combined_detect_llr2

    spc = get_chemistry_specific_config("rna002")
    b = make_reads(160, "rna002", spc.sig_preload_size, seed=seed, **kw)
    x = b.to_dense_pa()
    got = combined_detect_llr2(x, b.full_lens, spc)
    want = detect_ref.detect_llr2(x, b.full_lens, spc)
    assert diff_results(got, want) == []


def test_llr2_ragged_i16_ingest_matches_oracle():
    """native ingest (int16 ADC + calibration on the device) == oracle on the float32 matrix the same
    calibration produces on the host; two minibatches in one call."""
    from adapted_b200.detect import detect_reads

    spc = get_chemistry_specific_config("rna002")
    b = make_reads(120, "rna002", spc.sig_preload_size, seed=403)
    got, status = detect_reads(b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, spc, minibatch_size=60)
    assert not status.any()
    x = b.to_dense_pa()
    want = detect_ref.detect_llr2(x[:60], b.full_lens[:60], spc) + detect_ref.detect_llr2(x[60:], b.full_lens[60:], spc)
    assert diff_results(got, want) == []


def test_llr2_mad_zero_raises():
    from adapted_b200.detect import combined_detect_llr2

    spc = get_chemistry_specific_config("rna002")
    x = np.full((3, spc.sig_preload_size), 80.0, dtype=np.float32)
    with pytest.raises(ValueError, match="MAD normalization failed: scale is 0"):
        combined_detect_llr2(x, np.full(3, 30000, np.int32), spc)
