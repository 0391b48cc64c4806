"""Differential test of the oracle against the LIVE reference (needs /root/reference: build container
only; skipped on the GPU box).  Fresh seeds, every DetectResults field compared exactly."""
import logging
import warnings

import numpy as np
import pytest

from adapted_b200.config import flatten_config, get_chemistry_specific_config, start_peak_config
from adapted_b200.synth import make_reads
from oracle import build_ref, detect_ref
from tests.helpers import diff_results

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    build_ref.import_reference()
    logging.disable(logging.CRITICAL)
    warnings.simplefilter("ignore")
    from oracle.make_golden import reference_configs

    return reference_configs()


@pytest.mark.parametrize("chem", ["rna002", "rna004"])
def test_presets_match_reference_toml(ref, chem):
    assert flatten_config(get_chemistry_specific_config(chem)) == flatten_config(ref[("plain", chem)])
    assert flatten_config(start_peak_config(chem)) == flatten_config(ref[("start_peak", chem)])


@pytest.mark.parametrize("seed,kw", [(101, {}), (102, {"stress": True})])
def test_llr2_matches_reference(ref, seed, kw):
    from adapted.detect.combined import combined_detect_llr2

    spc = get_chemistry_specific_config("rna002")
    b = make_reads(40, "rna002", spc.sig_preload_size, seed=seed, **kw)
    x = b.to_dense_pa()
    want = combined_detect_llr2(x, b.full_lens, ref[("plain", "rna002")])
    got = detect_ref.detect_llr2(x, b.full_lens, spc)
    assert diff_results(got, want, exact_floats=True) == []


@pytest.mark.parametrize("seed,kw", [(201, {}), (202, {"short_frac": 0.3}), (203, {"stress": True})])
def test_cnn_matches_reference(ref, seed, kw):
    from adapted.detect.cnn import load_cnn_model
    from adapted.detect.combined import combined_detect_cnn

    spc_ref = ref[("plain", "rna004")]
    model = load_cnn_model(spc_ref.cnn_boundaries.model_name)
    w = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    spc = get_chemistry_specific_config("rna004")
    b = make_reads(60, "rna004", spc.sig_preload_size, seed=seed, **kw)
    x = b.to_dense_pa()
    want = combined_detect_cnn(x.copy(), b.full_lens, model, spc_ref)
    got = detect_ref.detect_cnn(x.copy(), b.full_lens, w, spc)
    assert diff_results(got, want, exact_floats=True) == []


@pytest.mark.parametrize("seed,kw", [(301, {}), (302, {"short_frac": 0.3})])
def test_start_peak_matches_reference(ref, seed, kw):
    from adapted.detect.combined import combined_detect_start_peak

    spc = start_peak_config("rna004")
    b = make_reads(40, "rna004", spc.sig_preload_size, seed=seed, **kw)
    x = b.to_dense_pa()
    want = combined_detect_start_peak(x, b.full_lens, ref[("start_peak", "rna004")])
    got = detect_ref.detect_start_peak(x, b.full_lens, spc)
    assert diff_results(got, want, exact_floats=True) == []
