"""The validation / CNN-prep / convolution-epilogue kernels of round 2 against the kernels they replaced.

Every replacement keeps its predecessor selectable (context options `hist_validate`, `no_fast_validate`; environment
switches read at launch time): the records of both sides must be BYTE-IDENTICAL -- the order statistics are exact in all
of them, the partition sums are the same integers through the same float64 formulas -- and the reads the tensor-core
histogram cannot settle (codes outside its range) must come back from the general kernel with the same bytes too.

References: adapted/detect/combined.py:358-631 (validate_boundaries), adapted/detect/cnn.py:70-82 (prepare_data),
adapted/detect/cnn.py:16-52 (BoundariesCNN, the transposed convolution of the head)."""
import os

import numpy as np
import pytest

from adapted_b200.config import get_chemistry_specific_config
from adapted_b200.records import records_to_results
from adapted_b200.synth import ReadBatch, make_reads
from tests.golden_io import load_cnn_weights
from tests.helpers import diff_results

pytestmark = pytest.mark.gpu

DEFAULTS = {"hist_validate": 1, "no_fast_validate": 0, "no_cand_followup": 0, "exact_global_select": 0}


def _records(b, spc, mbs, model=None, opts=None, env=None):
    from adapted_b200 import _lib
    from adapted_b200.detect import detect_reads

    ctx = _lib.default_context(0)
    opts, env = opts or {}, env or {}
    for k, v in opts.items():
        ctx.set_option(k, v)
    for k, v in env.items():
        os.environ[k] = v
    try:
        recs, st = detect_reads(b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, spc, model=model,
                                minibatch_size=mbs, return_records=True)
        handed = ctx.query("validate_handovers")
    finally:
        for k in opts:
            ctx.set_option(k, DEFAULTS[k])
        for k in env:
            os.environ.pop(k, None)
    assert not st.any()
    return np.asarray(recs).copy(), handed


def _results(recs, chem):
    return records_to_results(recs, 1 if chem == "rna004" else 0, None)


def _bytes(recs):
    return np.ascontiguousarray(recs).view(np.uint8).reshape(len(recs), -1)


def _same_bytes(a, b):
    bad = np.flatnonzero((_bytes(a) != _bytes(b)).any(axis=1))
    return [] if bad.size == 0 else [f"record {i} differs" for i in bad[:5]]


def _shift_pa(b, delta_pa):
    """the same ADC codes under a calibration that moves every pA value by delta_pa"""
    coff = (b.calib_offset + np.float32(delta_pa) / b.calib_scale).astype(np.float32)
    return ReadBatch(adc=b.adc, offsets=b.offsets, full_lens=b.full_lens, calib_offset=coff, calib_scale=b.calib_scale,
                     truth=b.truth, m=b.m)


@pytest.mark.parametrize("chem,kw", [("rna002", {}), ("rna002", {"stress": True}), ("rna004", {}), ("rna004", {"stress": True})])
def test_histogram_validation_equals_counting_validation(chem, kw):
    """validate_hist_kernel (tensor-core histograms) == validate_fast_kernel (counting passes), byte for byte; both agree
    with the general validate_kernel on every field (floats of the partition sums within the 1e-5 contract)"""
    spc = get_chemistry_specific_config(chem)
    model = load_cnn_weights() if chem == "rna004" else None
    b = make_reads(2000, chem, spc.sig_preload_size, seed=8101 + len(kw), **kw)
    hist, _ = _records(b, spc, 1000, model)
    counting, _ = _records(b, spc, 1000, model, opts={"hist_validate": 0})
    assert _same_bytes(hist, counting) == []
    general, _ = _records(b, spc, 1000, model, opts={"no_fast_validate": 1})
    assert diff_results(_results(hist, chem), _results(general, chem), exact_floats=False) == []


@pytest.mark.parametrize("delta,must_hand_over", [(-70.0, True), (-45.0, False), (95.0, True)])
def test_reads_outside_the_histogram_range_are_settled_by_the_general_kernel(delta, must_hand_over):
    """the histogram covers 25 .. 205 pA; with the whole signal moved down (adapter at 10 pA) or up (RNA beyond 205 pA)
    medians and MADs fall into its end bins: those reads must be handed over and come back with the bytes of the
    counting kernel.  At -45 pA only tails of the signal leave the range: no decisive probe touches an end bin, the
    histogram kernel settles the reads itself"""
    spc = get_chemistry_specific_config("rna002")
    b = _shift_pa(make_reads(300, "rna002", spc.sig_preload_size, seed=8111), delta)
    hist, handed = _records(b, spc, 300)
    counting, _ = _records(b, spc, 300, opts={"hist_validate": 0})
    assert _same_bytes(hist, counting) == []
    assert (handed > 0) or not must_hand_over  # the range really was left


def test_negative_and_mixed_sign_codes_go_through_the_histogram_kernel():
    """the counting kernel needs non-negative codes (fp16 bit patterns), the histogram kernel takes any int16: codes moved
    far below zero with the calibration compensating must give the records of the unshifted reads where the pA values
    are bit-identical, and the general kernel's otherwise"""
    spc = get_chemistry_specific_config("rna002")
    b = make_reads(200, "rna002", spc.sig_preload_size, seed=8121)
    adc = b.adc.astype(np.int32) - 600          # signal codes straddle zero
    coff = (b.calib_offset + np.float32(600.0)).astype(np.float32)
    b2 = ReadBatch(adc=adc.astype(np.int16), offsets=b.offsets, full_lens=b.full_lens, calib_offset=coff,
                   calib_scale=b.calib_scale, truth=b.truth, m=b.m)
    hist, _ = _records(b2, spc, 200)
    general, _ = _records(b2, spc, 200, opts={"no_fast_validate": 1})
    assert diff_results(_results(hist, "rna002"), _results(general, "rna002"), exact_floats=False) == []


def test_cnn_prep_warp_kernel_equals_cta_kernel():
    """cnn_prep_warp_kernel (warp per read, TMA-staged chunks, warp-level selects) == cnn_prep_kernel (CTA per read)"""
    spc = get_chemistry_specific_config("rna004")
    model = load_cnn_weights()
    b = make_reads(1500, "rna004", spc.sig_preload_size, seed=8131, short_frac=0.2)
    keep = b.full_lens >= spc.core.min_obs_adapter + 2 * spc.core.downscale_factor
    idx = np.flatnonzero(keep)
    chunks, offs = [], [0]
    for i in idx:
        chunks.append(b.adc[b.offsets[i]:b.offsets[i + 1]])
        offs.append(offs[-1] + len(chunks[-1]))
    b = ReadBatch(adc=np.concatenate(chunks), offsets=np.asarray(offs, np.int64), full_lens=b.full_lens[idx],
                  calib_offset=b.calib_offset[idx], calib_scale=b.calib_scale[idx], truth=b.truth[idx], m=b.m)
    warp, _ = _records(b, spc, 500, model)
    cta, _ = _records(b, spc, 500, model, env={"ADB_PREP_CTA": "1"})
    assert _same_bytes(warp, cta) == []


def test_transposed_convolution_weights_from_the_constant_bank():
    """layer 3's epilogue with its weights in the constant bank == the same epilogue with the weights in shared memory"""
    spc = get_chemistry_specific_config("rna004")
    model = load_cnn_weights()
    b = make_reads(1000, "rna004", spc.sig_preload_size, seed=8141)
    const, _ = _records(b, spc, 500, model)
    shared, _ = _records(b, spc, 500, model, env={"ADB_NO_CONST_CONVT": "1"})
    assert _same_bytes(const, shared) == []
