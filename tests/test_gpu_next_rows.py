"""SURVEY rows f2 / f3 on the GPU: the mvs_detect_overwrite branch of validate_boundaries (combined.py:517-562,
mean_var_shift_polyA_detect_at_loc mvs.py:181-338) and the native result tables written from GPU records."""
import csv
import io

import numpy as np
import pytest

from adapted_b200.config import config_as_dict, config_from_dict, get_chemistry_specific_config
from adapted_b200.synth import make_reads
from oracle import detect_ref
from tests.golden_io import load_case, load_cnn_weights
from tests.helpers import FLOAT_FIELDS, diff_results
from tests.test_gpu_cnn_path import _cnn_compare

pytestmark = pytest.mark.gpu


def _overwrite_config(chem):
    d = config_as_dict(get_chemistry_specific_config(chem))
    d["mvs_polya"]["mvs_detect_overwrite"] = True
    return config_from_dict(d)


@pytest.mark.parametrize("name", ["llr_rna002_overwrite", "llr_rna002_overwrite_stress"])
def test_llr2_overwrite_golden(name):
    from adapted_b200.detect import combined_detect_llr2

    rec = load_case(name)
    got = combined_detect_llr2(rec["batch"].to_dense_pa(), rec["batch"].full_lens, rec["spc"])
    assert diff_results(got, rec["results"]) == []


@pytest.mark.parametrize("name", ["cnn_rna004_overwrite", "cnn_rna004_overwrite_short"])
def test_cnn_overwrite_golden(name):
    from adapted_b200.detect import combined_detect_cnn

    rec = load_case(name)
    got = combined_detect_cnn(rec["batch"].to_dense_pa(), rec["batch"].full_lens, load_cnn_weights(), rec["spc"])
    _cnn_compare(got, rec["results"], rec["spc"].core.downscale_factor)


@pytest.mark.parametrize("seed,kw", [(701, {}), (702, {"stress": True}), (703, {"short_frac": 0.3})])
def test_llr2_overwrite_i16_matches_oracle(seed, kw):
    """int16 ingest with mvs_detect_overwrite: every read goes through the general validate kernel"""
    from adapted_b200.detect import detect_reads

    spc = _overwrite_config("rna002")
    b = make_reads(192, "rna002", spc.sig_preload_size, seed=seed, **kw)
    x = b.to_dense_pa()
    try:
        want = detect_ref.detect_llr2(x, b.full_lens, spc)
    except ValueError:
        pytest.skip("the reference loses this minibatch")
    got, status = detect_reads(b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, spc, minibatch_size=192)
    assert not status.any()
    assert diff_results(got, want) == []
    assert any(r["mvs_adapter_end"] for r in want)


def test_overwrite_bounded_mean_range_f32_compare():
    """an explicit pA_mean_range is compared in float32 against the moving mean (numpy weak-scalar rule)"""
    from adapted_b200.detect import combined_detect_llr2

    d = config_as_dict(_overwrite_config("rna002"))
    d["mvs_polya"]["pA_mean_range"] = (104.3, 111.7)
    d["mvs_polya"]["pA_var_range"] = (0.1, 17.3)
    spc = config_from_dict(d)
    b = make_reads(128, "rna002", spc.sig_preload_size, seed=704)
    x = b.to_dense_pa()
    want = detect_ref.detect_llr2(x, b.full_lens, spc)
    got = combined_detect_llr2(x, b.full_lens, spc)
    assert diff_results(got, want) == []


def _cells(text):
    rows = list(csv.reader(io.StringIO(text)))
    return rows[0], rows[1:]


@pytest.mark.parametrize("name,method", [("llr_rna002_basic", 0), ("llr_rna002_stress", 0),
                                         ("llr_rna002_overwrite_stress", 0), ("start_peak_rna004_basic", 2)])
def test_gpu_records_to_reference_csv(name, method):
    """GPU records -> native writer == the CSV files the executed reference wrote (configs[0]: CSV-for-CSV).
    Integer, boolean, text and array cells are identical; float cells may differ by the 1e-5 contract before
    rounding, i.e. by at most one unit of the third decimal."""
    from adapted_b200 import _lib
    from adapted_b200.config import flatten_config
    from adapted_b200.detect import _dense_batch, _run_flat
    from adapted_b200.output import format_detected_boundaries

    rec = load_case(name)
    x = rec["batch"].to_dense_pa()
    b, keep = _dense_batch(x, rec["batch"].full_lens)
    flat = flatten_config(rec["spc"])
    flat["primary_method"] = method
    recs, status, _ = _run_flat(b, flat, None, 0, keep)
    assert not status.any()
    ok = recs["success"] != 0
    log = rec["results"][0].get("llr_detect_log")
    for sel, with_reason, key in ((np.flatnonzero(ok), False, "csv_pass"), (np.flatnonzero(~ok), True, "csv_fail")):
        got = format_detected_boundaries(recs, rec["read_ids"], method, with_reason, sel, log).decode()
        if got == rec[key]:
            continue
        gh, grows = _cells(got)
        wh, wrows = _cells(rec[key])
        assert gh == wh and len(grows) == len(wrows)
        for gr, wr in zip(grows, wrows):
            for col, g, w in zip(gh, gr, wr):
                if g == w:
                    continue
                assert col in FLOAT_FIELDS, (col, g, w)
                assert abs(float(g) - float(w)) <= 0.0011 + 1e-5 * abs(float(w)), (col, g, w)


# ---- row f4: streaming poly(A) detector -----------------------------------------------------------------------------
from tests.golden_io import load_stream_cases  # noqa: E402

STREAM_CASES = load_stream_cases()


@pytest.mark.parametrize("case", STREAM_CASES, ids=[c["name"] for c in STREAM_CASES])
def test_stream_detector_golden(case):
    """GPU mean_var_shift_polyA_detect == executed reference (dense float32 and ragged int16 ingest), exact"""
    from adapted_b200.detect import mean_var_shift_polyA_detect_batch, mean_var_shift_polyA_detect_i16

    b = case["batch"]
    lens = np.minimum(b.full_lens, case["m"]).astype(np.int32)
    got = mean_var_shift_polyA_detect_batch(b.to_dense_pa(), lens, case["params"])
    assert np.array_equal(got, case["want"])
    got16 = mean_var_shift_polyA_detect_i16(b.adc, b.offsets, b.calib_offset, b.calib_scale, case["params"], window=case["m"])
    assert np.array_equal(got16, case["want"])


def test_stream_detector_matches_oracle_and_single_call():
    from adapted_b200.config import StreamingConfig
    from adapted_b200.detect import mean_var_shift_polyA_detect, mean_var_shift_polyA_detect_batch

    p = StreamingConfig(min_obs_adapter=1000, search_increment_step=250, polyA_local_range=(0.0, 9.1),
                        polyA_med_range=(101.7, 113.9), median_shift_range=(19.3, None))
    b = make_reads(96, "rna004", 17500, seed=801, short_frac=0.2)
    x = b.to_dense_pa()
    lens = np.minimum(b.full_lens, 17500).astype(np.int32)
    want = np.array([detect_ref.mvs_stream_detect(x[i, : lens[i]], p) for i in range(b.n)])
    got = mean_var_shift_polyA_detect_batch(x, lens, p)
    assert np.array_equal(got, want)
    assert (want > 0).any() and (want == 0).any()
    i = int(np.flatnonzero(want > 0)[0])
    assert mean_var_shift_polyA_detect(x[i, : lens[i]], p) == want[i]
    assert mean_var_shift_polyA_detect(x[i, :200], p) == 0  # shorter than min_obs_adapter + windows


# ---- file-level job: container -> GPU -> boundary tables (configs[0]: CSV-for-CSV) ---------------------------------
def _load_job():
    import gzip
    import hashlib
    import json
    import os

    from tests.golden_io import GOLDEN

    with gzip.open(os.path.join(GOLDEN, "job_llr_rna002.json.gz"), "rb") as f:
        doc = json.loads(f.read().decode())
    cfg = doc["config"]
    for sec in cfg.values():
        for k, v in sec.items():
            if isinstance(v, list):
                sec[k] = tuple(v)
    spc = config_from_dict(cfg)
    b = make_reads(doc["job"]["n"], "rna002", doc["m"], seed=doc["job"]["seed"], stress=True)
    assert hashlib.sha256(b.adc.tobytes()).hexdigest() == doc["adc_sha256"]
    return doc, spc, b


def _assert_csv_close(got, want):
    if got == want:
        return
    gh, grows = _cells(got)
    wh, wrows = _cells(want)
    assert gh == wh and len(grows) == len(wrows)
    for gr, wr in zip(grows, wrows):
        for col, g, w in zip(gh, gr, wr):
            if g != w:
                assert col in FLOAT_FIELDS, (col, g, w)
                assert abs(float(g) - float(w)) <= 0.0011 + 1e-5 * abs(float(w)), (col, g, w)


def test_detect_file_writes_the_reference_tables(tmp_path):
    """three minibatches, 64 reads per table: same files, same rows, same text as the executed reference (floats
    within the 1e-5 contract before rounding); then `continue` on a half-finished output directory"""
    import os

    from adapted_b200.ingest import detect_file, processed_read_ids, write_container

    doc, spc, b = _load_job()
    job = doc["job"]
    path = write_container(str(tmp_path / "reads"), b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, doc["read_ids"])
    out = str(tmp_path / "out")
    stats = detect_file(path, out, spc, minibatch_size=job["minibatch"], batch_size_output=job["batch_size_output"])
    assert stats["reads"] == job["n"] and stats["lost"] == 0
    written = sorted(os.path.join(sub, f) for sub in ("boundaries", "failed_reads") for f in os.listdir(os.path.join(out, sub)))
    assert written == sorted(doc["files"])
    for rel, want in doc["files"].items():
        with open(os.path.join(out, rel), newline="") as f:
            _assert_csv_close(f.read(), want)
    assert processed_read_ids(out) == set(doc["read_ids"])
    # continue: keep the first pass table only, rerun -> the missing reads are processed again, numbering continues
    os.remove(os.path.join(out, "boundaries", "detected_boundaries_1.csv"))
    os.remove(os.path.join(out, "failed_reads", "failed_reads_0.csv"))
    stats2 = detect_file(path, out, spc, minibatch_size=job["minibatch"], batch_size_output=job["batch_size_output"],
                         continue_run=True)
    assert stats2["reads"] == job["n"] - job["batch_size_output"]
    assert os.path.exists(os.path.join(out, "boundaries", "detected_boundaries_1.csv"))
    assert processed_read_ids(out) == set(doc["read_ids"])


def test_detect_files_minibatches_run_across_files(tmp_path):
    """the job's reads split over two containers (70 + 80): minibatches of 50 cross the file boundary like the
    reference's producer fills them (file_proc.py:159-187) -> the very same tables"""
    import os

    from adapted_b200.ingest import detect_files, write_container

    doc, spc, b = _load_job()
    job, ids = doc["job"], doc["read_ids"]
    paths = []
    for k, (a, e) in enumerate(((0, 70), (70, job["n"]))):
        o = b.offsets
        paths.append(write_container(str(tmp_path / f"part{k}"), b.adc[o[a]: o[e]], o[a: e + 1] - o[a], b.full_lens[a:e],
                                     b.calib_offset[a:e], b.calib_scale[a:e], ids[a:e]))
    out = str(tmp_path / "out")
    stats = detect_files(paths, out, spc, minibatch_size=job["minibatch"], batch_size_output=job["batch_size_output"],
                         minibatches_per_call=2)
    assert stats["reads"] == job["n"] and stats["lost"] == 0
    for rel, want in doc["files"].items():
        with open(os.path.join(out, rel), newline="") as f:
            _assert_csv_close(f.read(), want)
