"""the native table writer (adb_format_csv) against the CSV text the executed reference wrote
(save_detected_boundaries through pandas; captured in tests/golden by oracle/make_golden.py).  Runs without a GPU:
the golden results are packed into adb_record arrays (tests/helpers.results_to_records) and formatted natively."""
import os

import numpy as np
import pytest

from adapted_b200 import _lib
from adapted_b200.output import BoundaryTableWriter, format_detected_boundaries
from tests.golden_io import SEAM_CASES, load_case
from tests.helpers import results_to_records

METHOD = {"llr2": 0, "cnn": 1, "start_peak": 2}
CASES = [(s, n) for s, names in SEAM_CASES.items() for n in names]


@pytest.mark.parametrize("seam,name", CASES)
def test_native_csv_equals_reference_csv(seam, name):
    rec = load_case(name)
    if "raises" in rec:
        pytest.skip("minibatch lost by the reference: nothing is written")
    recs = results_to_records(rec["results"], METHOD[seam])
    ok = recs["success"] != 0
    log = rec["results"][0].get("llr_detect_log") if rec["results"] else None
    got_pass = format_detected_boundaries(recs, rec["read_ids"], METHOD[seam], False, np.flatnonzero(ok), log)
    got_fail = format_detected_boundaries(recs, rec["read_ids"], METHOD[seam], True, np.flatnonzero(~ok), log)
    assert got_pass.decode() == rec["csv_pass"]
    assert got_fail.decode() == rec["csv_fail"]


def test_int_columns_turn_float_when_a_read_died(tmp_path):
    """pandas infers float64 for an int column holding a None: one exception record changes every int cell."""
    rec = load_case("llr_rna002_stress")
    recs = results_to_records(rec["results"], 0)
    fail = np.flatnonzero(recs["success"] == 0)
    dead = recs[fail[:1]].copy()
    dead["valid"] = 0
    dead["fail_code"] = 21
    both = np.concatenate([recs[fail], dead])
    txt = format_detected_boundaries(both, [f"r{i}" for i in range(both.size)], 0, True).decode().splitlines()
    hdr = txt[0].split(",")
    first = txt[1].split(",")
    assert first[hdr.index("signal_len")].endswith(".0") and first[hdr.index("adapter_end")].endswith(".0")
    assert txt[-1].split(",")[-1] == "'NoneType' object is not iterable"
    assert txt[-1].split(",")[1:-1] == [""] * (len(hdr) - 2)


def test_table_writer_file_layout(tmp_path):
    """4000-reads-per-file batching, file names and index continuation of the saver threads (file_proc.py:312-457)."""
    rec = load_case("cnn_rna004_short")
    recs = results_to_records(rec["results"], 1)
    ids = rec["read_ids"]
    out = str(tmp_path)
    with BoundaryTableWriter(os.path.join(out, "boundaries"), os.path.join(out, "failed_reads"), 1,
                             batch_size_output=16) as w:
        for s in range(0, recs.size, 10):
            w.add(recs[s: s + 10], ids[s: s + 10])
    n_pass = int((recs["success"] != 0).sum())
    files = sorted(os.listdir(os.path.join(out, "boundaries")))
    assert files == [f"detected_boundaries_{i}.csv" for i in range((n_pass + 15) // 16)]
    rows = []
    for i in range(len(files)):
        with open(os.path.join(out, "boundaries", f"detected_boundaries_{i}.csv")) as f:
            lines = f.read().splitlines()
        assert len(lines) - 1 == (16 if i < len(files) - 1 or n_pass % 16 == 0 else n_pass % 16)
        rows += [l.split(",")[0] for l in lines[1:]]
    assert rows == [i for i, r in zip(ids, recs) if r["success"]]
    w2 = BoundaryTableWriter.continue_from(out, 1)
    assert w2.bidx["pass"] == len(files)
    assert w2.bidx["fail"] == len(os.listdir(os.path.join(out, "failed_reads")))


def test_empty_selection():
    assert format_detected_boundaries(np.zeros(0, _lib.RECORD_DTYPE), [], 0) == b"\n"


def _pandas_csv(results, read_ids, save_fail_reasons):
    """The reference's writer restated (adapted/output.py:26-51 + container_types.py:112-120): the checker."""
    import io

    import pandas as pd

    rows = []
    for r, i in zip(results, read_ids):
        d = r.to_dict()
        reason = d.pop("fail_reason", None)
        rows.append({"read_id": i, **d, "fail_reason": reason})
    df = pd.DataFrame(rows)
    if not df.empty:
        drop = ["success", "llr_trace"] + ([] if save_fail_reasons else ["fail_reason"])
        df = df.drop(columns=[c for c in drop if c in df.columns])
    buf = io.StringIO()
    df.round(3).to_csv(buf, index=False)
    return buf.getvalue()


@pytest.mark.parametrize("seed", range(6))
def test_native_csv_equals_pandas_on_random_records(seed):
    """Randomised records (rounding ties at the third decimal, float32 columns, long open-pore lists that numpy
    wraps, dead reads) through records_to_results -> pandas versus the native writer."""
    from adapted_b200.records import records_to_results

    rng = np.random.default_rng(seed)
    case = ["llr_rna002_stress", "cnn_rna004_short", "start_peak_rna004_basic", "cnn_rna004_overwrite_short"][seed % 4]
    seam = [s for s, names in SEAM_CASES.items() if case in names][0]
    base = results_to_records(load_case(case)["results"], METHOD[seam])
    recs = base[rng.integers(0, base.size, size=200)].copy()
    for name in ("stats", "mvs", "real", "med_shift"):
        v = recs[name]
        kind = rng.integers(0, 3, size=v.shape)
        ties = (rng.integers(-200000, 200000, size=v.shape) + 0.5) / 1000.0          # x.xxx5 in float64
        f32 = rng.normal(90, 40, size=v.shape).astype(np.float32).astype(np.float64)  # float32-born values
        recs[name] = np.where(kind == 0, ties, np.where(kind == 1, f32, np.round(f32, 1)))
    recs["real"][:, :2] = recs["real"][:, :2].astype(np.float32)
    recs["med_shift"] = recs["med_shift"].astype(np.float32)
    n_op = rng.integers(0, 21, size=recs.size)
    recs["n_open_pores"] = n_op
    recs["open_pores"] = rng.integers(0, 30000, size=recs["open_pores"].shape)
    if seed % 2:
        dead = rng.random(recs.size) < 0.05
        recs["valid"][dead] = 0
        recs["success"][dead] = 0
        recs["fail_code"][dead] = 21
    ids = [f"read-{i}" for i in range(recs.size)]
    res = records_to_results(recs, METHOD[seam], "" if seam == "llr2" else None)
    ok = recs["success"] != 0
    for mask, with_reason in ((ok, False), (~ok, True)):
        sel = np.flatnonzero(mask)
        got = format_detected_boundaries(recs, ids, METHOD[seam], with_reason, sel, "" if seam == "llr2" else None)
        want = _pandas_csv([res[i] for i in sel], [ids[i] for i in sel], with_reason)
        assert got.decode() == want


def test_container_roundtrip_and_processed_ids(tmp_path):
    """the native signal container and the `continue` scan of finished tables (file_proc.py:103-130), no GPU"""
    import gzip
    import json

    from adapted_b200.ingest import processed_read_ids, read_container, write_container
    from adapted_b200.synth import make_reads
    from tests.golden_io import GOLDEN

    b = make_reads(6, "rna004", 17500, seed=3)
    ids = [f"id-{i}" for i in range(6)]
    p = write_container(str(tmp_path / "c"), b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, ids)
    c = read_container(p)
    assert np.array_equal(c["adc"], b.adc) and np.array_equal(c["offsets"], b.offsets)
    assert np.array_equal(c["calib_scale"], b.calib_scale) and [x.decode() for x in c["read_ids"]] == ids
    with gzip.open(os.path.join(GOLDEN, "job_llr_rna002.json.gz"), "rb") as f:
        doc = json.loads(f.read().decode())
    for rel, text in doc["files"].items():
        os.makedirs(os.path.dirname(tmp_path / rel), exist_ok=True)
        with open(tmp_path / rel, "w", newline="") as f:
            f.write(text)
    assert processed_read_ids(str(tmp_path)) == set(doc["read_ids"])
    assert len(processed_read_ids(str(tmp_path), failed_only=True)) == 23


def test_yield_minibatches_selection_and_file_boundaries(tmp_path):
    """yield_signals_from_pod5 semantics (file_proc.py:143-190) on the int16 ingest: file order, minibatches running
    across files, truncation to the preload window, inclusion / exclusion sets"""
    from adapted_b200.ingest import write_container, yield_minibatches
    from adapted_b200.synth import make_reads

    files, all_ids, lens = [], [], []
    for f, n in enumerate((7, 5, 9)):
        b = make_reads(n, "rna004", 30000, seed=100 + f)
        ids = [f"f{f}-r{i}" for i in range(n)]
        files.append(write_container(str(tmp_path / f"c{f}"), b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, ids))
        all_ids += ids
        lens += b.full_lens.tolist()
    m = 17500
    got = list(yield_minibatches(files, None, None, 4, m))
    assert [len(g[5]) for g in got] == [4, 4, 4, 4, 4, 1]
    assert [i for g in got for i in g[5]] == all_ids
    for g in got:
        assert np.array_equal(np.diff(g[1]), np.minimum(g[2], m)) and g[0].size == g[1][-1]
    assert np.concatenate([g[2] for g in got]).tolist() == lens
    incl, excl = {all_ids[1], all_ids[8], all_ids[20], "missing"}, {all_ids[8]}
    assert [i for g in yield_minibatches(files, incl, None, 4, m) for i in g[5]] == [all_ids[1], all_ids[8], all_ids[20]]
    assert [i for g in yield_minibatches(files, incl, excl, 4, m) for i in g[5]] == [all_ids[1], all_ids[20]]
    assert [i for g in yield_minibatches(files, None, excl, 50, m) for i in g[5]] == [i for i in all_ids if i != all_ids[8]]


@pytest.mark.parametrize("seam,name", CASES)
def test_records_round_trip_to_results(seam, name):
    """adb_record -> DetectResults (adapted_b200.records.records_to_results, the host side of the seam) restores every
    field of what the executed reference returned, values and None pattern alike (CPU, no GPU needed)"""
    from adapted_b200.records import records_to_results
    from tests.helpers import diff_results

    rec = load_case(name)
    if "raises" in rec:
        pytest.skip("minibatch lost by the reference")
    recs = results_to_records(rec["results"], METHOD[seam])
    back = records_to_results(recs, METHOD[seam], rec["results"][0].get("llr_detect_log"))
    assert diff_results(back, rec["results"], exact_floats=True) == []
    for b, w in zip(back, rec["results"]):
        for k, v in w.items():
            if v is None:
                assert getattr(b, k) is None, k


def test_open_pore_lists_beyond_the_record_are_never_truncated():
    """A record keeps ADB_MAX_OPEN_PORES run starts and the true count.  Both host paths refuse a longer list without
    its overflow row (no silent truncation) and print the full array -- numpy's wrapped str(), like the reference
    (combined.py:412-419, output.py:26-51) -- when the row is supplied."""
    from adapted_b200.records import records_to_results

    rng = np.random.default_rng(5)
    base = results_to_records(load_case("llr_rna002_basic")["results"], 0)
    recs = base[base["success"] != 0][:6].copy()
    ids = [f"read-{i}" for i in range(recs.size)]
    full = {}
    for i, n in ((1, 49), (4, 137)):
        lst = np.sort(rng.choice(12000, size=n, replace=False)).astype(np.int32)
        recs["n_open_pores"][i] = n
        recs["open_pores"][i] = lst[: _lib.ADB_MAX_OPEN_PORES]
        full[i] = lst
    with pytest.raises(OverflowError):
        format_detected_boundaries(recs, ids, 0)
    with pytest.raises(OverflowError):
        records_to_results(recs, 0, "")
    with pytest.raises(OverflowError):  # a row of the wrong length is refused as well
        format_detected_boundaries(recs, ids, 0, open_pore_overflow={1: full[1], 4: full[4][:-1]})
    res = records_to_results(recs, 0, "", full)
    assert np.array_equal(res[4].open_pores, full[4]) and res[4].open_pores.dtype == np.int64
    got = format_detected_boundaries(recs, ids, 0, open_pore_overflow=full)
    assert got.decode() == _pandas_csv(res, ids, False)
    # a selection that leaves the long records out needs no rows
    sel = [0, 2, 3, 5]
    assert format_detected_boundaries(recs, ids, 0, sel=sel).decode() == _pandas_csv([res[i] for i in sel], [ids[i] for i in sel], False)


def test_table_writer_carries_overflow_rows(tmp_path):
    from adapted_b200.records import records_to_results

    base = results_to_records(load_case("llr_rna002_basic")["results"], 0)
    recs = base[base["success"] != 0][:10].copy()
    ids = [f"read-{i}" for i in range(recs.size)]
    lst = np.arange(100, 100 + 60 * 15, 15, dtype=np.int32)
    recs["n_open_pores"][7] = lst.size
    recs["open_pores"][7] = lst[: _lib.ADB_MAX_OPEN_PORES]
    out = str(tmp_path)
    with BoundaryTableWriter(os.path.join(out, "boundaries"), os.path.join(out, "failed_reads"), 0, batch_size_output=4,
                             llr_detect_log="") as w:
        w.add(recs[:5], ids[:5])
        w.add(recs[5:], ids[5:], {2: lst})  # index 2 of this minibatch = read 7
    res = records_to_results(recs, 0, "", {7: lst})
    text = "".join(open(os.path.join(out, "boundaries", f"detected_boundaries_{i}.csv"), newline="").read().split("\n", 1)[1]
                   for i in range(3))
    assert text == _pandas_csv(res, ids, False).split("\n", 1)[1]
