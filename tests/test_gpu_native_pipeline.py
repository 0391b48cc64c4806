"""the native overlapped file pipeline (adb_detect_files: reader thread -> pinned ring -> compressed H2D ->
device decode -> detection -> writer / formatter threads) against the synchronous python driver (ingest.detect_files),
whose tables are pinned to the executed reference's (tests/test_gpu_next_rows.py, job golden)."""
import os

import numpy as np
import pytest

from adapted_b200.config import get_chemistry_specific_config
from adapted_b200.synth import SCALE, make_reads
from tests.golden_io import load_cnn_weights

pytestmark = pytest.mark.gpu


def _tables(root):
    out = {}
    for sub in ("boundaries", "failed_reads"):
        d = os.path.join(root, sub)
        for fn in sorted(os.listdir(d)) if os.path.isdir(d) else []:
            with open(os.path.join(d, fn), "rb") as f:
                out[f"{sub}/{fn}"] = f.read()
    return out


def _write_job(tmp_path, chem, sizes, seed, form, short_frac=0.05):
    """the same reads as version-1 containers (python driver) and as ADBSIG02 containers (native pipeline)"""
    from adapted_b200.ingest import write_container, write_container_v2

    spc = get_chemistry_specific_config(chem)
    v1, v2 = [], []
    for k, n in enumerate(sizes):
        b = make_reads(n, chem, spc.sig_preload_size, seed=seed + k, short_frac=short_frac)
        ids = [f"{k:02d}-{i:06d}-read" for i in range(n)]
        if k == 0 and n > 40:  # one read with more open-pore runs than the record keeps
            o, start = b.offsets[17], b.truth[17, 0] + 150
            hi = np.int16(np.rint(230.0 / SCALE - b.calib_offset[17]))
            for r in range(70):
                b.adc[o + start + r * 14] = hi
        v1.append(write_container(str(tmp_path / f"v1_{k}"), b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, ids))
        v2.append(write_container_v2(str(tmp_path / f"v2_{k}"), b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, ids,
                                     compress=form != "raw", zstd=form == "zstd"))
    return spc, v1, v2


@pytest.mark.parametrize("chem,form", [("rna002", "svb16"), ("rna004", "svb16"), ("rna004", "zstd"), ("rna002", "raw")])
def test_native_pipeline_equals_python_driver(tmp_path, chem, form):
    """three files of 530 / 260 / 415 reads, minibatches of 100 running across the file boundaries, tables of 64 reads,
    chunks of 3 minibatches: every table byte-identical, incl. the read whose open-pore list overflows the record"""
    from adapted_b200.ingest import detect_files, detect_files_native

    spc, v1, v2 = _write_job(tmp_path, chem, (530, 260, 415), 300, form)
    w = load_cnn_weights() if chem == "rna004" else None
    a, b = str(tmp_path / "py"), str(tmp_path / "native")
    s1 = detect_files(v1, a, spc, model=w, minibatch_size=100, batch_size_output=64, minibatches_per_call=3)
    s2 = detect_files_native(v2, b, spc, model=w, minibatch_size=100, batch_size_output=64, chunk_minibatches=3)
    assert {k: s2[k] for k in ("reads", "pass", "fail", "lost", "files")} == {k: s1[k] for k in ("reads", "pass", "fail", "lost", "files")}
    ta, tb = _tables(a), _tables(b)
    assert sorted(ta) == sorted(tb) and len(ta) == s1["files"]
    for k in ta:
        assert ta[k] == tb[k], k
    assert s2["h2d_bytes"] < (1.3 if form == "raw" else 0.7) * sum(os.path.getsize(p) for p in v1)


def test_native_pipeline_selection_and_continue(tmp_path):
    """inclusion sets and `adapted continue` (file_proc.py:97-168): a first run over half of the reads, a continued run
    over everything -- together the tables of the python driver doing the same"""
    from adapted_b200.ingest import detect_files, detect_files_native, processed_read_ids

    spc, v1, v2 = _write_job(tmp_path, "rna002", (230, 170), 400, "svb16", short_frac=0.0)  # no lost minibatches
    all_ids = [f"{k:02d}-{i:06d}-read" for k, n in enumerate((230, 170)) for i in range(n)]
    first = set(all_ids[::2])
    a, b = str(tmp_path / "py"), str(tmp_path / "native")
    detect_files(v1, a, spc, read_ids_incl=first, minibatch_size=50, batch_size_output=40)
    detect_files_native(v2, b, spc, read_ids_incl=first, minibatch_size=50, batch_size_output=40, chunk_minibatches=2)
    assert _tables(a) == _tables(b)
    detect_files(v1, a, spc, minibatch_size=50, batch_size_output=40, continue_run=True)
    s = detect_files_native(v2, b, spc, minibatch_size=50, batch_size_output=40, chunk_minibatches=2, continue_run=True)
    assert s["reads"] == len(all_ids) - len(first)
    assert _tables(a) == _tables(b)
    assert processed_read_ids(b) >= set(all_ids)  # (the scan also picks up the wrapped lines of long array cells, like the reference's)


def test_native_pipeline_refuses_bad_containers(tmp_path):
    from adapted_b200 import _lib
    from adapted_b200.ingest import detect_files_native

    spc = get_chemistry_specific_config("rna002")
    bad = tmp_path / "bad.adbsig"
    bad.write_bytes(b"ADBSIG02" + b"\0" * 300)
    with pytest.raises(_lib.AdbError):
        detect_files_native([str(bad)], str(tmp_path / "o"), spc)
    with pytest.raises(_lib.AdbError):
        detect_files_native([str(tmp_path / "missing.adbsig")], str(tmp_path / "o"), spc)
