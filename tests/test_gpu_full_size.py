"""Full-size runs (BASELINE.json configs: 100 000 RNA002 reads / LLR path, RNA004 reads / CNN path) checked through
size-independent properties, plus an oracle spot check of whole minibatches:

  * the int16 fast paths (sampled one-pass global select, counting-based validate kernel) and the general kernels
    they stand in for give the same records on every read;
  * sharding independence: cutting the job at minibatch boundaries (what multi-GPU sharding and the pipelined ingest
    do) does not change a single byte of any record;
  * permutation invariance of the LLR path: reordering the reads INSIDE a minibatch permutes the records and changes
    nothing else (the minibatch-global median / MAD is a set statistic);
  * ground truth: the synthetic squiggles have known boundaries; the detector finds them.

ADB_FULL_SIZE_READS overrides the number of reads (default 100 000; the driver's GPU box has 180 GB)."""
import ctypes as C
import os

import numpy as np
import pytest

from adapted_b200.config import flatten_config, get_chemistry_specific_config
from tests.helpers import diff_results

pytestmark = pytest.mark.gpu

N_READS = int(os.environ.get("ADB_FULL_SIZE_READS", "100000"))
MBS = 1000

def _run(data, flat, weights=None, lo=0, hi=None, **opts):
    """adb_detect_dev on device-resident reads [lo, hi) (minibatch aligned); returns records (numpy), status."""
    import torch

    from adapted_b200 import _lib

    L = _lib.load()
    ctx = _lib.default_context(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        n_all = data["full_lens"].numel()
        hi = n_all if hi is None else hi
        n = hi - lo
        cfg = _lib.fill_config(flat)
        recs = torch.zeros(n * 512, dtype=torch.uint8, device="cuda")
        nb = (n + MBS - 1) // MBS
        status = torch.zeros(nb, dtype=torch.int32, device="cuda")
        offs = data["offsets"][lo:hi + 1].contiguous()
        wdev = None
        if weights is not None:
            wdev = torch.from_numpy(weights).cuda()
        keep = (data["full_lens"][lo:hi].contiguous(), data["calib_offset"][lo:hi].contiguous(), data["calib_scale"][lo:hi].contiguous())
        batch = _lib.AdbBatch(signal=data["adc"].data_ptr(), sig_type=_lib.SIG_I16, n_reads=n, m=int(flat["sig_preload_size"]),
                              batch_size=MBS, offsets=offs.data_ptr(), full_lens=keep[0].data_ptr(),
                              calib_offset=keep[1].data_ptr(), calib_scale=keep[2].data_ptr())
        torch.cuda.synchronize()
        _lib.check(L.adb_detect_dev(ctx.handle, C.byref(batch), C.byref(cfg), wdev.data_ptr() if wdev is not None else None,
                                    recs.data_ptr(), status.data_ptr(), None))
        torch.cuda.synchronize()
        fallbacks = ctx.query("global_select_fallbacks") if flat["primary_method"] == 0 and not opts.get("exact_global_select") else 0
        out = np.frombuffer(recs.cpu().numpy().tobytes(), dtype=_lib.RECORD_DTYPE).copy()
        return out, status.cpu().numpy(), fallbacks
    finally:
        for k in opts:
            ctx.set_option(k, 0)


def _assert_records_equivalent(a, b):
    """identical decisions, coordinates and order statistics; mean / std (summed differently) within 1e-6"""
    for name in a.dtype.names:
        if name in ("stats", "_reserved"):
            continue
        assert np.array_equal(a[name], b[name], equal_nan=True) if a[name].dtype.kind == "f" else np.array_equal(a[name], b[name]), name
    sa, sb = a["stats"], b["stats"]
    assert np.array_equal(sa[:, :, 2:], sb[:, :, 2:], equal_nan=True)  # med, mad
    assert np.allclose(sa[:, :, :2], sb[:, :, :2], rtol=1e-6, atol=0, equal_nan=True)  # mean, std


@pytest.fixture(scope="module")
def rna002():
    import torch

    from adapted_b200.synth import make_reads_torch

    spc = get_chemistry_specific_config("rna002")
    flat = flatten_config(spc)
    data = make_reads_torch(N_READS, "rna002", flat["sig_preload_size"], seed=77, device="cuda")
    torch.cuda.synchronize()
    base, st, fallbacks = _run(data, flat)
    return dict(spc=spc, flat=flat, data=data, base=base, status=st, fallbacks=fallbacks)


def test_llr_full_size_fast_paths_settle_everything(rna002):
    assert not rna002["status"].any()
    assert rna002["fallbacks"] == 0  # the sampled global select settled every one of the minibatches
    base = rna002["base"]
    assert 0.98 < base["success"].mean() <= 1.0
    truth = rna002["data"]["truth"].cpu().numpy()
    ok = base["success"] == 1
    ds = rna002["flat"]["downscale_factor"]
    assert np.median(np.abs(base["adapter_end"][ok] - truth[ok, 1])) <= ds
    assert np.median(np.abs(base["polya_end"][ok] - truth[ok, 2])) <= ds
    assert np.mean(np.abs(base["adapter_end"][ok] - truth[ok, 1]) <= 5 * ds) > 0.99


def test_llr_full_size_general_kernels_agree(rna002):
    gen, st, _ = _run(rna002["data"], rna002["flat"], exact_global_select=1, no_fast_validate=1)
    assert not st.any()
    _assert_records_equivalent(rna002["base"], gen)


def test_llr_full_size_sharding_independence(rna002):
    n = N_READS
    cut = (n // MBS) * 37 // 100 * MBS
    a, _, _ = _run(rna002["data"], rna002["flat"], lo=0, hi=cut)
    b, _, _ = _run(rna002["data"], rna002["flat"], lo=cut, hi=n)
    both = np.concatenate([a, b])
    assert both.tobytes() == rna002["base"].tobytes()


def test_llr_full_size_permutation_inside_minibatches(rna002):
    import torch

    data = rna002["data"]
    n = N_READS
    g = torch.Generator(device="cpu")
    g.manual_seed(5)
    perm = torch.cat([torch.randperm(min(MBS, n - s), generator=g) + s for s in range(0, n, MBS)])
    offs = data["offsets"].cpu()
    lens = (offs[1:] - offs[:-1])
    new_lens = lens[perm]
    new_offs = torch.zeros(n + 1, dtype=torch.int64)
    new_offs[1:] = torch.cumsum(new_lens, 0)
    pieces = [data["adc"][int(offs[i]):int(offs[i + 1])] for i in perm.tolist()]
    padc = torch.cat(pieces)
    del pieces
    dperm = perm.cuda()
    pdata = dict(adc=padc, offsets=new_offs.cuda(), full_lens=data["full_lens"][dperm].contiguous(),
                 calib_offset=data["calib_offset"][dperm].contiguous(), calib_scale=data["calib_scale"][dperm].contiguous())
    got, st, fb = _run(pdata, rna002["flat"])
    assert not st.any()
    want = rna002["base"][perm.numpy()]
    assert got.tobytes() == want.tobytes()


def test_llr_full_size_oracle_spot_check(rna002):
    """two whole minibatches (global statistics need all their reads) against the CPU oracle"""
    from adapted_b200.records import records_to_results
    from adapted_b200.synth import calibrate
    from oracle import detect_ref

    data, flat, spc = rna002["data"], rna002["flat"], rna002["spc"]
    m = flat["sig_preload_size"]
    offs = data["offsets"].cpu().numpy()
    nb = (N_READS + MBS - 1) // MBS
    for mb in sorted({0, nb // 2}):
        r0, r1 = mb * MBS, min((mb + 1) * MBS, N_READS)
        adc = data["adc"][int(offs[r0]):int(offs[r1])].cpu().numpy()
        coff = data["calib_offset"][r0:r1].cpu().numpy()
        cs = data["calib_scale"][r0:r1].cpu().numpy()
        lens = data["full_lens"][r0:r1].cpu().numpy()
        x = np.full((r1 - r0, m), np.nan, np.float32)
        for i in range(r1 - r0):
            a = adc[offs[r0 + i] - offs[r0]: offs[r0 + i + 1] - offs[r0]]
            x[i, :a.size] = calibrate(a, coff[i], cs[i])
        want = detect_ref.detect_llr2(x, lens, spc)
        got = records_to_results(rna002["base"][r0:r1], 0, "")
        assert diff_results(got, want) == []


@pytest.fixture(scope="module")
def rna004():
    import torch

    from adapted_b200.detect import flatten_cnn_weights
    from adapted_b200.synth import make_reads_torch
    from tests.golden_io import load_cnn_weights

    spc = get_chemistry_specific_config("rna004")
    flat = flatten_config(spc)
    n = min(N_READS, int(os.environ.get("ADB_FULL_SIZE_CNN_READS", "50000")))
    data = make_reads_torch(n, "rna004", flat["sig_preload_size"], seed=78, device="cuda")
    torch.cuda.synchronize()
    w = flatten_cnn_weights(load_cnn_weights())
    base, st, _ = _run(data, flat, weights=w)
    return dict(spc=spc, flat=flat, data=data, base=base, status=st, w=w, n=n)


def test_cnn_full_size_general_kernel_agrees_and_shards(rna004):
    assert not rna004["status"].any()
    base, n = rna004["base"], rna004["n"]
    assert 0.85 < base["success"].mean() < 1.0
    gen, st, _ = _run(rna004["data"], rna004["flat"], weights=rna004["w"], no_fast_validate=1)
    assert not st.any()
    _assert_records_equivalent(base, gen)
    cut = (n // MBS) // 2 * MBS
    a, _, _ = _run(rna004["data"], rna004["flat"], weights=rna004["w"], lo=0, hi=cut)
    b, _, _ = _run(rna004["data"], rna004["flat"], weights=rna004["w"], lo=cut, hi=n)
    assert np.concatenate([a, b]).tobytes() == base.tobytes()
    truth = rna004["data"]["truth"].cpu().numpy()
    ok = base["success"] == 1
    ds = rna004["flat"]["downscale_factor"]
    assert np.median(np.abs(base["adapter_end"][ok] - truth[ok, 1])) <= 2 * ds


# ---- BASELINE config 4: long poly(A) / truncated-preload stress set ---------------------------------------------------
N_STRESS = min(N_READS, int(os.environ.get("ADB_FULL_SIZE_STRESS_READS", "40000")))


def _dense_rows(data, m, r0, r1):
    from adapted_b200.synth import calibrate

    offs = data["offsets"].cpu().numpy()
    adc = data["adc"][int(offs[r0]):int(offs[r1])].cpu().numpy()
    coff = data["calib_offset"][r0:r1].cpu().numpy()
    cs = data["calib_scale"][r0:r1].cpu().numpy()
    x = np.full((r1 - r0, m), np.nan, np.float32)
    for i in range(r1 - r0):
        a = adc[offs[r0 + i] - offs[r0]: offs[r0 + i + 1] - offs[r0]]
        x[i, :a.size] = calibrate(a, coff[i], cs[i])
    return x, data["full_lens"][r0:r1].cpu().numpy()


@pytest.mark.parametrize("chem", ["rna002", "rna004"])
def test_stress_set_full_size(chem):
    """poly(A) lengths up to the preload limit and 10 % reads ending early: poly(A) running past the window, "not
    enough signal" after the adapter, failed first candidates (CNN hand-over incl. the second moving-statistics pass
    and the hail-mary fallback).  Fast paths == general kernels on every read, records independent of the cut, one
    minibatch against the oracle."""
    import torch

    from adapted_b200.detect import flatten_cnn_weights
    from adapted_b200.records import records_to_results
    from adapted_b200.synth import make_reads_torch
    from oracle import detect_ref
    from tests.golden_io import load_cnn_weights
    from tests.test_gpu_cnn_path import _cnn_compare

    spc = get_chemistry_specific_config(chem)
    flat = flatten_config(spc)
    n = N_STRESS
    # LLR path: a read without a single downscaled sample loses its whole minibatch in the reference (SURVEY A.11,
    # covered by the golden "lost minibatch" case); here the short reads keep at least a few trace samples
    data = make_reads_torch(n, chem, flat["sig_preload_size"], seed=91, device="cuda", stress=True, short_frac=0.1,
                            short_min=50 if flat["primary_method"] == 1 else flat["min_obs_adapter"] + 200)
    torch.cuda.synchronize()
    cnn = flat["primary_method"] == 1
    w = flatten_cnn_weights(load_cnn_weights()) if cnn else None
    base, st, _ = _run(data, flat, weights=w)
    assert not st.any()
    fails = base[base["success"] == 0]
    assert 0.02 < len(fails) / n < 0.9
    assert len(set(fails["fail_code"].tolist())) >= 3  # several different reasons are exercised
    gen, st2, _ = _run(data, flat, weights=w, no_fast_validate=1, **({} if cnn else {"exact_global_select": 1}))
    assert not st2.any()
    _assert_records_equivalent(base, gen)
    cut = (n // MBS) * 3 // 10 * MBS
    a, _, _ = _run(data, flat, weights=w, lo=0, hi=cut)
    b, _, _ = _run(data, flat, weights=w, lo=cut, hi=n)
    assert np.concatenate([a, b]).tobytes() == base.tobytes()
    r0 = (n // MBS // 2) * MBS
    x, lens = _dense_rows(data, flat["sig_preload_size"], r0, r0 + MBS)
    got = records_to_results(base[r0:r0 + MBS], flat["primary_method"], None if cnn else "")
    if cnn:
        want = detect_ref.detect_cnn(x, lens, load_cnn_weights(), spc)
        _cnn_compare(got, want, flat["downscale_factor"])
    else:
        want = detect_ref.detect_llr2(x, lens, spc)
        assert diff_results(got, want) == []
