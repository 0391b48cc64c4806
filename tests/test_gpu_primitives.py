"""Kernel-level parity (GPU): LLR trace incl. early-stop branches, global med/MAD, downscale -- through the C ABI."""
import numpy as np
import pytest

from oracle import detect_ref

pytestmark = pytest.mark.gpu


def _squiggle(rng, n):
    k1, k2 = sorted(rng.integers(20, n - 20, size=2))
    x = np.concatenate([rng.normal(-1.0, 1.0, k1), rng.normal(1.5, 0.3, k2 - k1), rng.normal(0.5, 1.4, n - k2)])
    return x.astype(np.float32).astype(np.float64), int(k1)


def _ulp_diff(a, b):
    """max |a-b| in units of the local float64 spacing (NaN/inf must agree exactly)."""
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(a[~fin & ~np.isnan(a)], b[~fin & ~np.isnan(b)])
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.spacing(np.maximum(np.abs(a[fin]), np.abs(b[fin])))))


def test_llr_trace_matches_oracle():
    from adapted_b200.detect import c_llr_trace_batch

    rng = np.random.default_rng(5)
    sigs, params = [], []
    for i in range(24):
        n = int(rng.integers(60, 1700))
        x, k1 = _squiggle(rng, n)
        for p in ((0, n - 1, 5, 5, 1, 0, 0, 0, 0, 0, 0), (k1, n - 1, 1, 1, 1, 0, 0, 0, 0, 0, 0),
                  (0, n - 1, 5, 5, 2, 0, 0, 0, 0, 0, 0)):
            sigs.append(x)
            params.append(p)
    out = c_llr_trace_batch(sigs, params, return_c_c2=True)
    worst = 0.0
    for x, p, (g, c, c2) in zip(sigs, params, out):
        g_ref, c_ref, c2_ref = detect_ref.llr_trace(x, *p, 1)
        # prefix sums are sequential IEEE adds: bit-exact
        assert np.array_equal(c, c_ref) and np.array_equal(c2, c2_ref)
        # gains: identical operation order, only `log` differs (CUDA <= 1 ulp vs glibc): a few ulp of the
        # magnitude of the summands n*log(var) (the gain itself is a difference of those)
        assert np.array_equal(g == 0, g_ref == 0)
        scale = np.nanmax(np.abs(g_ref[np.isfinite(g_ref)])) + x.size * 10
        assert np.nanmax(np.abs(g - g_ref)) <= 64 * np.spacing(scale)
        worst = max(worst, float(np.nanmax(np.abs(g - g_ref)) / np.spacing(scale)))
    print("worst gain deviation (ulp of summand scale):", worst)


def test_llr_trace_golden_vectors():
    """against the gains the reference's own Cython kernel produced (tests/golden/c_llr_trace.npz)"""
    from adapted_b200.detect import c_llr_trace

    with np.load("tests/golden/c_llr_trace.npz") as z:
        for i in range(6):
            x = z[f"x{i}"]
            for tag in ("full", "aes", "pes", "tail"):
                a = [int(v) for v in z[f"a{i}_{tag}"]]
                g = c_llr_trace(x, *a, 0)
                want = z[f"g{i}_{tag}"]
                # early-stop position (integer decision) must be identical: same zero pattern
                assert np.array_equal(g == 0, want == 0), (i, tag)
                assert np.allclose(g, want, rtol=0, atol=1e-8, equal_nan=True)


def test_llr_early_stop_matches_oracle():
    from adapted_b200.detect import c_llr_trace

    rng = np.random.default_rng(17)
    for i in range(8):
        n = int(rng.integers(400, 1700))
        x, _ = _squiggle(rng, n)
        for stride in (1, 2, 5):
            for aes, pes, aw, as_, pw, ps in ((1, 0, 100, 20, 0, 0), (1, 0, 50, 10, 0, 0), (1, 1, 100, 20, 30, 10),
                                             (0, 1, 60, 10, 20, 10)):
                g = c_llr_trace(x, 0, n - 1, 5, 5, stride, aes, aw, as_, pes, pw, ps, 0)
                want = detect_ref.llr_trace(x, 0, n - 1, 5, 5, stride, aes, aw, as_, pes, pw, ps, 0)
                assert np.array_equal(g == 0, want == 0), (i, stride, aes, pes)
                assert np.allclose(g, want, rtol=0, atol=1e-8, equal_nan=True)


def test_llr_degenerate_variance():
    from adapted_b200.detect import c_llr_trace

    x = np.concatenate([np.zeros(30), np.ones(30), np.full(40, 2.0)])
    g = c_llr_trace(x, 0, x.size - 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0)
    want = detect_ref.llr_trace(x, 0, x.size - 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0)
    assert np.array_equal(np.isnan(g), np.isnan(want))
    assert np.array_equal(np.isinf(g), np.isinf(want))


@pytest.mark.parametrize("seed", [0, 1])
def test_global_med_mad_bit_exact(seed):
    from adapted_b200.detect import global_med_mad
    from adapted_b200.synth import make_reads

    b = make_reads(37, "rna002", 26500, seed=seed, short_frac=0.3)
    x = b.to_dense_pa()
    med, mad = global_med_mad(x, b.full_lens, 25000)
    sub = x[:, :25000]
    want_med = float(np.nanmedian(sub))
    want_mad = float(np.nanmedian(np.abs(sub - want_med)))
    assert (med, mad) == (want_med, want_mad)


def test_global_med_mad_even_odd_and_ties():
    from adapted_b200.detect import global_med_mad

    rng = np.random.default_rng(3)
    for n, m in ((1, 7), (2, 8), (3, 1001), (5, 64)):
        x = np.round(rng.normal(90, 12, size=(n, m))).astype(np.float32)  # heavy ties
        lens = rng.integers(1, m + 1, size=n).astype(np.int32)
        for i, l in enumerate(lens):
            x[i, l:] = np.nan
        med, mad = global_med_mad(x, lens, m)
        want_med = float(np.nanmedian(x))
        assert med == want_med
        assert mad == float(np.nanmedian(np.abs(x - want_med)))


@pytest.mark.parametrize("chem,n,mbs,seed", [("rna002", 400, 200, 11), ("rna004", 300, 300, 12), ("rna002", 64, 64, 13)])
def test_global_med_mad_sampled_one_pass_select(chem, n, mbs, seed):
    """int16 ingest: the sampled one-pass select must settle these minibatches itself (no hand-over to the exact
    multi-pass select) and return numpy's order statistics bit for bit; the multi-pass select agrees."""
    from adapted_b200.config import get_chemistry_specific_config
    from adapted_b200.detect import global_med_mad_i16
    from adapted_b200.synth import make_reads

    spc = get_chemistry_specific_config(chem)
    b = make_reads(n, chem, spc.sig_preload_size, seed=seed, short_frac=0.1)
    tm = spc.core.max_obs_trace
    got, fallbacks = global_med_mad_i16(b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, b.m, tm, mbs)
    exact, _ = global_med_mad_i16(b.adc, b.offsets, b.full_lens, b.calib_offset, b.calib_scale, b.m, tm, mbs, exact=True)
    assert fallbacks == 0
    x = b.to_dense_pa()
    for i in range(got.shape[0]):
        sub = x[i * mbs:(i + 1) * mbs, :tm]
        med = np.float32(np.nanmedian(sub))
        mad = np.float32(np.nanmedian(np.abs(sub - med)))
        assert (got[i, 0], got[i, 1]) == (med, mad), (i, got[i], med, mad)
        assert (exact[i, 0], exact[i, 1]) == (med, mad)


def test_global_med_mad_sampled_select_hands_over_when_it_cannot_decide():
    """heavy ties / tiny minibatches / degenerate calibration: the sampled select must hand over, results stay exact"""
    from adapted_b200.detect import global_med_mad_i16

    rng = np.random.default_rng(21)
    n, m = 40, 3000
    lens = rng.integers(500, m, size=n).astype(np.int32)
    offs = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    # two-level signal: the median sits on a huge tie, MAD has few distinct values
    adc = np.where(rng.random(offs[-1]) < 0.5, 500, 520).astype(np.int16) + rng.integers(0, 2, size=offs[-1]).astype(np.int16)
    coff = np.full(n, -10.0, np.float32)
    cscale = np.full(n, 0.1755, np.float32)
    got, fallbacks = global_med_mad_i16(adc, offs, lens, coff, cscale, m, m, 20)
    for i in range(2):
        vals = np.concatenate([(adc[offs[r]:offs[r + 1]].astype(np.float32) + coff[r]) * cscale[r] for r in range(i * 20, (i + 1) * 20)])
        med = np.float32(np.median(vals))
        mad = np.float32(np.median(np.abs(vals - med)))
        assert (got[i, 0], got[i, 1]) == (med, mad)


@pytest.mark.parametrize("factor,col0", [(10, 1000), (20, 2000), (10, 0), (7, 3)])
def test_downscale_bit_exact(factor, col0):
    from adapted_b200.detect import downscale_signal
    from adapted_b200.synth import make_reads

    b = make_reads(12, "rna004", 17500, seed=9, short_frac=0.4)
    x = b.to_dense_pa()
    got = downscale_signal(x, b.full_lens, factor, col0)
    want = detect_ref.mean_pool(x[:, col0:], factor)
    assert got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)


def _four_level(rng, n):
    k = sorted(rng.integers(5, n - 5, size=3))
    lv, sd = rng.normal(0, 1.5, size=4), rng.uniform(0.2, 1.5, size=4)
    parts = [k[0], k[1] - k[0], k[2] - k[1], n - k[2]]
    x = np.concatenate([rng.normal(lv[i], sd[i], m) for i, m in enumerate(parts)])
    return x.astype(np.float32).astype(np.float64)


def test_legacy_three_split_detectors_match_oracle():
    """c_llr_detect_adapter / c_llr_detect_adapter_polya (_c_llr.pyx:239-363, the arg-max-of-LLR detectors) on the
    GPU against the oracle (itself pinned bit for bit against the reference's compiled kernel)"""
    from adapted_b200.detect import (c_llr_detect_adapter, c_llr_detect_adapter_polya, c_llr_detect_adapter_polya_trace,
                                     c_llr_detect_batch)

    rng = np.random.default_rng(41)
    sigs = [_four_level(rng, int(rng.integers(40, 1700))) for _ in range(120)]
    sigs.append(_four_level(rng, 30))       # too short for any split
    sigs.append(np.round(_four_level(rng, 600)))  # heavy ties in the medians
    for moa, bt, mop in ((20, 5, 10), (8, 2, 4), (50, 1, 30)):
        got = c_llr_detect_batch(sigs, moa, bt, mop)
        got2 = c_llr_detect_batch(sigs, moa, bt, None)
        for x, g, g2 in zip(sigs, got, got2):
            want = detect_ref.llr_detect_adapter_polya(x, moa, bt, mop)
            have = (0, 0) if g[3] == 2 else (int(g[0]), int(g[1]), int(g[2]))
            assert have == want, (x.size, moa, bt, mop)
            assert (int(g2[0]), int(g2[1])) == detect_ref.llr_detect_adapter(x, moa, bt)
    x = sigs[3]
    assert c_llr_detect_adapter(x, 20, 5) == detect_ref.llr_detect_adapter(x, 20, 5)
    assert c_llr_detect_adapter_polya(x, 20, 5, 10) == detect_ref.llr_detect_adapter_polya(x, 20, 5, 10)
    for g, w in zip(c_llr_detect_adapter_polya_trace(x, 20, 5, 10), detect_ref.llr_boundary_traces(x, 20, 5, 10)):
        assert np.array_equal(g == 0, w == 0)
        assert np.allclose(g, w, rtol=0, atol=1e-8, equal_nan=True)
