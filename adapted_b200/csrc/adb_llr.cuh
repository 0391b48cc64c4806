// LLR changepoint trace and boundary picking -- device side.
//
// Reference: adapted/detect/_c_llr.pyx:22-236 (var_c, _gains*, c_llr_trace[_gains]),
//            adapted/detect/llr.py:52-259,406-479 (LLRTrace, peak picking, corrections, poly(A) spike rule),
//            adapted/detect/combined.py:145-211 (per-read loop of combined_detect_llr2).
// Arithmetic contract: SURVEY.md A.2 -- sequential float64 prefix sums, individually rounded IEEE double
// operations (no FMA), gains written only where the reference's loop visits.
#pragma once
#include "adb_common.cuh"
#include "adb_peaks.cuh"

// _c_llr.pyx:22-37.  The start == 0 branch of the reference drops the subtrahends; subtracting an exact 0.0 gives the
// same doubles (x - 0.0 == x, also for -0.0 and non-finite x), so one straight-line body serves both cases.
__device__ __forceinline__ double var_c(int start, int end, const double *c, const double *c2) {
    if (start == end) return 0.0;
    const double s1 = start ? c[start - 1] : 0.0, s2 = start ? c2[start - 1] : 0.0;
    const double n = (double)(end - start);
    const double m = __ddiv_rn(__dsub_rn(c[end - 1], s1), n);
    return __dsub_rn(__ddiv_rn(__dsub_rn(c2[end - 1], s2), n), __dmul_rn(m, m));
}

// _c_llr.pyx:82-86 for every i the reference loop visits; gains[] zero elsewhere (np.zeros_like, :80).  CTA-wide.
// The two segment terms of a point run through ONE copy of the variance + log code (a rolled two-trip loop): with both
// inlined side by side the body of the point loop did not fit the instruction cache next to the scheduler (ncu:
// `no_instruction` was the largest stall of llr_primary_kernel, concentrated on these lines).
__device__ void cta_llr_gains(const double *c, const double *c2, int n, int start, int end, int head, int tail,
                              int stride, double *gains) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) gains[i] = 0.0;
    __syncthreads();
    const int i0 = start + head, i1 = end - tail;
    if (i0 < i1) {
        const double var_summed = __dmul_rn((double)(end - start), log(var_c(start, end, c, c2)));
        const int cnt = (i1 - i0 + stride - 1) / stride;
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const int i = i0 + k * stride;
            double h = 0.0, t = 0.0;
#pragma unroll 1
            for (int side = 0; side < 2; side++) {
                const int a = side ? i : start, b = side ? end : i;
                const double term = __dmul_rn((double)(b - a), log(var_c(a, b, c, c2)));
                if (side) t = term; else h = term;
            }
            gains[i] = __dsub_rn(var_summed, __dadd_rn(h, t));
        }
    }
    __syncthreads();
}

// mean(diff(gains[lo:hi:stride])) with python slice semantics and numpy's pairwise float64 mean; one thread.
__device__ double lane_mean_diff(const double *g, int n, int lo, int hi, int stride) {
    if (lo < 0) { lo += n; if (lo < 0) lo = 0; }
    if (hi > n) hi = n;
    int cnt = (hi > lo) ? (hi - lo + stride - 1) / stride : 0;
    int nd = cnt - 1;
    if (nd <= 0) return CUDART_NAN;
    double s = np_sum_f64([&](int k) { return __dsub_rn(g[lo + (k + 1) * stride], g[lo + k * stride]); }, nd);
    return __ddiv_rn(s, (double)nd);
}

// Early-stop variants (_c_llr.pyx:91-173).  The gains themselves do not depend on the stop rule (an early-stopped
// trace is a prefix of the full trace), so the full trace is computed in parallel, the stop position is found by
// evaluating the reference's predicates at the positions its loop would test them, and the tail is zeroed.
// mode 1: adapter early stop; mode 2: adapter + poly(A) early stop.  `tmp` = one int of shared memory.  CTA-wide.
__device__ void cta_llr_early_stop(double *gains, int n, int start, int end, int head, int tail, int stride,
                                   int mode, int a_window, int a_stride, int p_window, int p_stride, int *tmp) {
    const int i0 = start + head, i1 = end - tail;
    if (i0 >= i1) return;
    const int cnt = (i1 - i0 + stride - 1) / stride;
    if (threadIdx.x == 0) *tmp = 0x7fffffff;
    __syncthreads();
    // first loop index k (i = i0 + k*stride) whose adapter predicate fires
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        int i = i0 + k * stride;
        if (i >= i0 + a_window && ((i - i0) % a_stride) == 0) {
            if (lane_mean_diff(gains, n, i - a_window, i, stride) < 0) atomicMin(tmp, k);
        }
    }
    __syncthreads();
    int ka = *tmp;
    __syncthreads();
    int kstop = cnt;
    if (mode == 1) {
        kstop = min(ka, cnt);
    } else if (ka < cnt) {
        if (threadIdx.x == 0) *tmp = 0x7fffffff;
        __syncthreads();
        for (int k = ka + threadIdx.x; k < cnt; k += blockDim.x) {
            int i = i0 + k * stride;
            if (lane_mean_diff(gains, n, i - p_window, i, stride) > 0) atomicMin(tmp, k);
        }
        __syncthreads();
        kstop = min(*tmp, cnt);
        __syncthreads();
    }
    for (int k = kstop + threadIdx.x; k < cnt; k += blockDim.x) gains[i0 + k * stride] = 0.0;
    __syncthreads();
}

// LLRTrace._trace_start_end (llr.py:135-142): first / last index whose value is not <= 0 (NaN counts as positive);
// all <= 0 -> (0, n-1).  CTA-wide; `tmp` = 2 ints of shared memory.
__device__ void cta_trace_support(const double *g, int n, int &start, int &end, int *tmp) {
    if (threadIdx.x == 0) { tmp[0] = 0x7fffffff; tmp[1] = -1; }
    __syncthreads();
    int lo = 0x7fffffff, hi = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!(g[i] <= 0.0)) { lo = min(lo, i); hi = max(hi, i); }
    }
    if (hi >= 0) { atomicMin(&tmp[0], lo); atomicMax(&tmp[1], hi); }
    __syncthreads();
    start = (tmp[0] == 0x7fffffff) ? 0 : tmp[0];
    end = (tmp[1] < 0) ? n - 1 : tmp[1];
    __syncthreads();
}

// np.nanstd over g[lo:hi) (numpy/lib/_nanfunctions_impl.py _nanvar): NaN -> 0 in the sums, divided by the non-NaN
// count.  Warp-wide tree sums in float64: the result can differ from numpy's pairwise order in the last bits, which
// is far below the difference the two `log` implementations already put into every gain (it only scales the
// prominence threshold of adapter_end_from_trace, llr.py:221).  Returns the value to every lane.
__device__ double warp_nanstd(const double *g, int lo, int hi) {
    const int lane = threadIdx.x & 31;
    const int n = hi - lo;
    if (n <= 0) return CUDART_NAN;
    int cnt = 0;
    double s = 0.0;
    for (int i = lo + lane; i < hi; i += 32) { const double v = g[i]; if (v == v) { cnt++; s += v; } }
    cnt = warp_sum_i(cnt);
    s = warp_sum_d(s);
    if (cnt == 0) return CUDART_NAN;
    const double avg = s / (double)cnt;
    double ss = 0.0;
    for (int i = lo + lane; i < hi; i += 32) { const double v = g[i]; if (v == v) { const double d = v - avg; ss += d * d; } }
    ss = warp_sum_d(ss);
    return sqrt(ss / (double)cnt);
}

// correct_for_plateau (llr.py:145-177), warp-wide.  Returns the corrected absolute index.
__device__ int warp_plateau_fix(const double *g, int n, int peak) {
    const int lane = threadIdx.x & 31;
    const int L = min(peak + 500, n) - peak;  // len(trace_)
    const int nch = L - 1;                    // len(changes)
    const double thr = __dmul_rn(0.9, g[peak]);
    int best = -1;
    // candidates i = nch-10 .. 0 (descending); the first hit wins -> search from the top in chunks of 32
    for (int top = nch - 10; top >= 0 && best < 0; top -= 32) {
        int i = top - lane;
        bool ok = false;
        if (i >= 0) {
            ok = g[peak + i + 9] > thr;
            for (int k = 0; k < 9 && ok; k++) ok = (__dsub_rn(g[peak + i + k + 1], g[peak + i + k]) >= 0.0);
        }
        unsigned m = __ballot_sync(ADB_FULL, ok);
        if (m) best = top - (__ffs(m) - 1);
    }
    int plateau_end = (best >= 0) ? best + 9 : -1;
    return plateau_end > 0 ? peak + plateau_end : peak;
}

// adapter_end_from_trace (llr.py:204-259) -> cands[0] or -1 if there is no candidate: first peak of
// find_peaks(width, prominence, rel_height) on the support [start, end) of the trace, moved by correct_for_plateau
// (llr.py:145-177) and correct_for_split_peak (llr.py:180-201).  The candidate preparation of the two peak searches
// runs on every warp of the CTA, the selection walks on warp 0.  CTA-wide; `itmp` = 2 ints of shared memory.
__device__ int cta_adapter_end(const double *g, int n, int start, int end, double pmin, double wmin, double rel_height,
                               const PeakScratch &PS, int *itmp) {
    TraceView W;
    W.x = g + start;
    W.n = end - start;
    W.nan2num = 0;
    const int npk = cta_peaks_prepare(W, pmin, wmin, rel_height, PS, &itmp[0]);
    if (threadIdx.x < 32) {
        int out[2];
        int peak = -1;
        if (warp_peaks_select(W, npk, 0, pmin, wmin, rel_height, 1, out, PS) > 0) peak = warp_plateau_fix(g, n, out[0] + start);
        if (threadIdx.x == 0) itmp[1] = peak;
    }
    __syncthreads();
    int peak = itmp[1];
    if (peak < 0) return -1;
    // split-peak correction: first peak of find_peaks(trace[peak : peak + 500], width=10, prominence=1.0)
    TraceView W2;
    W2.x = g + peak;
    W2.n = min(peak + 500, n) - peak;
    W2.nan2num = 0;
    const int npk2 = cta_peaks_prepare(W2, 1.0, 10.0, 0.5, PS, &itmp[0]);
    if (threadIdx.x < 32) {
        int out[2];
        int res = peak;
        if (warp_peaks_select(W2, npk2, 0, 1.0, 10.0, 0.5, 1, out, PS) > 0 && g[out[0] + peak] >= __dmul_rn(0.9, g[peak]))
            res = out[0] + peak;
        if (threadIdx.x == 0) itmp[1] = res;
    }
    __syncthreads();
    peak = itmp[1];
    __syncthreads();
    return peak;
}

// the two-peak rule of detect_full_polya_trace_peak_with_spike (llr.py:452-477) given the first k (<= 2) peaks of the
// poly(A) trace g (raw values, not nan_to_num).  Warp-wide.  Returns the downscaled index (0 = none).
__device__ int warp_polya_spike(const double *g, int k, const int pk[2]) {
    const int lane = threadIdx.x & 31;
    if (k == 0) return 0;
    if (k == 1) return pk[0];
    const double h0 = g[pk[0]], h1 = g[pk[1]];  // raw trace, not nan_to_num (llr.py:459)
    if (h1 > h0) return pk[1];
    if (h1 < __dmul_rn(h0, 0.5)) return pk[0];
    // idx_min = argmin(trace[p0:p1]) (first minimum; numpy's argmin returns the first NaN if there is one)
    double bv = CUDART_INF;
    int bi = 0x7fffffff, nan_i = 0x7fffffff;
    for (int i = pk[0] + lane; i < pk[1]; i += 32) {
        double v = g[i];
        if (!(v == v)) nan_i = min(nan_i, i);
        else if (v < bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(ADB_FULL, bv, o);
        int oi = __shfl_xor_sync(ADB_FULL, bi, o);
        int on = __shfl_xor_sync(ADB_FULL, nan_i, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        nan_i = min(nan_i, on);
    }
    const int idx_min = (nan_i != 0x7fffffff) ? nan_i : bi;
    const int cnt = pk[1] - idx_min;  // >= 1; a single point gives ssxm == 0 -> r = 0 -> 0 (scipy returns nan, same decision)
    // r^2 of the regression of g[idx_min:p1] on its index (scipy.stats.linregress: r = ssxym / sqrt(ssxm*ssym))
    double sx = 0, sy = 0;
    for (int i = idx_min + lane; i < pk[1]; i += 32) { sx += (double)i; sy += g[i]; }
    sx = warp_sum_d(sx); sy = warp_sum_d(sy);
    const double mx = sx / cnt, my = sy / cnt;
    double sxx = 0, sxy = 0, syy = 0;
    for (int i = idx_min + lane; i < pk[1]; i += 32) {
        double dx = (double)i - mx, dy = g[i] - my;
        sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
    }
    sxx = warp_sum_d(sxx); sxy = warp_sum_d(sxy); syy = warp_sum_d(syy);
    double r;
    if (sxx == 0.0 || syy == 0.0) r = 0.0;
    else {
        r = sxy / sqrt(sxx * syy);
        if (r > 1.0) r = 1.0; else if (r < -1.0) r = -1.0;
    }
    return (r * r >= 0.99) ? pk[1] : 0;
}

// detect_full_polya_trace_peak_with_spike (llr.py:406-479) on the full-length trace g[0..n).  CTA-wide (candidate
// preparation on every warp, selection and the two-peak rule on warp 0); `itmp` = 2 ints of shared memory.  Returns
// the downscaled index (0 = none) to every thread.
__device__ int cta_polya_end(double *g, int n, const PeakScratch &PS, int *itmp) {
    const int lane = threadIdx.x & 31;
    // np.nan_to_num(trace, nan=0) (llr.py:445): non-finite gains are rare (a one-sample head segment has variance
    // "zero" up to rounding, so its log is -inf or NaN).  Up to four of them are replaced in place for the peak
    // search and restored afterwards, so that the search reads plain doubles; more than four keep the trace as it
    // is and convert on every access.  Warp 0 owns the side list.
    int nf_idx[4] = {-1, -1, -1, -1};
    double nf_val[4] = {0, 0, 0, 0};
    int nnf = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const double v = (i < n) ? g[i] : 0.0;
            unsigned m = __ballot_sync(ADB_FULL, !(fabs(v) <= DBL_MAX));
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const double bv = __shfl_sync(ADB_FULL, v, l);
#pragma unroll
                for (int q = 0; q < 4; q++) if (nnf == q) { nf_idx[q] = base + l; nf_val[q] = bv; }
                nnf++;
            }
        }
        if (nnf <= 4 && lane == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q < nnf) { const double v = nf_val[q]; g[nf_idx[q]] = (v != v) ? 0.0 : (v > 0 ? DBL_MAX : -DBL_MAX); }
        }
        if (threadIdx.x == 0) itmp[1] = (nnf <= 4) ? 1 : 0;
    }
    __syncthreads();
    const bool patched = itmp[1] != 0;
    TraceView W;
    W.x = g;
    W.n = n;
    W.nan2num = patched ? 0 : 1;
    const int npk = cta_peaks_prepare(W, 1.0, 10.0, 0.5, PS, &itmp[0]);
    if (threadIdx.x < 32) {
        int pk[2] = {0, 0};
        const int k = warp_peaks_select(W, npk, 10, 1.0, 10.0, 0.5, 2, pk, PS);
        if (patched && nnf > 0) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) if (q < nnf) g[nf_idx[q]] = nf_val[q];
            }
            __syncwarp();
        }
        const int res = warp_polya_spike(g, k, pk);
        if (lane == 0) itmp[1] = res;
    }
    __syncthreads();
    const int res = itmp[1];
    __syncthreads();
    return res;
}
