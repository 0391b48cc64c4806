// Warp-level restatement of the scipy.signal.find_peaks subset the reference uses:
// local maxima with plateau midpoints, distance suppression by descending height, prominence with bases,
// width at a relative height -- evaluated in that order (scipy/signal/_peak_finding.py:976-1008).
//
// The reference only ever consumes the FIRST one or two peaks that survive all filters
// (llr.py:197-200, combined.py:184, llr.py:452-477), so instead of materialising every property the warp
//   1. lists all local maxima,
//   2. classifies each with a budgeted two-sided walk (a finished side bounds the prominence from above, which
//      rejects noise wiggles after a few steps),
//   3. evaluates the survivors exactly (warp-cooperative walks) in ascending order until enough peaks passed,
//   4. for distance > 0 resolves "kept by _select_by_peak_distance" lazily, only for those survivors.
// Distance suppression depends on peak heights alone, prominence/width on the signal alone, so filtering in
// this order is equivalent to scipy's.
//
// All functions are called by ONE full warp (32 converged lanes); the trace lives in shared memory.
#pragma once
#include <float.h>

#include "adb_common.cuh"

struct TraceView {
    const double *x;  // shared memory
    int n;
    int nan2num;  // np.nan_to_num(trace, nan=0): nan -> 0, +-inf -> +-DBL_MAX (llr.py:445)
    __device__ __forceinline__ double at(int i) const {
        double v = x[i];
        if (nan2num) {
            if (v != v) v = 0.0;
            else if (v == CUDART_INF) v = DBL_MAX;
            else if (v == -CUDART_INF) v = -DBL_MAX;
        }
        return v;
    }
};

// ---- 1. local maxima (scipy _local_maxima_1d) ---------------------------------------------------------------
__device__ int warp_local_maxima(const TraceView &V, unsigned short *pk, int cap) {
    const int lane = threadIdx.x & 31;
    int count = 0;
    const int imax = V.n - 1;
    for (int base = 1; base < imax; base += 32) {
        int i = base + lane;
        int mid = -1;
        if (i < imax) {
            double xi = V.at(i);
            if (V.at(i - 1) < xi) {
                int ia = i + 1;
                while (ia < imax && V.at(ia) == xi) ia++;
                if (V.at(ia) < xi) mid = (i + ia - 1) / 2;
            }
        }
        unsigned m = __ballot_sync(ADB_FULL, mid >= 0);
        if (mid >= 0) {
            int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) pk[pos] = (unsigned short)mid;
        }
        count += __popc(m);
    }
    return min(count, cap);
}

// ---- 3a. exact prominence + bases of one peak, warp-cooperative (scipy _peak_prominences, wlen=-1) -----------
struct PeakProps {
    double prominence;
    int left_base, right_base;
};

__device__ PeakProps warp_prominence(const TraceView &V, int peak) {
    const int lane = threadIdx.x & 31;
    const double xp = V.at(peak);
    PeakProps P;
    // left walk: i = peak, peak-1, ... while i >= 0 and x[i] <= xp; strict '<' keeps the first minimum met.
    // Every lane keeps the minimum of the positions it visits (its own sub-sequence is in walk order, so a strict
    // '<' keeps the first one met); ONE warp reduction at the end of the walk merges the 32 candidates.
    double best = CUDART_INF;
    int bidx = -1;
    for (int cur = peak; cur >= 0; cur -= 32) {
        const int i = cur - lane;
        const double v = (i >= 0) ? V.at(i) : 0.0;
        const bool stop = (i < 0) || !(v <= xp);
        const unsigned sm = __ballot_sync(ADB_FULL, stop);
        const int nvalid = sm ? (__ffs(sm) - 1) : 32;
        if (lane < nvalid && v < best) { best = v; bidx = i; }
        if (sm) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(ADB_FULL, best, o);
        const int oi = __shfl_xor_sync(ADB_FULL, bidx, o);
        // smaller value wins; on ties the element met first in walk order (larger index on the left walk)
        if (ov < best || (ov == best && oi > bidx)) { best = ov; bidx = oi; }
    }
    double lmin = xp;
    int lbase = peak;
    if (best < lmin) { lmin = best; lbase = bidx; }
    best = CUDART_INF;
    bidx = 0x7fffffff;
    for (int cur = peak; cur < V.n; cur += 32) {
        const int i = cur + lane;
        const double v = (i < V.n) ? V.at(i) : 0.0;
        const bool stop = (i >= V.n) || !(v <= xp);
        const unsigned sm = __ballot_sync(ADB_FULL, stop);
        const int nvalid = sm ? (__ffs(sm) - 1) : 32;
        if (lane < nvalid && v < best) { best = v; bidx = i; }
        if (sm) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(ADB_FULL, best, o);
        const int oi = __shfl_xor_sync(ADB_FULL, bidx, o);
        if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    double rmin = xp;
    int rbase = peak;
    if (best < rmin) { rmin = best; rbase = bidx; }
    P.prominence = __dsub_rn(xp, fmax(lmin, rmin));
    P.left_base = lbase;
    P.right_base = rbase;
    return P;
}

// ---- 3b. width of one peak at height = x[peak] - prominence * rel_height (scipy _peak_widths) ----------------
__device__ double warp_width(const TraceView &V, int peak, const PeakProps &P, double rel_height) {
    const int lane = threadIdx.x & 31;
    const double height = __dsub_rn(V.at(peak), __dmul_rn(P.prominence, rel_height));
    // left: i = peak; while (i_min < i && height < x[i]) i--;
    int li = peak;
    for (int cur = peak;; cur -= 32) {
        int i = cur - lane;
        bool stop = !(i > P.left_base) || !(height < V.at(max(i, 0)));
        unsigned sm = __ballot_sync(ADB_FULL, stop);
        if (sm) { li = cur - (__ffs(sm) - 1); break; }
    }
    double left_ip = (double)li;
    {
        double xl = V.at(li);
        if (xl < height) left_ip = __dadd_rn(left_ip, __ddiv_rn(__dsub_rn(height, xl), __dsub_rn(V.at(li + 1), xl)));
    }
    int ri = peak;
    for (int cur = peak;; cur += 32) {
        int i = cur + lane;
        bool stop = !(i < P.right_base) || !(height < V.at(min(i, V.n - 1)));
        unsigned sm = __ballot_sync(ADB_FULL, stop);
        if (sm) { ri = cur + (__ffs(sm) - 1); break; }
    }
    double right_ip = (double)ri;
    {
        double xr = V.at(ri);
        if (xr < height) right_ip = __dsub_rn(right_ip, __ddiv_rn(__dsub_rn(height, xr), __dsub_rn(V.at(ri - 1), xr)));
    }
    return __dsub_rn(right_ip, left_ip);
}

// ---- 2. budgeted classification, one lane per peak -----------------------------------------------------------
// returns true if the peak is certainly rejected by `pmin <= prominence` or by `wmin <= width`:
//   * a side whose walk finished within the budget bounds the prominence from above (prom <= x[peak] - min of that
//     side); if that bound is already below pmin the peak is out (a NaN pmin rejects all);
//   * the width is measured at height x[peak] - prominence * rel_height, and a lower evaluation height can only widen
//     it (the bases only stop the walk earlier): with the prominence bound the lowest possible height is known, and
//     if the samples at or below it are found within the budget on both sides and are no more than wmin apart, the
//     true width is below wmin.
__device__ __forceinline__ bool lane_quick_reject(const TraceView &V, int peak, double pmin, double wmin,
                                                  double rel_height, int budget) {
    if (!(pmin == pmin)) return true;
    const double xp = V.at(peak);
    double prom_ub = CUDART_INF;
    double m = xp;
    int i = peak - 1, steps = 0;
    bool done = false;
    while (steps < budget) {
        if (i < 0) { done = true; break; }
        double v = V.at(i);
        if (!(v <= xp)) { done = true; break; }
        m = fmin(m, v);
        i--; steps++;
    }
    if (done) {
        prom_ub = __dsub_rn(xp, m);
        if (!(pmin <= prom_ub)) return true;
    }
    m = xp; i = peak + 1; steps = 0; done = false;
    while (steps < budget) {
        if (i >= V.n) { done = true; break; }
        double v = V.at(i);
        if (!(v <= xp)) { done = true; break; }
        m = fmin(m, v);
        i++; steps++;
    }
    if (done) {
        const double pr = __dsub_rn(xp, m);
        if (!(pmin <= pr)) return true;
        prom_ub = fmin(prom_ub, pr);
    }
    if (prom_ub < CUDART_INF && wmin > 0.0) {
        // lowest possible evaluation height; a NaN / inf height never rejects
        const double h = __dsub_rn(xp, __dmul_rn(prom_ub, rel_height));
        if (h == h && h > -CUDART_INF) {
            int il = -1, ir = -1;
            for (int k = 1; k <= budget; k++) {
                const int j = peak - k;
                if (j < 0) break;
                if (!(h < V.at(j))) { il = j; break; }
            }
            for (int k = 1; k <= budget; k++) {
                const int j = peak + k;
                if (j >= V.n) break;
                if (!(h < V.at(j))) { ir = j; break; }
            }
            // the walks of _peak_widths stop at il / ir at the latest: width <= ir - il (strictly less unless both
            // samples sit exactly at h, in which case it is equal) -- reject only with one sample to spare
            if (il >= 0 && ir >= 0 && (double)(ir - il) < wmin) return true;
        }
    }
    return false;
}

// ---- 4. lazy _select_by_peak_distance ---------------------------------------------------------------------------
// status: 0 unknown, 1 kept, 2 removed.  Priority = height, ties -> higher index first (np.argsort order read
// backwards; exact ties are a documented hazard of the reference).  Executed by lane 0.
__device__ bool lane_distance_kept(const TraceView &V, const unsigned short *pk, int npk, int j0, int dist,
                                   unsigned char *status, unsigned short *stack) {
    if (status[j0]) return status[j0] == 1;
    int sp = 0;
    stack[sp++] = (unsigned short)j0;
    while (sp > 0) {
        int t = stack[sp - 1];
        if (status[t]) { sp--; continue; }
        const double xt = V.at(pk[t]);
        int pending = -1;
        bool any_kept = false;
        for (int k = t - 1; k >= 0 && (int)pk[t] - (int)pk[k] < dist; k--) {
            double xk = V.at(pk[k]);
            if (xk > xt) {  // (ties: lower index has lower priority)
                if (status[k] == 0) { pending = k; break; }
                if (status[k] == 1) any_kept = true;
            }
        }
        if (pending < 0) {
            for (int k = t + 1; k < npk && (int)pk[k] - (int)pk[t] < dist; k++) {
                double xk = V.at(pk[k]);
                if (xk >= xt) {
                    if (status[k] == 0) { pending = k; break; }
                    if (status[k] == 1) any_kept = true;
                }
            }
        }
        if (pending >= 0) {
            stack[sp++] = (unsigned short)pending;
        } else {
            status[t] = any_kept ? 2 : 1;
            sp--;
        }
    }
    return status[j0] == 1;
}

// ---- driver: first `want` (<= 2) peaks of find_peaks(x, distance, prominence=pmin, width=wmin, rel_height) -------
// Scratch (shared): pk[cap] u16, stack[cap] u16, status[cap] u8, flags[cap] u8.  Returns the number found (uniform);
// out[] holds view-relative indices (uniform across the warp).
struct PeakScratch {
    unsigned short *pk, *stack;
    unsigned char *status, *flags;
    int cap;
};

// phase 3: the first `want` peaks among the prepared candidates (S.pk[0..npk), S.flags).  Warp-wide.
__device__ int warp_peaks_select(const TraceView &V, int npk, int dist, double pmin, double wmin, double rel_height,
                                 int want, int *out, const PeakScratch &S) {
    const int lane = threadIdx.x & 31;
    int found = 0;
    for (int base = 0; base < npk && found < want; base += 32) {
        int j = base + lane;
        unsigned todo = __ballot_sync(ADB_FULL, j < npk && S.flags[j]);
        while (todo && found < want) {
            int l = __ffs(todo) - 1;
            todo &= todo - 1;
            int jj = base + l;
            int peak = S.pk[jj];
            PeakProps P = warp_prominence(V, peak);
            if (!(pmin <= P.prominence)) continue;
            double w = warp_width(V, peak, P, rel_height);
            if (!(wmin <= w)) continue;
            int kept = 1;
            if (dist > 0) {
                if (lane == 0) kept = lane_distance_kept(V, S.pk, npk, jj, dist, S.status, S.stack) ? 1 : 0;
                kept = __shfl_sync(ADB_FULL, kept, 0);
            }
            if (kept) out[found++] = peak;
        }
    }
    return found;
}

__device__ int warp_find_first_peaks(const TraceView &V, int dist, double pmin, double wmin, double rel_height,
                                     int want, int *out, const PeakScratch &S) {
    const int lane = threadIdx.x & 31;
    if (V.n < 3) return 0;
    const int npk = warp_local_maxima(V, S.pk, S.cap);
    __syncwarp();
    for (int j = lane; j < npk; j += 32) {
        S.flags[j] = lane_quick_reject(V, S.pk[j], pmin, wmin, rel_height, 12) ? 0 : 1;
        S.status[j] = 0;
    }
    __syncwarp();
    return warp_peaks_select(V, npk, dist, pmin, wmin, rel_height, want, out, S);
}

// phases 1 + 2 with every warp of the CTA: local maxima chunk by chunk (32 positions per chunk, warps take chunks
// round robin; the at most 16 maxima of a chunk go to a fixed slot of `stack`, the counts to `status`), compaction in
// position order into S.pk, quick classification one thread per maximum.  `tmp` = one int of shared memory.
// CTA-wide; returns the number of maxima (uniform).  The view must not exceed (cap / 16 - 1) * 32 positions.
__device__ int cta_peaks_prepare(const TraceView &V, double pmin, double wmin, double rel_height, const PeakScratch &S,
                                 int *tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (V.n < 3) return 0;
    const int imax = V.n - 1;
    const int nch = (imax - 1 + 31) >> 5;  // positions 1 .. imax - 1
    for (int ch = warp; ch < nch; ch += nw) {
        const int i = 1 + (ch << 5) + lane;
        int mid = -1;
        if (i < imax) {
            const double xi = V.at(i);
            if (V.at(i - 1) < xi) {
                int ia = i + 1;
                while (ia < imax && V.at(ia) == xi) ia++;
                if (V.at(ia) < xi) mid = (i + ia - 1) / 2;
            }
        }
        const unsigned m = __ballot_sync(ADB_FULL, mid >= 0);
        if (mid >= 0) S.stack[(ch << 4) + __popc(m & ((1u << lane) - 1u))] = (unsigned short)mid;
        if (lane == 0) S.status[ch] = (unsigned char)__popc(m);
    }
    __syncthreads();
    if (warp == 0) {
        int total = 0;
        for (int base = 0; base < nch; base += 32) {
            const int ch = base + lane;
            const int cnt = (ch < nch) ? S.status[ch] : 0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(ADB_FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const int off = total + incl - cnt;
            for (int k = 0; k < cnt; k++) if (off + k < S.cap) S.pk[off + k] = S.stack[(ch << 4) + k];
            total += __shfl_sync(ADB_FULL, incl, 31);
        }
        if (lane == 0) *tmp = min(total, S.cap);
    }
    __syncthreads();
    const int npk = *tmp;
    for (int j = threadIdx.x; j < npk; j += blockDim.x) {
        S.flags[j] = lane_quick_reject(V, S.pk[j], pmin, wmin, rel_height, 12) ? 0 : 1;
        S.status[j] = 0;
    }
    __syncthreads();
    return npk;
}
