// validate_boundaries + per-segment statistics for int16 sources WITHOUT shared-memory atomics.
//
// Reference: adapted/detect/combined.py:358-631 (control flow as in adb_validate.cuh).
//
// The histogram-based statistics of adb_validate.cuh are bound by the shared-memory atomic unit (~2 cycles per
// sample and histogram, ncu: profiles/r1_validate_kernel_v3.txt).  Here every order statistic of a read is found by
// COUNTING: the staged window of non-negative int16 ADC codes is read as packed half-precision bit patterns (for
// 0 <= code < 0x7c00 the integer order and the fp16 order coincide), one HSET2.LE compares two samples with a probe
// and the 0xffff / 0 lane masks are summed with plain integer adds.  All statistics of a read advance together:
//   round A  bisection on the code value for every rank needed (medians, p15 / p85 of numpy's linear percentile),
//   round B  bisection for the k-th smallest |pA(code) - med| (MAD): deviations are monotone on either side of the
//            median, so "how many samples deviate at most t" is the count of one code interval whose ends are found
//            by evaluating the float32 deviations exactly; the candidates of both sides are halved together by
//            probing the middle candidate of the longer side.
// Sums for mean / std are exact integer sums of the codes.  The float32 moving-statistics series (bottleneck
// recurrences, mvs_series_kernel) are staged in shared memory and their medians found by bisection on the ordered
// float bits.  Anything outside this path (float32 sources, codes outside [0, 0x7c00), no precomputed series, further
// poly(A) candidates after a failed first one, the CNN path's hail-mary fallback) is left to validate_kernel, which
// skips the reads flagged done here.
#pragma once
#include <cuda_fp16.h>

#include "adb_common.cuh"
#include "adb_gsample.cuh"
#include "adb_validate.cuh"
#include "adb_series_median.cuh"

#define VF_THREADS 256
#define VF_MAX_TASKS 12
#define VF_KEY_LIMIT 0x7c00

struct VfTask {
    int a, b;      // sample range [a, b) of the window
    int kind;      // 0: rank (count of codes <= idx); 1: k-th smallest deviation from the median (MAD)
    int k;         // 0-based rank looked for
    int lo, hi;    // search interval over the index; hi = first index whose count exceeds k (or the sentinel)
    int cnt_hi;    // count at hi
    int pv;        // deviation search: first code whose value is >= med
    float med;
    int sent;      // sentinel: hi == sent means "no index of this task satisfies the predicate"
    int smin, smax;  // code bounds of the samples of the segment
    // deviation search: candidates still undecided on the right side (codes pv + i, i in [rlo, rhi)) and on the left
    // side (codes pv - 1 - i, i in [llo, lhi)); candidates below the lows fail the predicate, from the his on they pass
    int rlo, rhi, llo, lhi;
    int jL, jR;    // codes pv - jL .. pv + jR - 1 deviate at most t_probe
    float t_probe; // deviation probed in the running pass
    float t_best;  // smallest deviation found so far that passes, and the count of samples deviating at most that much
    int c_best;
    int pA, pB;    // probes of the running pass: count the codes in [pA, pB]
    int mid;
    int active;
    // scan geometry, fixed per task: full 16-byte vectors [v0, v1) of W16, `voff` = offset of v0 in the flattened
    // vector space of all tasks of the round, boundary elements [hb, he) and [tb, te) (W16 indices, < 8 each)
    int v0, v1, voff;
    int hb, he, tb, te;
};

struct VfScratch {  // shared memory
    VfTask task[VF_MAX_TASKS];
    unsigned cnt[VF_MAX_TASKS];
    int ntask;
    int vtotal;    // vectors of all tasks of the round
    ushort4 items[VF_THREADS / 32][VF_MAX_TASKS];  // per warp: (task, first vector, end vector, owns the boundary elements)
    int nitems[VF_THREADS / 32];
    unsigned tmin[VF_MAX_TASKS], tmax[VF_MAX_TASKS];  // per-task code bounds (vf_task_bounds)
    uint32_t cand[2][64];  // keys left inside the brackets of vf_series_medians
    int ncand[2];
    int itmp[16];
    unsigned wtot[8];
    long long ltmp[8];
    float ftmp[8];
    double dtmp[8];
    uint32_t utmp[8];
};

struct VfRead {      // per-read constants (registers)
    const uint16_t *W16;  // 16-byte aligned base of the staged window
    int s0;               // index of sample 0 in W16
    int n;                // samples in the window
    float coff, cscale;
    int kmin, kmax;       // code bounds of the window
};

__device__ __forceinline__ float vf_pa(const VfRead &R, int code) { return gsb_pa(code, R.coff, R.cscale); }

__device__ __forceinline__ unsigned vf_decode(unsigned s) {
    // s = sum of HSET2 lane masks (0xffff per true half): recover the number of true halves
    const unsigned nlo = (0u - s) & 0xffffu;
    const unsigned t = ((s + nlo) >> 16) & 0xffffu;
    const unsigned nhi = (nlo - t) & 0xffffu;
    return nlo + nhi;
}

__device__ __forceinline__ __half2 vf_h2(unsigned code) {
    const unsigned w = (code & 0xffffu) * 0x00010001u;
    return *reinterpret_cast<const __half2 *>(&w);
}

// scan geometry of the window range [a, b) (a < b)
__device__ __forceinline__ void vf_geometry(const VfRead &R, VfTask &t, int a, int b, int voff) {
    const int i0 = R.s0 + a, i1 = R.s0 + b;
    const int v0 = (i0 + 7) >> 3, v1 = max(i1 >> 3, v0);
    t.v0 = v0; t.v1 = v1; t.voff = voff;
    t.hb = i0; t.he = min(i1, v0 << 3);
    t.tb = max(v1 << 3, t.he); t.te = i1;
}

// per-lane partial of: number of samples in the full vectors [vb, ve) of W16 whose code lies in [pA, pB] (pA <= pB).
// Two vectors per iteration, both loads issued before the compares.
__device__ __forceinline__ int vf_count_vectors(const VfRead &R, int vb, int ve, int pA, int pB) {
    const int lane = threadIdx.x & 31;
    const uint4 *V = reinterpret_cast<const uint4 *>(R.W16);
    const __half2 hB = vf_h2((unsigned)pB);
    unsigned sB = 0, sA = 0;
    if (pA > 0) {
        const __half2 hA = vf_h2((unsigned)(pA - 1));
        auto eat = [&](const uint4 &q) {
            const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const __half2 h = *reinterpret_cast<const __half2 *>(&w[t]);
                sB += __hle2_mask(h, hB);
                sA += __hle2_mask(h, hA);
            }
        };
        int v = vb + lane;
        for (; v + 32 < ve; v += 64) {
            const uint4 qa = V[v], qb = V[v + 32];
            eat(qa);
            eat(qb);
        }
        if (v < ve) eat(V[v]);
        return (int)vf_decode(sB) - (int)vf_decode(sA);
    }
    auto eat1 = [&](const uint4 &q) {
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int t = 0; t < 4; t++) sB += __hle2_mask(*reinterpret_cast<const __half2 *>(&w[t]), hB);
    };
    int v = vb + lane;
    for (; v + 32 < ve; v += 64) {
        const uint4 qa = V[v], qb = V[v + 32];
        eat1(qa);
        eat1(qb);
    }
    if (v < ve) eat1(V[v]);
    return (int)vf_decode(sB);
}

// float32 deviation of a code from the median, as the reference computes it on the pA values
__device__ __forceinline__ float vf_dev(const VfRead &R, int code, float med) { return fabsf(__fsub_rn(vf_pa(R, code), med)); }

// deviation of the i-th candidate of a side (right: code pv + i, left: code pv - 1 - i); non-decreasing in i
__device__ __forceinline__ float vf_side_dev(const VfRead &R, const VfTask &t, bool right, int i) {
    return vf_dev(R, right ? t.pv + i : t.pv - 1 - i, t.med);
}

// first index of a side (n candidates) whose deviation is > t (strict) or >= t: the calibration is nearly linear, so
// t / scale is within a step or two of the answer; the guess is corrected by exact float32 evaluations
__device__ int vf_side_first(const VfRead &R, const VfTask &t, bool right, int n, float thr, bool strict) {
    float gf = thr / R.cscale;
    int g = (gf == gf && gf < 1e9f) ? (int)gf : n;
    g = min(max(g, 0), n);
    int guard = 0;
    while (g > 0 && guard++ < 100000) {
        const float d = vf_side_dev(R, t, right, g - 1);
        if (strict ? (d > thr) : (d >= thr)) g--; else break;
    }
    while (g < n && guard++ < 100000) {
        const float d = vf_side_dev(R, t, right, g);
        if (strict ? !(d > thr) : !(d >= thr)) g++; else break;
    }
    return g;
}

__device__ __forceinline__ bool vf_pending(const VfTask &t) {
    return t.kind == 0 ? (t.lo < t.hi) : (t.rlo < t.rhi || t.llo < t.lhi);
}

// choose the probe of the next pass (one thread); false when the task has converged
__device__ bool vf_prepare(const VfRead &R, VfTask &t) {
    if (!vf_pending(t)) { t.active = 0; return false; }
    t.active = 1;
    if (t.kind == 0) { t.mid = (t.lo + t.hi) >> 1; t.pA = 0; t.pB = t.mid; return true; }
    // the middle candidate of the longer side: by symmetry of the two sides it halves the other one as well
    const int nR = t.smax + 1 - t.pv, nL = t.pv - t.smin;
    const bool right = (t.rhi - t.rlo) >= (t.lhi - t.llo);
    const int i = right ? (t.rlo + t.rhi) >> 1 : (t.llo + t.lhi) >> 1;
    const float thr = vf_side_dev(R, t, right, i);
    t.t_probe = thr;
    t.jR = vf_side_first(R, t, true, nR, thr, true);
    t.jL = vf_side_first(R, t, false, nL, thr, true);
    t.pA = t.pv - t.jL;
    t.pB = t.pv + t.jR - 1;
    return true;
}

// take the count of the pass (one thread)
__device__ void vf_update(const VfRead &R, VfTask &t, int c) {
    if (t.kind == 0) {
        if (c > t.k) { t.hi = t.mid; t.cnt_hi = c; } else t.lo = t.mid + 1;
        return;
    }
    if (c > t.k) {
        // every candidate deviating at least t_probe passes
        const int nR = t.smax + 1 - t.pv, nL = t.pv - t.smin;
        t.rhi = min(t.rhi, vf_side_first(R, t, true, nR, t.t_probe, false));
        t.lhi = min(t.lhi, vf_side_first(R, t, false, nL, t.t_probe, false));
        t.rlo = min(t.rlo, t.rhi);
        t.llo = min(t.llo, t.lhi);
        t.t_best = t.t_probe;
        t.c_best = c;
    } else {
        // every candidate deviating at most t_probe fails
        t.rlo = min(max(t.rlo, t.jR), t.rhi);
        t.llo = min(max(t.llo, t.jL), t.lhi);
    }
}

__device__ void vf_build_items(VfScratch &S, bool all_tasks) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntask = S.ntask;
    const int nw = VF_THREADS / 32;
    if (lane == 0) {
        const int wbeg = (int)(((long long)S.vtotal * warp) / nw), wend = (int)(((long long)S.vtotal * (warp + 1)) / nw);
        int n = 0;
        for (int q = 0; q < ntask; q++) {
            const VfTask &t = S.task[q];
            if (!all_tasks && !vf_pending(t)) continue;
            const int fb = max(t.voff, wbeg), fe = min(t.voff + (t.v1 - t.v0), wend);
            const int edge = (warp == (q & (nw - 1))) && (t.hb < t.he || t.tb < t.te);
            if (fb < fe) S.items[warp][n++] = make_ushort4((unsigned short)q, (unsigned short)(t.v0 + (fb - t.voff)), (unsigned short)(t.v0 + (fe - t.voff)), (unsigned short)edge);
            else if (edge) S.items[warp][n++] = make_ushort4((unsigned short)q, 0, 0, 1);
        }
        S.nitems[warp] = n;
    }
    __syncwarp();
}

// code bounds of the samples of every (rank) task: the bisections then start from the segment's own range instead of
// the window's (an open-pore stretch or a spike elsewhere in the read would cost every task three more passes).
// CTA-wide; sets lo / hi / smin / smax of every task.
__device__ void vf_task_bounds(const VfRead &R, VfScratch &S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    const int ntask = S.ntask;
    if (tid < ntask) { S.tmin[tid] = 0xffffu; S.tmax[tid] = 0u; }
    vf_build_items(S, true);
    __syncthreads();
    const uint4 *V = reinterpret_cast<const uint4 *>(R.W16);
    const int nit = S.nitems[warp];
    for (int it = 0; it < nit; it++) {
        const ushort4 item = S.items[warp][it];
        const VfTask &t = S.task[item.x];
        unsigned mn = 0xffffffffu, mx = 0u;
        for (int v = item.y + lane; v < item.z; v += 32) {
            const uint4 q = V[v];
            mn = __vminu2(__vminu2(mn, q.x), __vminu2(q.y, __vminu2(q.z, q.w)));
            mx = __vmaxu2(__vmaxu2(mx, q.x), __vmaxu2(q.y, __vmaxu2(q.z, q.w)));
        }
        unsigned lo = min(mn & 0xffffu, mn >> 16), hi = max(mx & 0xffffu, mx >> 16);
        if (item.w && lane < 16) {
            const int i = (lane < 8) ? t.hb + lane : t.tb + (lane - 8);
            const int iend = (lane < 8) ? t.he : t.te;
            if (i < iend) { const unsigned code = R.W16[i]; lo = min(lo, code); hi = max(hi, code); }
        }
        lo = __reduce_min_sync(ADB_FULL, lo);
        hi = __reduce_max_sync(ADB_FULL, hi);
        if (lane == 0) { atomicMin(&S.tmin[item.x], lo); atomicMax(&S.tmax[item.x], hi); }
    }
    __syncthreads();
    if (tid < ntask) {
        VfTask &t = S.task[tid];
        t.smin = (int)S.tmin[tid]; t.smax = (int)S.tmax[tid];
        t.lo = t.smin; t.hi = t.smax; t.sent = t.smax + 1;
    }
}

__device__ void vf_run(const VfRead &R, VfScratch &S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    const int ntask = S.ntask;
    vf_build_items(S, false);
    bool mine = false;
    if (tid < ntask) {
        VfTask &t = S.task[tid];
        S.cnt[tid] = 0;
        mine = vf_prepare(R, t);
    }
    const int nit = S.nitems[warp];
    while (__syncthreads_or(mine)) {
        for (int it = 0; it < nit; it++) {
            const ushort4 item = S.items[warp][it];
            const VfTask &t = S.task[item.x];
            if (!t.active) continue;
            const int pA = t.pA, pB = t.pB;
            if (pB < pA) continue;
            int c = 0;
            if (item.y < item.z) c = vf_count_vectors(R, item.y, item.z, pA, pB);
            if (item.w && lane < 16) {
                const int i = (lane < 8) ? t.hb + lane : t.tb + (lane - 8);
                const int iend = (lane < 8) ? t.he : t.te;
                if (i < iend) { const int code = R.W16[i]; c += (code >= pA && code <= pB); }
            }
            c = __reduce_add_sync(ADB_FULL, c);
            if (lane == 0 && c) atomicAdd(&S.cnt[item.x], (unsigned)c);
        }
        __syncthreads();
        mine = false;
        if (tid < ntask) {
            VfTask &t = S.task[tid];
            if (t.active) {
                const int c = (int)S.cnt[tid];
                S.cnt[tid] = 0;
                vf_update(R, t, c);
                mine = vf_prepare(R, t);
            }
        }
    }
}

// smallest code > v (succ) / largest code < v (pred) among the samples [a, b); CTA-wide, rare path.  -1 if none.
__device__ int vf_neighbour(const VfRead &R, VfScratch &S, int a, int b, int v, bool succ) {
    __syncthreads();
    if (threadIdx.x == 0) S.itmp[15] = succ ? 0x7fffffff : -1;
    __syncthreads();
    int best = succ ? 0x7fffffff : -1;
    for (int j = a + threadIdx.x; j < b; j += blockDim.x) {
        const int c = R.W16[R.s0 + j];
        if (succ) { if (c > v) best = min(best, c); } else { if (c < v) best = max(best, c); }
    }
    if (succ) { if (best != 0x7fffffff) atomicMin(&S.itmp[15], best); } else { if (best >= 0) atomicMax(&S.itmp[15], best); }
    __syncthreads();
    const int r = S.itmp[15];
    __syncthreads();
    return (r == 0x7fffffff) ? -1 : r;
}

// ---- staging helpers ---------------------------------------------------------------------------------------------
// min / max code of the window (packed int16 min / max, VIMNMX.S16x2).  CTA-wide.
__device__ void vf_minmax(const VfRead &R, VfScratch &S, int &kmin, int &kmax) {
    const int tid = threadIdx.x;
    const int i0 = R.s0, i1 = R.s0 + R.n;
    const int v0 = (i0 + 7) >> 3, v1 = i1 >> 3;
    const int head_end = min(i1, v0 << 3), tail_beg = max(v1 << 3, head_end);
    int lo = 0x7fff, hi = -0x8000;
    if (tid < 8) { const int i = i0 + tid; if (i < head_end) { const int c = (int16_t)R.W16[i]; lo = min(lo, c); hi = max(hi, c); } }
    else if (tid < 16) { const int i = tail_beg + tid - 8; if (i < i1) { const int c = (int16_t)R.W16[i]; lo = min(lo, c); hi = max(hi, c); } }
    unsigned mn = 0x7fff7fffu, mx = 0x80008000u;
    const uint4 *V = reinterpret_cast<const uint4 *>(R.W16);
    for (int v = v0 + tid; v < v1; v += VF_THREADS) {
        const uint4 q = V[v];
        mn = __vmins2(__vmins2(mn, q.x), __vmins2(q.y, __vmins2(q.z, q.w)));
        mx = __vmaxs2(__vmaxs2(mx, q.x), __vmaxs2(q.y, __vmaxs2(q.z, q.w)));
    }
    lo = min(lo, min((int)(int16_t)(mn & 0xffffu), (int)(int16_t)(mn >> 16)));
    hi = max(hi, max((int)(int16_t)(mx & 0xffffu), (int)(int16_t)(mx >> 16)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(ADB_FULL, lo, o));
        hi = max(hi, __shfl_xor_sync(ADB_FULL, hi, o));
    }
    __syncthreads();
    if ((tid & 31) == 0) { S.itmp[tid >> 5] = lo; S.itmp[8 + (tid >> 5)] = hi; }
    __syncthreads();
    lo = 0x7fff; hi = -0x8000;
    for (int w = 0; w < VF_THREADS / 32; w++) { lo = min(lo, S.itmp[w]); hi = max(hi, S.itmp[8 + w]); }
    __syncthreads();
    kmin = lo; kmax = hi;
}

// find_open_pores (anomalies.py:15-35) on codes: a sample is an open-pore sample iff code >= c200.  Same structure
// as open_pores_scan in adb_validate.cuh.
__device__ int vf_open_pores(const VfRead &R, VfScratch &S, int a, int b, int c200, adb_record *rec, int *last) {
    clip_seg(a, b, R.n);
    const int n = b - a;
    const int T = blockDim.x, tid = threadIdx.x;
    const int chunk = (n + T - 1) / T;
    const int j0 = min(tid * chunk, n), j1 = min(j0 + chunk, n);
    const uint16_t *w = R.W16 + R.s0 + a;
    int *sh = S.itmp;  // [0]=hits [1]=first hit [2]=last hit [4]=last valid
    __syncthreads();
    if (tid == 0) { sh[0] = 0; sh[1] = 0x7fffffff; sh[2] = -1; sh[3] = 0; sh[4] = -1; }
    __syncthreads();
    int hits = 0, first = 0x7fffffff, lastp = -1, ncand = 0, lastc = -1;
    for (int j = j0; j < j1; j++) {
        if ((int)w[j] >= c200) {
            hits++;
            first = min(first, j);
            lastp = j;
            bool gap = true;
            for (int k = 1; k < 10 && gap; k++)
                if (j - k >= 0 && (int)w[j - k] >= c200) gap = false;
            if (gap) { ncand++; lastc = j; }
        }
    }
    if (hits) {
        atomicAdd(&sh[0], hits);
        atomicMin(&sh[1], first);
        atomicMax(&sh[2], lastp);
    }
    __syncthreads();
    const int tot_hits = sh[0], first_hit = sh[1], last_hit = sh[2];
    int result_n;
    if (tot_hits == 0) {
        result_n = 0;
    } else if (tot_hits == 1) {
        result_n = 1;
        if (tid == 0) rec->open_pores[0] = a + first_hit;
        *last = a + first_hit;
    } else {
        if (first_hit >= j0 && first_hit < j1) { ncand--; if (lastc == first_hit) lastc = -1; }
        int incl = ncand;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(ADB_FULL, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        __syncthreads();
        if ((tid & 31) == 31) S.wtot[tid >> 5] = (unsigned)incl;
        if (lastc >= 0) atomicMax(&sh[4], lastc);
        __syncthreads();
        int wbase = 0, total = 0;
        for (int q = 0; q < (int)((T + 31) >> 5); q++) {
            if (q < (tid >> 5)) wbase += (int)S.wtot[q];
            total += (int)S.wtot[q];
        }
        int pos = wbase + incl - ncand;
        if (total > 0) {
            if (ncand > 0 && pos < ADB_MAX_OPEN_PORES) {
                for (int j = j0; j < j1 && pos < ADB_MAX_OPEN_PORES; j++) {
                    if (j == first_hit) continue;
                    if ((int)w[j] >= c200) {
                        bool gap = true;
                        for (int k = 1; k < 10 && gap; k++)
                            if (j - k >= 0 && (int)w[j - k] >= c200) gap = false;
                        if (gap) rec->open_pores[pos++] = a + j;
                    }
                }
            }
            result_n = total;
            *last = a + sh[4];
        } else {
            result_n = 1;
            if (tid == 0) rec->open_pores[0] = a + last_hit;
            *last = a + last_hit;
        }
    }
    __syncthreads();
    return result_n;
}

// numpy float32 mean of the `n` samples starting at a0 and at a1 (real_range_check: first / last mean_window samples)
// in numpy's pairwise order.  For 128 < n <= 512 the (at most four) leaves of the pairwise tree are summed by
// different threads and combined in tree order; other sizes by one thread per mean.  CTA-wide.
template <class T>  // uint16_t: non-negative codes of the staged window; int16_t: any code (adb_vhist.cuh)
__device__ void vf_mean_pair_t(const T *w, float coff, float cscale, VfScratch &S, int a0, int a1, int n, float &m0, float &m1) {
    __syncthreads();
    const int tid = threadIdx.x;
    if (n > 128 && n <= 512) {
        // tree: n -> (h0, n - h0); each part p > 128 -> (p2, p - p2)
        if (tid < 8) {
            const int which = tid >> 2, leaf = tid & 3;
            const int base = which ? a1 : a0;
            int h0 = n / 2; h0 -= h0 % 8;
            const int part = leaf >> 1;              // 0: left half, 1: right half
            const int pbase = part ? h0 : 0, plen = part ? n - h0 : h0;
            int lbase, llen;
            if (plen > 128) {
                int p2 = plen / 2; p2 -= p2 % 8;
                lbase = pbase + ((leaf & 1) ? p2 : 0);
                llen = (leaf & 1) ? plen - p2 : p2;
            } else {
                lbase = pbase;
                llen = (leaf & 1) ? 0 : plen;       // single leaf: the odd slot is unused
            }
            float sum = 0.f;
            if (llen > 0) {
                const T *p = w + base + lbase;
                sum = np_sum_f32_leaf([&](int i) { return __fmul_rn(__fadd_rn((float)(int)p[i], coff), cscale); }, llen);
            }
            S.ftmp[tid] = sum;
        }
        __syncthreads();
        auto combine = [&](int which) {
            int h0 = n / 2; h0 -= h0 % 8;
            const float *f = S.ftmp + which * 4;
            const float left = (h0 > 128) ? __fadd_rn(f[0], f[1]) : f[0];
            const float right = (n - h0 > 128) ? __fadd_rn(f[2], f[3]) : f[2];
            return __fdiv_rn(__fadd_rn(left, right), (float)n);
        };
        m0 = combine(0);
        m1 = combine(1);
    } else {
        if (tid == 0 || tid == 32) {
            const T *p = w + (tid == 0 ? a0 : a1);
            const float s = np_sum_f32([&](int i) { return __fmul_rn(__fadd_rn((float)(int)p[i], coff), cscale); }, n);
            S.ftmp[tid == 0 ? 0 : 1] = __fdiv_rn(s, (float)n);
        }
        __syncthreads();
        m0 = S.ftmp[0];
        m1 = S.ftmp[1];
    }
    __syncthreads();
}

__device__ __forceinline__ void vf_mean_pair(const VfRead &R, VfScratch &S, int a0, int a1, int n, float &m0, float &m1) {
    vf_mean_pair_t<uint16_t>(R.W16 + R.s0, R.coff, R.cscale, S, a0, a1, n, m0, m1);
}

// exact integer sums of the codes of up to three segments [sa[i], sb[i]) (clipped; skipped unless on[i]) -> mean /
// population std of the pA values in float64 (signal_partitions.py:91-92; numpy sums float32 pairwise -- the
// contract for these statistics is 1e-5 relative).  CTA-wide.
__device__ void vf_mean_std3(const VfRead &R, VfScratch &S, const int sa[3], const int sb[3], const bool on[3],
                             double mean_out[3], double std_out[3]) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid < 6) S.ltmp[tid] = 0;
    __syncthreads();
    const uint16_t *w = R.W16 + R.s0;
    int na[3];
    for (int sgm = 0; sgm < 3; sgm++) {
        int a = sa[sgm], b = sb[sgm];
        clip_seg(a, b, R.n);
        na[sgm] = on[sgm] ? b - a : 0;
        if (na[sgm] <= 0) continue;
        // 16-byte vectors over the aligned body, the (< 8 each) head / tail elements by the first threads; codes are
        // non-negative and below 0x7c00: a thread's sum of codes fits 32 bits, the squares accumulate in 64
        unsigned s1u = 0;
        unsigned long long s2u = 0;
        {
            const int i0 = R.s0 + a, i1 = R.s0 + b;
            const int v0 = (i0 + 7) >> 3, v1 = max(i1 >> 3, v0);
            const int he = min(i1, v0 << 3), tb = max(v1 << 3, he);
            const uint4 *V = reinterpret_cast<const uint4 *>(R.W16);
            for (int v = v0 + tid; v < v1; v += VF_THREADS) {
                const uint4 q = V[v];
                const unsigned ww[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const unsigned lo = ww[t] & 0xffffu, hi = ww[t] >> 16;
                    s1u += lo + hi;
                    s2u += (unsigned long long)lo * lo;
                    s2u += (unsigned long long)hi * hi;
                }
            }
            if (tid < 16) {
                const int i = (tid < 8) ? i0 + tid : tb + (tid - 8);
                const int iend = (tid < 8) ? he : i1;
                if (i < iend) { const unsigned c = R.W16[i]; s1u += c; s2u += (unsigned long long)c * c; }
            }
        }
        long long s1 = (long long)s1u, s2 = (long long)s2u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(ADB_FULL, s1, o); s2 += __shfl_xor_sync(ADB_FULL, s2, o); }
        if ((tid & 31) == 0) {
            atomicAdd((unsigned long long *)&S.ltmp[2 * sgm], (unsigned long long)s1);
            atomicAdd((unsigned long long *)&S.ltmp[2 * sgm + 1], (unsigned long long)s2);
        }
    }
    __syncthreads();
    for (int sgm = 0; sgm < 3; sgm++) {
        const int n = na[sgm];
        if (n <= 0) { mean_out[sgm] = CUDART_NAN; std_out[sgm] = CUDART_NAN; continue; }
        const double mk = (double)S.ltmp[2 * sgm] / n;
        double vk = (double)S.ltmp[2 * sgm + 1] / n - mk * mk;
        if (vk < 0) vk = 0;
        mean_out[sgm] = (double)(float)((mk + (double)R.coff) * (double)R.cscale);
        std_out[sgm] = (double)(float)(sqrt(vk) * fabs((double)R.cscale));
    }
    __syncthreads();
}

// np.nanmedian of up to two NaN-free float32 series (global memory) staged as ordered keys in `buf` (shared memory):
// bisection on the key value, both series advancing in the same passes, until at most VF_NCAND keys are left inside
// the bracket; those are gathered and ranked directly (float keys are nearly all distinct, so isolating a single key
// by bisection would take ~22 passes; isolating 64 of ~3000 takes ~6).  CTA-wide.
#define VF_NCAND 64
// STAGED = false (series longer than `buf`: poly(A) segments beyond half the window, the long-poly(A) stress set): every
// pass reads the 16-byte aligned rows from the pools (L2 resident) and forms the keys on the fly -- same passes, same
// result.
template <bool STAGED>
__device__ void vf_series_medians(VfScratch &S, const float *g0, int n0, const float *g1, int n1, uint32_t *buf, float out[2]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *k0 = buf, *k1 = buf + ((n0 + 3) & ~3);
    __syncthreads();
    uint32_t mn[2] = {0xffffffffu, 0xffffffffu}, mx[2] = {0u, 0u};
    for (int j = tid; j < n0; j += VF_THREADS) { const uint32_t k = f32_key(g0[j]); if (STAGED) k0[j] = k; mn[0] = min(mn[0], k); mx[0] = max(mx[0], k); }
    for (int j = tid; j < n1; j += VF_THREADS) { const uint32_t k = f32_key(g1[j]); if (STAGED) k1[j] = k; mn[1] = min(mn[1], k); mx[1] = max(mx[1], k); }
    for (int q = 0; q < 2; q++) { mn[q] = warp_min_u(mn[q]); mx[q] = warp_max_u(mx[q]); }
    if (tid < 4) S.utmp[tid] = (tid < 2) ? 0xffffffffu : 0u;
    if (tid < 6) S.cnt[tid] = 0;
    if (tid < 2) S.ncand[tid] = 0;
    __syncthreads();
    if (lane == 0) {
        atomicMin(&S.utmp[0], mn[0]); atomicMin(&S.utmp[1], mn[1]);
        atomicMax(&S.utmp[2], mx[0]); atomicMax(&S.utmp[3], mx[1]);
    }
    __syncthreads();
    // search state in registers (identical in every thread): keys in [lo, hi] hold the ranks cnt_lo .. cnt_hi - 1
    uint32_t lo[2] = {S.utmp[0], S.utmp[1]}, hi[2] = {S.utmp[2], S.utmp[3]};
    const int nn[2] = {n0, n1};
    const uint32_t *kk[2] = {k0, k1};
    const float *gg[2] = {g0, g1};
    auto key_at = [&](int q, int j) -> uint32_t { return STAGED ? kk[q][j] : f32_key(gg[q][j]); };
    const unsigned rank[2] = {n0 > 0 ? (unsigned)(n0 - 1) / 2 : 0u, n1 > 0 ? (unsigned)(n1 - 1) / 2 : 0u};
    unsigned cnt_hi[2] = {(unsigned)n0, (unsigned)n1}, cnt_lo[2] = {0u, 0u};
    auto open = [&](int q) { return nn[q] > 0 && lo[q] < hi[q] && cnt_hi[q] - cnt_lo[q] > VF_NCAND; };
    int pass = 0;
    while (open(0) || open(1)) {
        uint32_t mid[2];
        for (int q = 0; q < 2; q++) {
            mid[q] = lo[q] + ((hi[q] - lo[q]) >> 1);
            if (!open(q)) continue;
            int c = 0;
            const int nv = nn[q] >> 2;
            if (STAGED) {
                const uint4 *V = reinterpret_cast<const uint4 *>(kk[q]);
                for (int v = tid; v < nv; v += VF_THREADS) {
                    const uint4 x = V[v];
                    c += (x.x <= mid[q]) + (x.y <= mid[q]) + (x.z <= mid[q]) + (x.w <= mid[q]);
                }
            } else {
                const float4 *V = reinterpret_cast<const float4 *>(gg[q]);
                for (int v = tid; v < nv; v += VF_THREADS) {
                    const float4 x = V[v];
                    c += (f32_key(x.x) <= mid[q]) + (f32_key(x.y) <= mid[q]) + (f32_key(x.z) <= mid[q]) + (f32_key(x.w) <= mid[q]);
                }
            }
            const int j = (nv << 2) + tid;
            if (j < nn[q]) c += (key_at(q, j) <= mid[q]);
            c = __reduce_add_sync(ADB_FULL, c);
            if (lane == 0 && c) atomicAdd(&S.cnt[q + 2 * (pass % 3)], (unsigned)c);
        }
        __syncthreads();
        for (int q = 0; q < 2; q++) {
            if (!open(q)) continue;
            const unsigned c = S.cnt[q + 2 * (pass % 3)];
            if (c > rank[q]) { hi[q] = mid[q]; cnt_hi[q] = c; } else { lo[q] = mid[q] + 1; cnt_lo[q] = c; }
        }
        // three rotating counter sets: the set of the previous pass has been read by everybody (this pass's barrier
        // lies in between) and is not used again before the pass after next
        if (tid < 2) S.cnt[tid + 2 * ((pass + 2) % 3)] = 0;
        pass++;
    }
    // gather the keys left in the brackets (cnt_hi - cnt_lo of them, or all equal keys when lo == hi)
    for (int q = 0; q < 2; q++) {
        if (nn[q] <= 0 || lo[q] == hi[q]) continue;
        for (int j = tid; j < nn[q]; j += VF_THREADS) {
            const uint32_t k = key_at(q, j);
            if (k >= lo[q] && k <= hi[q]) { const int p = atomicAdd(&S.ncand[q], 1); if (p < VF_NCAND) S.cand[q][p] = k; }
        }
    }
    __syncthreads();
    // warp q ranks the candidates of series q: the keys at ranks r and r + 1 inside the bracket
    if (warp < 2) {
        const int q = warp;
        if (nn[q] > 0) {
            uint32_t a = hi[q], b = 0xffffffffu;  // lower / upper middle key; b unknown yet
            bool have_b = false;
            if (lo[q] == hi[q]) {
                have_b = cnt_hi[q] > rank[q] + 1;
                b = hi[q];
            } else {
                const int m = min((int)(cnt_hi[q] - cnt_lo[q]), VF_NCAND);
                const int r = (int)(rank[q] - cnt_lo[q]);
                for (int i = lane; i < m; i += 32) {
                    const uint32_t ki = S.cand[q][i];
                    int pos = 0;
                    for (int j = 0; j < m; j++) { const uint32_t kj = S.cand[q][j]; pos += (kj < ki) || (kj == ki && j < i); }
                    if (pos == r) S.utmp[4 + 2 * q] = ki;
                    if (pos == r + 1) S.utmp[5 + 2 * q] = ki;
                }
                __syncwarp();
                a = S.utmp[4 + 2 * q];
                have_b = (r + 1 < m);
                if (have_b) b = S.utmp[5 + 2 * q];
            }
            if (lane == 0) { S.utmp[4 + 2 * q] = a; S.utmp[5 + 2 * q] = b; S.itmp[q] = have_b ? 1 : 0; }
        }
    }
    __syncthreads();
    for (int q = 0; q < 2; q++) {
        if (nn[q] <= 0) { out[q] = CUDART_NAN_F; continue; }
        const float a = key_f32(S.utmp[4 + 2 * q]);
        if (nn[q] & 1) { out[q] = a; continue; }
        uint32_t up = S.utmp[5 + 2 * q];
        if (!S.itmp[q]) {
            // the upper middle element lies beyond the bracket: smallest key above hi (rare)
            __syncthreads();
            if (tid == 0) S.utmp[3] = 0xffffffffu;
            __syncthreads();
            uint32_t best = 0xffffffffu;
            for (int j = tid; j < nn[q]; j += VF_THREADS) { const uint32_t k = key_at(q, j); if (k > hi[q]) best = min(best, k); }
            best = warp_min_u(best);
            if (lane == 0) atomicMin(&S.utmp[3], best);
            __syncthreads();
            up = S.utmp[3];
            __syncthreads();
        }
        out[q] = __fdiv_rn(__fadd_rn(a, key_f32(up)), 2.0f);
    }
    __syncthreads();
    if (tid < 6) S.cnt[tid] = 0;
    __syncthreads();
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
struct VfastArgs {
    BatchDev B;
    const int *given;           // [n_reads][given_stride] = adapter_end, polya_end(= topk[0]), topk[1..]
    int given_stride;
    int given_ntopk;            // entries per read when ntopk_per_read == nullptr (-1: None)
    const int *ntopk_per_read;
    int mode;                   // ADB_METHOD_*
    int win_bytes;              // capacity of the staged window (bytes)
    adb_record *out;
    const int *batch_status;
    const float *pre_var, *pre_mean;  // compact pools of precomputed moving statistics (mvs_series_kernel)
    const long long *pre_off;
    const int *pre_meta;
    const float *series_med;    // [n_reads][2] medians of the two series of the row (series_median_kernel)
    unsigned char *done;        // [n_reads], zeroed before the launch; 1 = record written by this kernel,
                                // 2 = record written with the FIRST poly(A) candidate's outcome, further candidates pending
    int cand_followup;          // 1: validate_cand_kernel runs behind this kernel (CNN path with several candidates)
};

__host__ __device__ inline size_t vfast_smem_bytes(int win_bytes) {
    return (((size_t)win_bytes + 48 + 15) & ~(size_t)15);  // dynamic part: the window; the scratch is static
}

// one rank task over the window range [a, b) (clipped); returns the task index or -1 if the range is empty
__device__ __forceinline__ int vf_add_rank(VfScratch &S, const VfRead &R, int &nt, int &rot, int a, int b, int k) {
    clip_seg(a, b, R.n);
    const int n = b - a;
    if (n <= 0) return -1;
    const int q = nt++;
    const int nvec = max(((R.s0 + b) >> 3) - ((R.s0 + a + 7) >> 3), 0);
    if (threadIdx.x == 0) {
        VfTask &t = S.task[q];
        t.a = a; t.b = b; t.kind = 0; t.k = k; t.lo = R.kmin; t.hi = R.kmax; t.cnt_hi = n; t.pv = 0; t.med = 0.f;
        t.active = 0; t.sent = R.kmax + 1; t.smin = R.kmin; t.smax = R.kmax;
        vf_geometry(R, t, a, b, rot);
    }
    rot += nvec;
    return q;
}

// value of the order statistics k and k + 1 (if two) of a finished rank task -> codes v0, v1.  CTA-wide (uniform).
__device__ __forceinline__ void vf_rank_pair(const VfRead &R, VfScratch &S, int q, bool two, int &v0, int &v1) {
    const VfTask &t = S.task[q];
    v0 = t.hi;
    v1 = v0;
    const int a = t.a, b = t.b, k = t.k, c = t.cnt_hi;
    if (two && !(c > k + 1)) v1 = vf_neighbour(R, S, a, b, v0, true);
}

// numpy median of the range of a finished rank task with k = (n - 1) / 2
__device__ float vf_median_of(const VfRead &R, VfScratch &S, int q) {
    if (q < 0) return CUDART_NAN_F;
    const int n = S.task[q].b - S.task[q].a;
    int v0, v1;
    vf_rank_pair(R, S, q, (n & 1) == 0, v0, v1);
    const float x0 = vf_pa(R, v0);
    if (n & 1) return x0;
    return __fdiv_rn(__fadd_rn(x0, vf_pa(R, v1)), 2.0f);
}

struct VfDevOut { float mad; };

// add the deviation-search task of a segment (median known, code bounds [smin, smax] of its samples); returns its
// index or -1
__device__ __forceinline__ int vf_add_mad(VfScratch &S, const VfRead &R, int &nt, int &rot, int a, int b, float med,
                                          int smin, int smax) {
    clip_seg(a, b, R.n);
    const int n = b - a;
    if (n <= 0 || !(med == med)) return -1;
    const int q = nt++;
    int ok = 1;
    int pv = gsb_code_at(med, false, R.coff, R.cscale, &ok);
    pv = min(max(pv, smin), smax + 1);
    const int nvec = max(((R.s0 + b) >> 3) - ((R.s0 + a + 7) >> 3), 0);
    if (threadIdx.x == 0) {
        VfTask &t = S.task[q];
        t.a = a; t.b = b; t.kind = 1; t.k = (n - 1) / 2; t.cnt_hi = -1; t.pv = pv; t.med = med;
        t.smin = smin; t.smax = smax;
        t.lo = 0; t.hi = 0; t.sent = 0;
        t.rlo = 0; t.rhi = smax + 1 - pv; t.llo = 0; t.lhi = pv - smin;
        t.t_best = CUDART_INF_F; t.c_best = -1;
        t.active = 0;
        vf_geometry(R, t, a, b, rot);
    }
    rot += nvec;
    return q;
}

// median of |x - med| from the finished deviation task q.  CTA-wide (uniform).
__device__ float vf_mad_of(const VfRead &R, VfScratch &S, int q) {
    if (q < 0) return CUDART_NAN_F;
    const VfTask t = S.task[q];
    const int n = t.b - t.a, k = t.k, pv = t.pv;
    const float med = t.med;
    const float d0 = t.t_best;  // the last probe that passed is the smallest passing candidate (see vf_update)
    if (n & 1) return d0;
    float d1 = d0;
    if (!(t.c_best > k + 1)) {
        // the next larger deviation: first occupied code beyond the interval [l, r] of the codes deviating <= d0
        int lo = R.kmin, hi = pv;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (vf_dev(R, m, med) <= d0) hi = m; else lo = m + 1; }
        const int l = lo;  // == pv if no left code deviates <= d0
        lo = pv; hi = R.kmax + 1;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (vf_dev(R, m, med) > d0) hi = m; else lo = m + 1; }
        const int r = lo - 1;  // == pv - 1 if no right code deviates <= d0
        const int up = vf_neighbour(R, S, t.a, t.b, r, true);
        const int dn = vf_neighbour(R, S, t.a, t.b, l, false);
        d1 = CUDART_INF_F;
        if (up >= 0) d1 = fminf(d1, vf_dev(R, up, med));
        if (dn >= 0) d1 = fminf(d1, vf_dev(R, dn, med));
    }
    return __fdiv_rn(__fadd_rn(d0, d1), 2.0f);
}

__global__ void __launch_bounds__(VF_THREADS, 4) validate_fast_kernel(VfastArgs A, adb_config cfg) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *winbuf = smem;
    __shared__ VfScratch S;
    __shared__ uint64_t bar_storage;
    uint64_t *bar = &bar_storage;
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t phase = 0;

    for (int r = blockIdx.x; r < A.B.n_reads; r += gridDim.x) {
        const int mb = r / A.B.batch_size;
        adb_record *rec = A.out + r;
        // generic-proxy accesses to the window memory of the previous read are ordered before the next bulk copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (A.batch_status[mb] != ADB_OK) {  // minibatch lost (host raises): zero record
            for (int w = tid; w < (int)(sizeof(adb_record) / 4); w += blockDim.x) ((uint32_t *)rec)[w] = 0;
            if (tid == 0) A.done[r] = 1;
            continue;
        }
        const ReadSrc gsrc = make_src(A.B, r);
        if (!(gsrc.i16 != nullptr && gsrc.cscale > 0.0f && isfinite(gsrc.cscale) && isfinite(gsrc.coff))) continue;
        const int full_len = A.B.full_lens[r];
        const int size = gsrc.n;
        const int *g = A.given + (size_t)r * A.given_stride;
        const int a_end = g[0], pe_best = g[1];
        const int n_topk = A.ntopk_per_read ? A.ntopk_per_read[r] : A.given_ntopk;
        const int pe0 = (n_topk >= 1) ? g[1] : 0;
        const int topk1 = (n_topk >= 2) ? g[2] : 0;
        const int msw = cfg.median_shift_window;
        const bool haveA = (a_end != 0);
        const bool mvs_geom = cfg.mvs_detect_check && !(pe0 == 0 || a_end == 0 || pe0 < a_end || pe0 - a_end <= 2) &&
                              !(size < a_end + msw);
        const bool win_var = !(pe0 - a_end <= cfg.pA_var_window + 2), win_mean = !(pe0 - a_end <= cfg.pA_mean_window + 2);
        float smed_var = 0.f, smed_mean = 0.f;  // medians of the two moving-statistics series (series_median_kernel)
        if (mvs_geom && (win_var || win_mean)) {
            const long long po = A.pre_off ? A.pre_off[r] : -1;
            if (!(po >= 0 && A.pre_meta[2 * r] == a_end && A.pre_meta[2 * r + 1] == pe0)) continue;  // not precomputed
            smed_var = A.series_med[2 * r];
            smed_mean = A.series_med[2 * r + 1];
        }
        // ---- stage the preload window in shared memory (TMA bulk copy) ----
        VfRead R;
        {
            unsigned char *w = cta_stage_window(winbuf, (const unsigned char *)gsrc.i16, size * 2, bar, phase);
            R.W16 = reinterpret_cast<const uint16_t *>(winbuf);
            R.s0 = (int)(w - winbuf) >> 1;
        }
        R.n = size; R.coff = gsrc.coff; R.cscale = gsrc.cscale;
        if (size <= 0) continue;
        vf_minmax(R, S, R.kmin, R.kmax);
        if (R.kmin < 0 || R.kmax >= VF_KEY_LIMIT) continue;  // codes must read as non-negative finite halves
        for (int w = tid; w < (int)(sizeof(adb_record) / 4); w += blockDim.x) ((uint32_t *)rec)[w] = 0;

        // ---- speculative inputs of the checks, all from the staged window ----
        int n_open = 0, op_last = 0;
        if (haveA && cfg.detect_open_pores) {
            int ok = 1;
            const int c200 = gsb_code_at(200.0f, false, R.coff, R.cscale, &ok);
            n_open = vf_open_pores(R, S, 0, a_end, c200, rec, &op_last);
        }
        const int a_start1 = (n_open > 0) ? op_last : 0;
        int ra = a_start1, rb = a_end;
        clip_seg(ra, rb, size);
        const int rlen = rb - ra;
        const bool rr_geom = haveA && cfg.real_signal_check && rlen >= 2 * cfg.mean_window;
        float rm0 = 0.f, rm1 = 0.f;
        if (rr_geom) vf_mean_pair(R, S, ra, rb - cfg.mean_window, cfg.mean_window, rm0, rm1);
        const int lrw = min(cfg.max_obs_local_range, rlen);
        // np.percentile positions of the two local ranges
        const int nLR = lrw;
        const double vLR85 = __dmul_rn((double)(nLR - 1), 0.85), vLR15 = __dmul_rn((double)(nLR - 1), 0.15);
        int pa_ = a_end, pb_ = pe0;
        clip_seg(pa_, pb_, size);
        const int nP = pb_ - pa_;
        const double vP85 = __dmul_rn((double)(nP - 1), 0.85), vP15 = __dmul_rn((double)(nP - 1), 0.15);

        // ---- round A: every rank ----
        int nt = 0, rot = 0;
        __syncthreads();
        const int tA0 = haveA ? vf_add_rank(S, R, nt, rot, 0, a_end, 0) : -1;
        const int tA1 = (haveA && a_start1 != 0) ? vf_add_rank(S, R, nt, rot, a_start1, a_end, 0) : -1;
        const int tL15 = rr_geom ? vf_add_rank(S, R, nt, rot, rb - lrw, rb, (int)floor(vLR15)) : -1;
        const int tL85 = rr_geom ? vf_add_rank(S, R, nt, rot, rb - lrw, rb, (int)floor(vLR85)) : -1;
        const bool needP = mvs_geom || (pe_best > a_end);
        const int tP = needP ? vf_add_rank(S, R, nt, rot, a_end, pe_best, 0) : -1;
        const int tP15 = (mvs_geom && nP > 0) ? vf_add_rank(S, R, nt, rot, a_end, pe0, (int)floor(vP15)) : -1;
        const int tP85 = (mvs_geom && nP > 0) ? vf_add_rank(S, R, nt, rot, a_end, pe0, (int)floor(vP85)) : -1;
        const int tAF = mvs_geom ? vf_add_rank(S, R, nt, rot, a_end, min(a_end + msw, size), 0) : -1;
        const int tBF = mvs_geom ? vf_add_rank(S, R, nt, rot, max(a_end - msw, 0), a_end, 0) : -1;
        const int tR = (size > pe_best) ? vf_add_rank(S, R, nt, rot, pe_best, size, 0) : -1;
        const bool ms_geom = cfg.detect_med_shift && haveA;
        const int tMA = ms_geom ? vf_add_rank(S, R, nt, rot, a_end, min(a_end + cfg.med_shift_window, full_len), 0) : -1;
        const int tMB = ms_geom ? vf_add_rank(S, R, nt, rot, max(a_end - cfg.med_shift_window, 0), a_end, 0) : -1;
        __syncthreads();
        if (tid < nt) {  // medians: rank (n - 1) / 2 (the percentile tasks carry their own rank)
            VfTask &t = S.task[tid];
            if (tid != tL15 && tid != tL85 && tid != tP15 && tid != tP85) t.k = (t.b - t.a - 1) / 2;
        }
        if (tid == 0) { S.ntask = nt; S.vtotal = rot; }
        vf_task_bounds(R, S);
        vf_run(R, S);
        const float medA0 = vf_median_of(R, S, tA0), medA1 = vf_median_of(R, S, tA1);
        const float medP = vf_median_of(R, S, tP), medR = vf_median_of(R, S, tR);
        const float medAF = vf_median_of(R, S, tAF), medBF = vf_median_of(R, S, tBF);
        const float medMA = vf_median_of(R, S, tMA), medMB = vf_median_of(R, S, tMB);
        auto local_range = [&](int q15, int q85, int n, double v15, double v85) -> double {
            if (q15 < 0 || q85 < 0) return CUDART_NAN;
            const int l15 = (int)floor(v15), l85 = (int)floor(v85);
            int a15, b15, a85, b85;
            vf_rank_pair(R, S, q15, min(l15 + 1, n - 1) != l15, a15, b15);
            vf_rank_pair(R, S, q85, min(l85 + 1, n - 1) != l85, a85, b85);
            const double p85 = np_lerp_f32(vf_pa(R, a85), vf_pa(R, b85), __dsub_rn(v85, (double)l85));
            const double p15 = np_lerp_f32(vf_pa(R, a15), vf_pa(R, b15), __dsub_rn(v15, (double)l15));
            return __dsub_rn(p85, p15);
        };
        const double lrA = local_range(tL15, tL85, nLR, vLR15, vLR85);
        const double lrP = local_range(tP15, tP85, nP, vP15, vP85);

        // ---- round B: every MAD ----
        int bmin[4] = {0, 0, 0, 0}, bmax[4] = {0, 0, 0, 0};
        {
            const int src[4] = {tA0, tA1, tP, tR};
            for (int q = 0; q < 4; q++) if (src[q] >= 0) { bmin[q] = S.task[src[q]].smin; bmax[q] = S.task[src[q]].smax; }
        }
        nt = 0; rot = 0;
        __syncthreads();
        const int dA0 = (tA0 >= 0) ? vf_add_mad(S, R, nt, rot, 0, a_end, medA0, bmin[0], bmax[0]) : -1;
        const int dA1 = (tA1 >= 0) ? vf_add_mad(S, R, nt, rot, a_start1, a_end, medA1, bmin[1], bmax[1]) : -1;
        const int dP = (tP >= 0) ? vf_add_mad(S, R, nt, rot, a_end, pe_best, medP, bmin[2], bmax[2]) : -1;
        const int dR = (tR >= 0) ? vf_add_mad(S, R, nt, rot, pe_best, size, medR, bmin[3], bmax[3]) : -1;
        __syncthreads();
        if (tid == 0) { S.ntask = nt; S.vtotal = rot; }
        vf_run(R, S);
        const float madA0 = vf_mad_of(R, S, dA0), madA1 = vf_mad_of(R, S, dA1);
        const float madP = vf_mad_of(R, S, dP), madR = vf_mad_of(R, S, dR);

        // ---- the checks (combined.py:394-580), part 1: everything that decides adapter_start ----
        int a_start = 0;
        bool success = true;
        int fail = ADB_FAIL_NONE, fail_mask = 0;
        uint32_t valid = ADB_V_FIELDS;
        double mvs_v[5] = {0, 0, 0, 0, 0}, real_v[3] = {0, 0, 0}, med_shift = 0.0;
        int n_open_rep = 0;
        if (a_end == 0) { success = false; fail = ADB_FAIL_NO_ADAPTER; }
        if (success && (madA0 != 0.0f) && !in_range_d((double)madA0, cfg.adapter_mad_range)) { success = false; fail = ADB_FAIL_ADAPTER_MAD; }
        if (success && cfg.detect_open_pores) {
            n_open_rep = n_open;
            valid |= ADB_V_OPEN_PORES;
            if (n_open > 0) {
                a_start = op_last;
                if (a_end - a_start < cfg.min_obs_adapter) { success = false; fail = ADB_FAIL_OPEN_PORE; }
            }
        }
        // partition sums (signal_partitions.py:65-96) while the window is still staged
        double pmean[3], pstd[3];
        {
            const int sa[3] = {a_start, a_end, pe_best}, sb[3] = {a_end, pe_best, size};
            const bool on[3] = {a_end > a_start, pe_best > a_end, size > pe_best};
            vf_mean_std3(R, S, sa, sb, on, pmean, pstd);
        }
        if (success && cfg.real_signal_check) {
            if (rlen < 2 * cfg.mean_window) {
                success = false; fail = ADB_FAIL_REAL_RANGE;
            } else {
                real_v[0] = (double)rm0; real_v[1] = (double)rm1;
                valid |= ADB_V_REAL_MEANS;
                if (in_range_d((double)rm0, cfg.mean_start_range) && in_range_d((double)rm1, cfg.mean_end_range)) {
                    real_v[2] = lrA;
                    valid |= ADB_V_REAL_RANGE;
                    if (!in_range_d(lrA, cfg.local_range)) { success = false; fail = ADB_FAIL_REAL_RANGE; }
                } else {
                    success = false; fail = ADB_FAIL_REAL_RANGE;
                }
            }
        }
        bool exception = false, defer = false, need_mvs = false, followup = false;
        double mlo = cfg.pA_mean_range[0], mhi = cfg.pA_mean_range[1];
        if (success && cfg.mvs_detect_check) {
            if (pe_best == 0) {
                success = false; fail = ADB_FAIL_NO_POLYA;
            } else {
                if (cfg.pA_mean_range_empty && !cfg.pA_mean_scale_range_empty) {
                    mlo = __dmul_rn(cfg.pA_mean_scale_range[0], (double)medA0);
                    mhi = __dmul_rn(cfg.pA_mean_scale_range[1], (double)medA0);
                } else if (cfg.pA_mean_range_empty) {
                    exception = true; fail = ADB_FAIL_EXC_PA_MEAN_RANGE;
                }
                if (!exception && n_topk < 0) { exception = true; fail = ADB_FAIL_EXC_TOPK_NONE; }
                need_mvs = !exception && n_topk >= 1 && pe0 != 0;
            }
        }
        if (need_mvs) {
            // first candidate (mvs.py:45-158); further candidates after a failure are left to validate_kernel
            valid |= ADB_V_MVS;
            bool ok = false;
            if (mvs_geom) {
                const int L = nP;
                __syncthreads();
                if (!win_var || !win_mean) {
                    // exact numpy mean / variance of a short segment (one thread, pairwise order)
                    if (tid == 0) {
                        const uint16_t *p = R.W16 + R.s0 + pa_;
                        const float co = R.coff, cs = R.cscale;
                        const float mean = __fdiv_rn(np_sum_f32([&](int i) { return __fmul_rn(__fadd_rn((float)(int)p[i], co), cs); }, L), (float)L);
                        S.ftmp[0] = mean;
                        S.ftmp[1] = __fdiv_rn(np_sum_f32([&](int i) { const float d = __fsub_rn(__fmul_rn(__fadd_rn((float)(int)p[i], co), cs), mean); return __fmul_rn(d, d); }, L), (float)L);
                    }
                    __syncthreads();
                }
                const float small_mean = S.ftmp[0], small_var = S.ftmp[1];
                __syncthreads();
                const float var32 = win_var ? smed_var : small_var;
                const float mean32 = win_mean ? smed_mean : small_mean;
                const float shift32 = __fsub_rn(medAF, medBF);
                mvs_v[0] = (double)mean32; mvs_v[1] = (double)var32; mvs_v[2] = (double)medP; mvs_v[3] = lrP; mvs_v[4] = (double)shift32;
                const double mr[2] = {mlo, mhi};
                int mask = 0;
                if (!in_range_d(mvs_v[0], mr)) mask |= 1;
                if (!in_range_d(mvs_v[1], cfg.pA_var_range)) mask |= 2;
                if (!in_range_d(mvs_v[2], cfg.polyA_med_range)) mask |= 4;
                if (!in_range_d(mvs_v[3], cfg.polyA_local_range)) mask |= 8;
                if (!in_range_d(mvs_v[4], cfg.median_shift_range)) mask |= 16;
                ok = (mask == 0);
                if (!ok) {
                    success = false;
                    if (mvs_v[0] == 0.0) { fail = ADB_FAIL_MVS_NOT_ENOUGH; fail_mask = 0; }  // combined.py:492-495 keys on the value
                    else { fail = ADB_FAIL_MVS_CHECKS; fail_mask = mask; }
                }
            } else {
                success = false; fail = ADB_FAIL_MVS_NOT_ENOUGH; fail_mask = 0;
            }
            if (!ok && topk1 != 0) {
                // the reference goes on to the next candidates (combined.py:464-515): their moving statistics are
                // prefixes of one more series pass, so validate_cand_kernel finishes the read from this record
                const bool fits = A.cand_followup != 0;  // (series beyond the window memory are bisected from the pools there)
                if (fits) followup = true; else defer = true;
            }
        }
        if (!exception && success && cfg.detect_med_shift) {
            const float sh = __fsub_rn(medMA, medMB);
            med_shift = (double)sh;
            valid |= ADB_V_MED_SHIFT;
            if (!in_range_d(med_shift, cfg.med_shift_range)) { success = false; fail = ADB_FAIL_MED_SHIFT; }
        }
        if (A.mode == ADB_METHOD_CNN && cfg.fallback_to_llr_short_reads && !exception && !success && a_end > 0 && pe_best > 0 &&
            pe_best - a_end > 1000 && full_len < 2 * cfg.max_obs_adapter)
            defer = true;  // "hail mary" LLR fallback (combined.py:251-301) lives in validate_kernel
        if (defer) continue;  // (uniform) validate_kernel redoes this read from scratch
        __syncthreads();
        if (!(valid & ADB_V_OPEN_PORES) || exception) {
            if (tid < ADB_MAX_OPEN_PORES) rec->open_pores[tid] = 0;  // the scan was speculative
        }
        if (exception) {
            // combined.py:225-226 / 304-305 / 350-351: DetectResults(success=False, fail_reason=str(e)), all else None
            if (tid == 0) {
                rec->success = 0; rec->fail_code = fail; rec->mvs_fail_mask = 0; rec->valid = 0;
                rec->signal_len = full_len; rec->preloaded = min(full_len, size);
                A.done[r] = 1;
            }
            continue;
        }
        if (tid == 0) {
            double st[3][4];
            for (int p = 0; p < 3; p++) for (int q = 0; q < 4; q++) st[p][q] = 0.0;
            if (a_end > a_start) {
                st[0][0] = pmean[0]; st[0][1] = pstd[0];
                st[0][2] = (double)(a_start == 0 ? medA0 : medA1);
                st[0][3] = (double)(a_start == 0 ? madA0 : madA1);
                valid |= ADB_V_ADAPTER_STATS;
            }
            if (pe_best > a_end) {
                st[1][0] = pmean[1]; st[1][1] = pstd[1]; st[1][2] = (double)medP; st[1][3] = (double)madP;
                valid |= ADB_V_POLYA_STATS;
            }
            if (size > pe_best) {
                st[2][0] = pmean[2]; st[2][1] = pstd[2]; st[2][2] = (double)medR; st[2][3] = (double)madR;
                valid |= ADB_V_RNA_STATS;
            }
            rec->success = success ? 1 : 0;
            rec->fail_code = fail;
            rec->mvs_fail_mask = fail_mask;
            rec->valid = valid | (n_topk >= 0 ? ADB_V_CAND : 0);
            rec->signal_len = full_len;
            rec->preloaded = min(full_len, size);
            rec->adapter_start = a_start;
            rec->adapter_end = a_end;
            rec->polya_end = pe_best;
            rec->primary_adapter_end = a_end;
            rec->primary_polya_end = pe_best;
            rec->mvs_adapter_end = 0;
            rec->n_cand = max(n_topk, 0);
            for (int t = 0; t < ADB_MAX_CAND; t++) rec->cand[t] = (t < n_topk) ? g[1 + t] : 0;
            rec->n_open_pores = n_open_rep;
            for (int p = 0; p < 3; p++) for (int q = 0; q < 4; q++) rec->stats[p][q] = st[p][q];
            for (int i = 0; i < 5; i++) rec->mvs[i] = mvs_v[i];
            for (int i = 0; i < 3; i++) rec->real[i] = real_v[i];
            rec->med_shift = med_shift;
            if (followup) *reinterpret_cast<float *>(rec->_reserved) = medA0;  // scales the mean range of the later candidates
            __threadfence();
            A.done[r] = followup ? 2 : 1;
        }
    }
}

// ---- further poly(A) candidates of the reads whose first one failed (CNN path, combined.py:464-515) ---------------------
// `success` is never reset once a candidate has failed, so the loop of the reference runs to its last non-zero
// candidate whatever happens: the mvs_detect_* values of the record are those of the LAST candidate, the fail reason
// that of the last candidate that failed (scanning backwards from the end until one fails).  Everything else of the
// record -- incl. polya_end = the first candidate and the partition statistics -- stands as validate_fast_kernel wrote
// it.  One evaluation = mean_var_shift_polyA_check (mvs.py:45-158) of one candidate: poly(A) median, p85 - p15, the two
// medians around adapter_end, and the medians of the moving statistics, which are prefixes of the rows the second
// mvs_series_kernel pass computed up to the largest candidate.  A read this kernel cannot settle (series not
// precomputed) goes back to validate_kernel (done = 0).
__global__ void __launch_bounds__(VF_THREADS, 4) validate_cand_kernel(VfastArgs A, adb_config cfg, const int *pending,
                                                                      const int *n_pending) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *winbuf = smem;
    const size_t win_cap = (((size_t)A.win_bytes + 48 + 15) & ~(size_t)15);
    __shared__ VfScratch S;
    __shared__ uint64_t bar_storage;
    uint64_t *bar = &bar_storage;
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t phase = 0;
    const int n_list = *n_pending;
    const int msw = cfg.median_shift_window;
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int r = pending[li];
        if (A.done[r] != 2) continue;
        adb_record *rec = A.out + r;
        const ReadSrc gsrc = make_src(A.B, r);
        const int size = gsrc.n;
        const int *g = A.given + (size_t)r * A.given_stride;
        const int a_end = g[0];
        const int n_topk = A.ntopk_per_read ? A.ntopk_per_read[r] : A.given_ntopk;
        int n_eval = 0;
        while (n_eval < n_topk && g[1 + n_eval] != 0) n_eval++;
        const long long po = A.pre_off ? A.pre_off[r] : -1;
        const bool have_rows = po >= 0 && A.pre_meta[2 * r] == a_end;
        const int row_pe = have_rows ? A.pre_meta[2 * r + 1] : 0;
        const float medA0 = *reinterpret_cast<const float *>(rec->_reserved);
        double mlo = cfg.pA_mean_range[0], mhi = cfg.pA_mean_range[1];
        if (cfg.pA_mean_range_empty && !cfg.pA_mean_scale_range_empty) {
            mlo = __dmul_rn(cfg.pA_mean_scale_range[0], (double)medA0);
            mhi = __dmul_rn(cfg.pA_mean_scale_range[1], (double)medA0);
        }
        double last_v[5] = {0, 0, 0, 0, 0};
        int fail = rec->fail_code, fail_mask = rec->mvs_fail_mask;  // the first candidate's, unless a later one failed
        bool bail = false, fail_known = false;
        for (int t = n_eval - 1; t >= 1 && !fail_known && !bail; t--) {
            const int pe = g[1 + t];
            double v[5] = {0, 0, 0, 0, 0};
            bool ok = false;
            const bool geom = !(pe == 0 || a_end == 0 || pe < a_end || pe - a_end <= 2) && !(size < a_end + msw);
            if (geom) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                VfRead R;
                {
                    unsigned char *w = cta_stage_window(winbuf, (const unsigned char *)gsrc.i16, size * 2, bar, phase);
                    R.W16 = reinterpret_cast<const uint16_t *>(winbuf);
                    R.s0 = (int)(w - winbuf) >> 1;
                }
                R.n = size; R.coff = gsrc.coff; R.cscale = gsrc.cscale;
                vf_minmax(R, S, R.kmin, R.kmax);  // (the first pass has verified 0 <= code < VF_KEY_LIMIT)
                int pa_ = a_end, pb_ = pe;
                clip_seg(pa_, pb_, size);
                const int nP = pb_ - pa_;
                const bool win_var = !(pe - a_end <= cfg.pA_var_window + 2), win_mean = !(pe - a_end <= cfg.pA_mean_window + 2);
                const int nv = win_var ? nP - (cfg.pA_var_window - 1) : 0, nm = win_mean ? nP - (cfg.pA_mean_window - 1) : 0;
                if ((win_var || win_mean) && !(have_rows && row_pe >= pe)) { bail = true; break; }
                const bool big = (size_t)(((max(nv, 0) + 3) & ~3) + max(nm, 0)) * 4 > win_cap;  // series beyond the window memory
                const double vP85 = __dmul_rn((double)(nP - 1), 0.85), vP15 = __dmul_rn((double)(nP - 1), 0.15);
                int nt = 0, rot = 0;
                __syncthreads();
                const int tP = vf_add_rank(S, R, nt, rot, a_end, pe, 0);
                const int tP15 = (nP > 0) ? vf_add_rank(S, R, nt, rot, a_end, pe, (int)floor(vP15)) : -1;
                const int tP85 = (nP > 0) ? vf_add_rank(S, R, nt, rot, a_end, pe, (int)floor(vP85)) : -1;
                const int tAF = vf_add_rank(S, R, nt, rot, a_end, min(a_end + msw, size), 0);
                const int tBF = vf_add_rank(S, R, nt, rot, max(a_end - msw, 0), a_end, 0);
                __syncthreads();
                if (tid < nt && tid != tP15 && tid != tP85) { VfTask &q = S.task[tid]; q.k = (q.b - q.a - 1) / 2; }
                if (tid == 0) { S.ntask = nt; S.vtotal = rot; }
                vf_task_bounds(R, S);
                vf_run(R, S);
                const float medP = vf_median_of(R, S, tP), medAF = vf_median_of(R, S, tAF), medBF = vf_median_of(R, S, tBF);
                double lrP = CUDART_NAN;
                if (tP15 >= 0 && tP85 >= 0) {
                    const int l15 = (int)floor(vP15), l85 = (int)floor(vP85);
                    int a15, b15, a85, b85;
                    vf_rank_pair(R, S, tP15, min(l15 + 1, nP - 1) != l15, a15, b15);
                    vf_rank_pair(R, S, tP85, min(l85 + 1, nP - 1) != l85, a85, b85);
                    const double p85 = np_lerp_f32(vf_pa(R, a85), vf_pa(R, b85), __dsub_rn(vP85, (double)l85));
                    const double p15 = np_lerp_f32(vf_pa(R, a15), vf_pa(R, b15), __dsub_rn(vP15, (double)l15));
                    lrP = __dsub_rn(p85, p15);
                }
                __syncthreads();
                if (!win_var || !win_mean) {  // exact numpy mean / variance of a short segment (one thread, pairwise order)
                    if (tid == 0) {
                        const uint16_t *p = R.W16 + R.s0 + pa_;
                        const float co = R.coff, cs = R.cscale;
                        const float mean = __fdiv_rn(np_sum_f32([&](int i) { return __fmul_rn(__fadd_rn((float)(int)p[i], co), cs); }, nP), (float)nP);
                        S.ftmp[0] = mean;
                        S.ftmp[1] = __fdiv_rn(np_sum_f32([&](int i) { const float d = __fsub_rn(__fmul_rn(__fadd_rn((float)(int)p[i], co), cs), mean); return __fmul_rn(d, d); }, nP), (float)nP);
                    }
                    __syncthreads();
                }
                const float small_mean = S.ftmp[0], small_var = S.ftmp[1];
                __syncthreads();
                float smed[2] = {0.f, 0.f};
                if (big) vf_series_medians<false>(S, A.pre_var + po, max(nv, 0), A.pre_mean + po, max(nm, 0), (uint32_t *)winbuf, smed);
                else if (nv > 0 || nm > 0) vf_series_medians<true>(S, A.pre_var + po, max(nv, 0), A.pre_mean + po, max(nm, 0), (uint32_t *)winbuf, smed);
                v[0] = (double)(win_mean ? smed[1] : small_mean);
                v[1] = (double)(win_var ? smed[0] : small_var);
                v[2] = (double)medP; v[3] = lrP; v[4] = (double)__fsub_rn(medAF, medBF);
                const double mr[2] = {mlo, mhi};
                int mask = 0;
                if (!in_range_d(v[0], mr)) mask |= 1;
                if (!in_range_d(v[1], cfg.pA_var_range)) mask |= 2;
                if (!in_range_d(v[2], cfg.polyA_med_range)) mask |= 4;
                if (!in_range_d(v[3], cfg.polyA_local_range)) mask |= 8;
                if (!in_range_d(v[4], cfg.median_shift_range)) mask |= 16;
                ok = (mask == 0);
                if (!ok) {
                    if (v[0] == 0.0) { fail = ADB_FAIL_MVS_NOT_ENOUGH; fail_mask = 0; }
                    else { fail = ADB_FAIL_MVS_CHECKS; fail_mask = mask; }
                    fail_known = true;
                }
            } else {  // the early return of mean_var_shift_polyA_check: zeros, "not enough signal"
                fail = ADB_FAIL_MVS_NOT_ENOUGH; fail_mask = 0;
                fail_known = true;
            }
            if (t == n_eval - 1) for (int i = 0; i < 5; i++) last_v[i] = v[i];
        }
        __syncthreads();
        if (tid == 0) {
            if (bail) {
                A.done[r] = 0;  // validate_kernel redoes the read from scratch
            } else {
                for (int i = 0; i < 5; i++) rec->mvs[i] = last_v[i];
                rec->fail_code = fail;
                rec->mvs_fail_mask = fail_mask;
                for (int i = 0; i < 8; i++) rec->_reserved[i] = 0;
                __threadfence();
                A.done[r] = 1;
            }
        }
    }
}
