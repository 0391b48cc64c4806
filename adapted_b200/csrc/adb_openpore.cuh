// Full open-pore lists for the reads whose list does not fit the fixed-size record (n_open_pores > ADB_MAX_OPEN_PORES).
//
// Reference: find_open_pores(signal[adapter_start:adapter_end]) adapted/detect/anomalies.py:15-35 as called from
// validate_boundaries (adapted/detect/combined.py:411-419).  With more than one hit the function reports every hit
// pos[i] (i >= 1) whose predecessor lies at least min_obs_diff = 10 samples back, i.e. the first sample of every run
// of hits except the very first hit.  The validate kernels keep the first ADB_MAX_OPEN_PORES of them in the record
// together with the true count; this kernel lists all of them for the (rare) reads beyond that, so that neither the
// python seam nor the table writer has to truncate or refuse them.  One CTA per selected read, two launches (count,
// then write at the prefix-summed offsets).
#pragma once
#include "adb_common.cuh"

#define ADB_OP_THREADS 256

struct OpenPoreArgs {
    BatchDev B;
    const int *seg;        // [n_sel][2] = scanned range [a, b) of the read (clipped to the preload window here)
    long long *counts;     // [n_sel] (count launch) or nullptr
    const long long *offs; // [n_sel + 1] (write launch)
    int *out;
    long long cap;
};

__global__ void __launch_bounds__(ADB_OP_THREADS) open_pores_full_kernel(OpenPoreArgs A) {
    __shared__ int wtot[ADB_OP_THREADS / 32];
    __shared__ int sh_first, sh_last, sh_hits;
    const int r = blockIdx.x, tid = threadIdx.x;
    const ReadSrc src = make_src(A.B, r);
    int a = A.seg[2 * r], b = A.seg[2 * r + 1];
    a = max(a, 0);
    b = min(b, src.n);
    const int n = max(b - a, 0);
    const int chunk = (n + ADB_OP_THREADS - 1) / ADB_OP_THREADS;
    const int j0 = min(tid * chunk, n), j1 = min(j0 + chunk, n);
    if (tid == 0) { sh_first = 0x7fffffff; sh_last = -1; sh_hits = 0; }
    __syncthreads();
    auto hit = [&](int j) { return src.pa(a + j) >= 200.0f; };
    auto run_start = [&](int j) {  // a hit whose previous hit lies >= 10 samples back (or that has none)
        if (!hit(j)) return false;
        for (int k = 1; k < 10; k++)
            if (j - k >= 0 && hit(j - k)) return false;
        return true;
    };
    int first = 0x7fffffff, last = -1, hits = 0, ncand = 0;
    for (int j = j0; j < j1; j++) {
        if (hit(j)) { first = min(first, j); last = j; hits++; }
        if (run_start(j)) ncand++;
    }
    if (hits) { atomicMin(&sh_first, first); atomicMax(&sh_last, last); atomicAdd(&sh_hits, hits); }
    __syncthreads();
    const int first_hit = sh_first;
    if (first_hit >= j0 && first_hit < j1) ncand--;  // the first hit is never reported (the loop starts at i = 1)
    int incl = ncand;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(ADB_FULL, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) wtot[tid >> 5] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < ADB_OP_THREADS / 32; w++) {
        if (w < (tid >> 5)) wbase += wtot[w];
        total += wtot[w];
    }
    // a single hit is returned as it is; several hits without a reported run start give [last hit] (anomalies.py:30-31)
    const int single = (sh_hits == 1) ? first_hit : ((sh_hits > 1 && total == 0) ? sh_last : -1);
    if (A.counts) {
        if (tid == 0) A.counts[r] = single >= 0 ? 1 : total;
        return;
    }
    if (single >= 0) {
        if (tid == 0 && A.offs[r] < A.cap) A.out[A.offs[r]] = a + single;
        return;
    }
    long long pos = A.offs[r] + wbase + incl - ncand;
    for (int j = j0; j < j1; j++) {
        if (j == first_hit) continue;
        if (run_start(j)) {
            if (pos < A.cap) A.out[pos] = a + j;
            pos++;
        }
    }
}
