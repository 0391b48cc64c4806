// CTA-wide exact order statistics over an implicit sequence of 32-bit keys.
//
// Multi-level histogram (radix) select with an adaptive window: level 0 bins the key range [kmin, kmax] into
// NB bins of 2^s keys; every requested rank is located by a prefix sum over the bins and, while s > 0, refined
// inside its bin.  For raw int16 ADC keys the per-read range is ~1-2 k values, so one pass resolves the rank;
// for float32 keys of ADC-quantised pA data two to three passes do.  Ties, duplicates and adversarial inputs
// only change the number of passes, never the result: it is the exact order statistic that np.partition
// (np.median / np.percentile) returns.
#pragma once
#include "adb_common.cuh"

#define ADB_SEL_NB 2048
#define ADB_SEL_MAXRANKS 4

struct SelScratch {
    uint32_t *hist;      // [ADB_SEL_NB] shared
    uint32_t *warp_tot;  // [32] shared
    // work list of refinement groups (shared)
    uint32_t *g_lo;      // [8]
    uint32_t *g_span;    // [8]  span-1 (inclusive width), so a full 2^32 range fits
    int *g_rbeg, *g_rend, *g_off;  // [8] each
    int *g_count;        // [1]
    int *r_bin, *r_before;  // [ADB_SEL_MAXRANKS]
};

#define ADB_SEL_SMEM_BYTES (ADB_SEL_NB * 4 + 32 * 4 + 8 * 4 * 5 + 16 + ADB_SEL_MAXRANKS * 8)

__device__ __forceinline__ SelScratch sel_scratch_from(unsigned char *base) {
    SelScratch S;
    S.hist = (uint32_t *)base;
    S.warp_tot = S.hist + ADB_SEL_NB;
    S.g_lo = S.warp_tot + 32;
    S.g_span = S.g_lo + 8;
    S.g_rbeg = (int *)(S.g_span + 8);
    S.g_rend = S.g_rbeg + 8;
    S.g_off = S.g_rend + 8;
    S.g_count = S.g_off + 8;
    S.r_bin = S.g_count + 4;
    S.r_before = S.r_bin + ADB_SEL_MAXRANKS;
    return S;
}

// min / max key over the sequence (all threads call; result broadcast)
template <class KeyF>
__device__ void cta_key_minmax(KeyF key, int n, uint32_t &kmin, uint32_t &kmax, SelScratch &S) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    {
        const int T = blockDim.x;
        int j = threadIdx.x;
        for (; j + 3 * T < n; j += 4 * T) {
            const uint32_t k0 = key(j), k1 = key(j + T), k2 = key(j + 2 * T), k3 = key(j + 3 * T);
            lo = min(min(lo, k0), min(k1, min(k2, k3)));
            hi = max(max(hi, k0), max(k1, max(k2, k3)));
        }
        for (; j < n; j += T) {
            const uint32_t k = key(j);
            lo = min(lo, k);
            hi = max(hi, k);
        }
    }
    lo = warp_min_u(lo);
    hi = warp_max_u(hi);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        S.warp_tot[threadIdx.x >> 5] = lo;
        S.hist[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    lo = 0xffffffffu; hi = 0u;
    for (int w = 0; w < nw; w++) {
        lo = min(lo, S.warp_tot[w]);
        hi = max(hi, S.hist[w]);
    }
    __syncthreads();
    kmin = lo;
    kmax = hi;
}

// Exact keys at the (ascending, 0-based) ranks[0..nr) of the n keys key(0..n).  kmin/kmax must bound all keys
// (they need not be tight).  All threads of the CTA call; out[] is written by thread 0 into shared memory the
// caller provides and is valid after the function returns (it ends with __syncthreads()).
template <class KeyF>
__device__ void cta_select_ranks(KeyF key, int n, uint32_t kmin, uint32_t kmax, const int *ranks, int nr,
                                 uint32_t *out, SelScratch &S) {
    const int T = blockDim.x, tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) {
        S.g_lo[0] = kmin;
        S.g_span[0] = kmax - kmin;
        S.g_rbeg[0] = 0;
        S.g_rend[0] = nr;
        S.g_off[0] = 0;
        *S.g_count = 1;
    }
    __syncthreads();
    while (true) {
        int gc = *S.g_count;
        if (gc == 0) break;
        // pop the last group (all threads read the same values)
        const uint32_t lo = S.g_lo[gc - 1], span = S.g_span[gc - 1];
        const int rbeg = S.g_rbeg[gc - 1], rend = S.g_rend[gc - 1], off = S.g_off[gc - 1];
        int s = 0;
        while ((span >> s) >= (uint32_t)ADB_SEL_NB) s++;
        __syncthreads();
        for (int b = tid; b < ADB_SEL_NB; b += T) S.hist[b] = 0;
        if (tid == 0) *S.g_count = gc - 1;
        __syncthreads();
        {
            // keys below lo wrap to large values and are filtered by the span test; 4 loads in flight per thread
            int j = tid;
            for (; j + 3 * T < n; j += 4 * T) {
                const uint32_t d0 = key(j) - lo, d1 = key(j + T) - lo, d2 = key(j + 2 * T) - lo, d3 = key(j + 3 * T) - lo;
                if (d0 <= span) atomicAdd(&S.hist[d0 >> s], 1u);
                if (d1 <= span) atomicAdd(&S.hist[d1 >> s], 1u);
                if (d2 <= span) atomicAdd(&S.hist[d2 >> s], 1u);
                if (d3 <= span) atomicAdd(&S.hist[d3 >> s], 1u);
            }
            for (; j < n; j += T) {
                const uint32_t d = key(j) - lo;
                if (d <= span) atomicAdd(&S.hist[d >> s], 1u);
            }
        }
        __syncthreads();
        // exclusive prefix over the bins: each thread owns a contiguous chunk
        const int per = (ADB_SEL_NB + T - 1) / T;
        const int b0 = tid * per, b1 = min(b0 + per, ADB_SEL_NB);
        uint32_t local = 0;
        for (int b = b0; b < b1; b++) local += S.hist[b];
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(ADB_FULL, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) S.warp_tot[tid >> 5] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < (tid >> 5); w++) wbase += S.warp_tot[w];
        const uint32_t excl = wbase + incl - local;
        for (int r = rbeg; r < rend; r++) {
            uint32_t rel = (uint32_t)(ranks[r] - off);
            if (rel >= excl && rel < excl + local) {
                uint32_t acc = excl;
                for (int b = b0; b < b1; b++) {
                    uint32_t h = S.hist[b];
                    if (rel < acc + h) {
                        S.r_bin[r] = b;
                        S.r_before[r] = (int)acc;
                        break;
                    }
                    acc += h;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (s == 0) {
                for (int r = rbeg; r < rend; r++) out[r] = lo + (uint32_t)S.r_bin[r];
            } else {
                int g = *S.g_count;
                int r = rbeg;
                while (r < rend) {
                    int r2 = r + 1;
                    while (r2 < rend && S.r_bin[r2] == S.r_bin[r]) r2++;
                    uint32_t nlo = lo + ((uint32_t)S.r_bin[r] << s);
                    uint32_t nspan = min((1u << s) - 1u, span - ((uint32_t)S.r_bin[r] << s));
                    S.g_lo[g] = nlo;
                    S.g_span[g] = nspan;
                    S.g_rbeg[g] = r;
                    S.g_rend[g] = r2;
                    S.g_off[g] = off + S.r_before[r];
                    g++;
                    r = r2;
                }
                *S.g_count = g;
            }
        }
        __syncthreads();
    }
    __syncthreads();
}

// median of n float32 values with numpy semantics: odd -> s[n/2]; even -> f32(f32(a+b)/2).
// n == 0 -> NaN (np.median of an empty slice).  `kbuf` = 4 uint32 + 2 int of shared memory.
template <class KeyF, class ValF>
__device__ float cta_median_keys(KeyF key, ValF val_of_key, int n, uint32_t kmin, uint32_t kmax, SelScratch &S,
                                 uint32_t *kbuf) {
    if (n <= 0) return CUDART_NAN_F;
    int *ranks = (int *)(kbuf + 4);
    __syncthreads();
    if (threadIdx.x == 0) {
        ranks[0] = (n - 1) / 2;
        ranks[1] = n / 2;
    }
    __syncthreads();
    int nr = (n & 1) ? 1 : 2;
    cta_select_ranks(key, n, kmin, kmax, ranks, nr, kbuf, S);
    float a = val_of_key(kbuf[0]);
    if (n & 1) return a;
    float b = val_of_key(kbuf[1]);
    return __fdiv_rn(__fadd_rn(a, b), 2.0f);
}
