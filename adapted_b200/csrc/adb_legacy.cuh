// Legacy three-split LLR detectors of the Cython module (no caller inside the reference; library-level operators):
//   c_llr_detect_adapter(raw_signal, min_obs_adapter, border_trim)                       _c_llr.pyx:239-288
//   c_llr_detect_adapter_polya(raw_signal, min_obs_adapter, border_trim, min_obs_polya)  _c_llr.pyx:290-363
// built on _best_split (_c_llr.pyx:40-64): the position of the largest positive LLR gain of a split range -- the
// "argmax of the LLR trace" of the north star.  One CTA per signal: sequential float64 prefix sums on two lanes, the
// gain loop spread over the CTA with a (gain, first index) arg-max reduction, the four segment medians by bisection
// on the ordered float64 bits.
#pragma once
#include "adb_common.cuh"
#include "adb_llr.cuh"

#define ADB_LEGACY_THREADS 128

__device__ __forceinline__ unsigned long long f64_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct LegacyScratch {
    double dred[ADB_LEGACY_THREADS / 32];
    int ired[ADB_LEGACY_THREADS / 32];
    unsigned long long kred[2];
    unsigned cnt[3];
    int flag;
};

// _best_split: x = -1, gain = 0 if no gain of the range is positive.  CTA-wide; results uniform.
__device__ void cta_best_split(const double *c, const double *c2, int start, int end, int head, int tail, int &x,
                               double &gain, LegacyScratch &S) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int i0 = start + head, i1 = end - tail;
    double best = 0.0;
    int bx = -1;
    if (i0 < i1) {
        const double var_summed = __dmul_rn((double)(end - start), log(var_c(start, end, c, c2)));
        for (int i = i0 + tid; i < i1; i += T) {
            const double h = __dmul_rn((double)(i - start), log(var_c(start, i, c, c2)));
            const double t = __dmul_rn((double)(end - i), log(var_c(i, end, c, c2)));
            const double g = __dsub_rn(var_summed, __dadd_rn(h, t));
            if (g > best) { best = g; bx = i; }  // strict: the first maximum of this thread's ascending positions
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(ADB_FULL, best, o);
        const int ox = __shfl_xor_sync(ADB_FULL, bx, o);
        if (ox >= 0 && (bx < 0 || ob > best || (ob == best && ox < bx))) { best = ob; bx = ox; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { S.dred[tid >> 5] = best; S.ired[tid >> 5] = bx; }
    __syncthreads();
    best = 0.0; bx = -1;
    for (int w = 0; w < (T >> 5); w++) {
        const double ob = S.dred[w];
        const int ox = S.ired[w];
        if (ox >= 0 && (bx < 0 || ob > best || (ob == best && ox < bx))) { best = ob; bx = ox; }
    }
    __syncthreads();
    x = bx;
    gain = (bx >= 0) ? best : 0.0;
}

// np.median(x[a:b]) for float64 (NaN if the slice is empty or holds a NaN).  CTA-wide; result uniform.
__device__ double cta_median_f64(const double *x, int a, int b, LegacyScratch &S) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int n = b - a;
    if (n <= 0) return CUDART_NAN;
    __syncthreads();
    if (tid == 0) { S.kred[0] = ~0ull; S.kred[1] = 0ull; S.flag = 0; }
    __syncthreads();
    unsigned long long mn = ~0ull, mx = 0ull;
    bool nan = false;
    for (int j = a + tid; j < b; j += T) {
        const double v = x[j];
        nan |= !(v == v);
        const unsigned long long k = f64_key(v);
        mn = min(mn, k); mx = max(mx, k);
    }
    atomicMin(&S.kred[0], mn);
    atomicMax(&S.kred[1], mx);
    if (nan) S.flag = 1;
    __syncthreads();
    if (S.flag) { __syncthreads(); return CUDART_NAN; }
    unsigned long long lo = S.kred[0], hi = S.kred[1];
    const unsigned k0 = (unsigned)(n - 1) / 2, k1 = (unsigned)n / 2;
    unsigned cnt_hi = (unsigned)n;
    int pass = 0;
    __syncthreads();
    if (tid < 3) S.cnt[tid] = 0;
    __syncthreads();
    while (lo < hi) {
        const unsigned long long mid = lo + ((hi - lo) >> 1);
        unsigned c = 0;
        for (int j = a + tid; j < b; j += T) c += (f64_key(x[j]) <= mid);
        c = __reduce_add_sync(ADB_FULL, c);
        if ((tid & 31) == 0 && c) atomicAdd(&S.cnt[pass % 3], c);
        __syncthreads();
        const unsigned tot = S.cnt[pass % 3];
        if (tot > k0) { hi = mid; cnt_hi = tot; } else lo = mid + 1;
        if (tid == 0) S.cnt[(pass + 2) % 3] = 0;
        pass++;
    }
    const double v0 = key_f64(hi);
    if (n & 1) return v0;
    double v1 = v0;
    if (!(cnt_hi > k1)) {
        __syncthreads();
        if (tid == 0) S.kred[0] = ~0ull;
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int j = a + tid; j < b; j += T) { const unsigned long long k = f64_key(x[j]); if (k > hi) best = min(best, k); }
        atomicMin(&S.kred[0], best);
        __syncthreads();
        v1 = key_f64(S.kred[0]);
        __syncthreads();
    }
    return __ddiv_rn(__dadd_rn(v0, v1), 2.0);
}

// params per signal: min_obs_adapter, border_trim, min_obs_polya (< 0: adapter only)
// out per signal: adapter_start, adapter_end, polya_end, tuple length of the reference's return value (2 or 3)
__global__ void __launch_bounds__(ADB_LEGACY_THREADS) llr_legacy_detect_kernel(const double *signals, const int64_t *offs,
                                                                               const int64_t *params, double *c_all,
                                                                               double *c2_all, int64_t *out) {
    __shared__ LegacyScratch S;
    const int t = blockIdx.x;
    const int64_t o = offs[t];
    const int n = (int)(offs[t + 1] - o);
    const int moa = (int)params[3 * t], bt = (int)params[3 * t + 1], mop = (int)params[3 * t + 2];
    const double *x = signals + o;
    double *c = c_all + o, *c2 = c2_all + o;
    int64_t *res = out + 4 * (size_t)t;
    if (threadIdx.x < 2) {
        const bool sq = threadIdx.x == 1;
        double *dst = sq ? c2 : c;
        double s = 0.0;
        for (int i = 0; i < n; i++) { double v = x[i]; if (sq) v = __dmul_rn(v, v); s = __dadd_rn(s, v); dst[i] = s; }
    }
    __threadfence_block();
    __syncthreads();
    const int length = n - 1;
    int x_first = -1, x_head = -1, x_tail = -1;
    double g_first = 0.0, gain_head = 0.0, gain_tail = 0.0;
    if (length > 0) cta_best_split(c, c2, 0, length, moa + bt, bt, x_first, g_first, S);
    if (x_first == -1) {  // empty signal: the reference returns (0, 0) from both functions (_c_llr.pyx:258-260, 313-315)
        if (threadIdx.x == 0) { res[0] = 0; res[1] = 0; res[2] = 0; res[3] = 2; }
        return;
    }
    cta_best_split(c, c2, 0, x_first, bt, moa, x_head, gain_head, S);
    cta_best_split(c, c2, x_first, length, moa, bt, x_tail, gain_tail, S);
    if (x_head == -1) x_head = 1;
    if (x_tail == -1) x_tail = x_first + 1;
    double med[4];
    med[0] = cta_median_f64(x, 0, min(x_head, n), S);
    med[1] = cta_median_f64(x, min(x_head, n), min(x_first, n), S);
    med[2] = cta_median_f64(x, min(x_first, n), min(x_tail, n), S);
    med[3] = cta_median_f64(x, min(x_tail, n), n, S);
    const double mean4 = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(med[0], med[1]), med[2]), med[3]), 4.0);
    int a_start = 0, a_end = 0;
    if (__dsub_rn(med[2], med[1]) > 0) {
        if (med[0] >= mean4) { a_start = x_head; a_end = x_first; } else { a_start = 0; a_end = x_first; }
    } else if (gain_tail > gain_head) {
        a_start = x_first; a_end = x_tail;
    }
    int polya_end = 0;
    if (mop >= 0 && a_end != 0) {
        double gp;
        cta_best_split(c, c2, a_end, length, mop, bt, polya_end, gp, S);
        if (polya_end == -1) polya_end = 0;
    }
    if (threadIdx.x == 0) { res[0] = a_start; res[1] = a_end; res[2] = polya_end; res[3] = 3; }
}
