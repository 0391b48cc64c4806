// Streaming / read-until poly(A) detector: mean_var_shift_polyA_detect, adapted/detect/mvs.py:341-426.
// No caller inside the reference (a library operator for tools that keep an accumulating read-until cache); built
// on the pieces of the validate kernel: window staged in shared memory, bottleneck moving statistics, histogram-based
// exact medians / percentiles.
//
// One CTA per read.  The moving mean / variance of signal[min_obs_adapter:] are computed once, reduced to a bit mask
// "mean and variance in range" (the series themselves are not needed afterwards), and the reference's search loop --
// first match at or after `offset`, median / local-range / median-shift checks there, else advance the offset by
// search_increment_step past the rejected position -- runs on the mask with one parallel find-first per iteration.
#pragma once
#include "adb_validate.cuh"

struct StreamArgs {
    BatchDev B;
    int win_bytes;
    float *series;   // [gridDim.x][2][m] scratch (moving variance / mean; the mask reuses the variance row)
    int32_t *out;    // [n_reads] poly(A) start, 0: none
};

// in_range(np.float32 value, python-float bounds) compares in float32 (numpy casts the weak scalar, utils.py:16-28)
__device__ __forceinline__ bool in_range_f32w(float v, const double r[2]) { return (float)r[0] <= v && v <= (float)r[1]; }

__global__ void __launch_bounds__(ADB_VAL_THREADS, 3) mvs_stream_kernel(StreamArgs A, adb_stream_config P) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *winbuf = smem;
    const size_t win_cap = (((size_t)A.win_bytes + 48 + 15) & ~(size_t)15);
    unsigned char *scratch = smem + win_cap;
    const size_t scratch_sz = ((size_t)ADB_SEL_SMEM_BYTES + 64 + 15) & ~(size_t)15;
    unsigned char *small = scratch + scratch_sz;
    uint32_t *kbuf = (uint32_t *)small;
    int *itmp = (int *)(small + 32);
    double *dtmp = (double *)(small + 64);
    uint64_t *bar = (uint64_t *)(small + 192);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    uint32_t phase = 0;

    ValCtx C;
    C.cfg = nullptr;
    C.S = sel_scratch_from(scratch);
    C.kbuf = kbuf;
    C.itmp = itmp;
    C.dtmp = dtmp;
    C.series_a = A.series + (size_t)blockIdx.x * 2 * A.B.m;
    C.series_b = C.series_a + A.B.m;
    C.pre_var = C.pre_mean = nullptr;
    C.pre_ae = C.pre_pe = -1;
    const int wm = P.pA_mean_window, wv = P.pA_var_window, moa = P.min_obs_adapter;

    for (int r = blockIdx.x; r < A.B.n_reads; r += gridDim.x) {
        __syncthreads();
        const ReadSrc gsrc = make_src(A.B, r);
        const int n = gsrc.n;  // calibrated_signal.size
        if (n < moa + max(max(wm, wv), max(P.min_obs_post_loc, P.polyA_window))) {  // mvs.py:352-363
            if (threadIdx.x == 0) A.out[r] = 0;
            continue;
        }
        ReadSrc src = gsrc;
        {
            const int esz = gsrc.f32 ? 4 : 2;
            const unsigned char *g = gsrc.f32 ? (const unsigned char *)gsrc.f32 : (const unsigned char *)gsrc.i16;
            unsigned char *w = cta_stage_window(winbuf, g, gsrc.n * esz, bar, phase);
            if (gsrc.f32) src.f32 = (const float *)w; else src.i16 = (const int16_t *)w;
        }
        C.src = src;
        C.int_keys = (src.i16 != nullptr) && (src.cscale > 0.0f);
        val_window_bounds(C);
        const int L = n - moa;
        seg_moving_stats(C, moa, L, wv, wm, true, true);
        // match mask over the series index i (position moa + i): bit set iff both moving statistics are in range;
        // entries before a window is full are NaN in the reference and never match
        uint32_t *mask = (uint32_t *)C.series_a;  // written word by word after the word's 32 entries were read
        const int nwords = (L + 31) >> 5;
        const int i0 = max(wv, wm) - 1;
        for (int w0 = 0; w0 < nwords; w0 += blockDim.x) {
            const int w = w0 + threadIdx.x;
            uint32_t bits = 0;
            if (w < nwords) {
                for (int b = 0; b < 32; b++) {
                    const int i = w * 32 + b;
                    if (i >= i0 && i < L) {
                        const float mv = C.series_a[i - (wv - 1)], mm = C.series_b[i - (wm - 1)];
                        if (in_range_f32w(mm, P.pA_mean_range) && in_range_f32w(mv, P.pA_var_range)) bits |= 1u << b;
                    }
                }
            }
            // series_a[j] is read for j = i - (wv - 1) <= i: a word is only overwritten by masks of entries at or
            // before it, and every thread finishes reading its 32 entries before the block-wide barrier
            __syncthreads();
            if (w < nwords) mask[w] = bits;
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();
        int offset = max(wm, wv), result = 0;
        while (offset < L) {  // mvs.py:381-425
            // first match at or after offset
            if (threadIdx.x == 0) itmp[7] = 0x7fffffff;
            __syncthreads();
            for (int w = (offset >> 5) + threadIdx.x; w < nwords; w += blockDim.x) {
                uint32_t bits = mask[w];
                if (w == (offset >> 5)) bits &= 0xffffffffu << (offset & 31);
                if (bits) { atomicMin(&itmp[7], w * 32 + __ffs(bits) - 1); break; }
            }
            __syncthreads();
            const int hit = itmp[7];
            __syncthreads();
            if (hit == 0x7fffffff) break;                     // "didn't find a match, return 0"
            const int idx = moa + hit;
            if (n - idx < P.min_obs_post_loc) break;          // not enough signal left: return 0
            const SegStats Q = seg_stats(C, idx, min(idx + P.polyA_window, n), SS_MED | SS_LR);
            const float m_after = seg_stats(C, idx, min(idx + P.median_shift_window, n), SS_MED).med;
            const float m_before = seg_stats(C, max(idx - P.median_shift_window, 0), idx, SS_MED).med;
            const double shift = (double)__fsub_rn(m_after, m_before);
            if (in_range_f32w(Q.med, P.polyA_med_range) && in_range_d(Q.lr, P.polyA_local_range) &&
                in_range_d(shift, P.median_shift_range)) {
                result = idx;
                break;
            }
            offset = idx - moa + P.search_increment_step;
        }
        if (threadIdx.x == 0) A.out[r] = result;
    }
}
