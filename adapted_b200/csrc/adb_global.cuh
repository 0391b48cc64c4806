// Minibatch-global median / MAD of normalize_signal(batch[:, :max_obs_trace], with_nan=True)
// (adapted/detect/normalize.py:15-22,54; call site combined.py:128-132): ONE exact order statistic over all
// non-NaN samples of the first `max_obs_trace` columns of every read of a minibatch (~25 M values).
//
// Exact 3-pass radix select (11 + 11 + 10 key bits) with one histogram set per minibatch in global memory:
//   gsel_hist  : streams the samples (HBM/L2 bound), filters by the prefix found so far, accumulates a per-CTA
//                shared-memory histogram with run-length aggregated atomics, flushes it to the global histogram;
//   gsel_scan  : one CTA per minibatch locates the bins of the two middle ranks and extends the prefixes.
// The same two kernels run a second time on |x - med| for the MAD.  All minibatches of a call are processed
// by the same launches (blockIdx.y = minibatch).
#pragma once
#include "adb_common.cuh"

#define GSEL_BINS 2048

struct GselState {       // one per minibatch, device memory
    unsigned long long count;  // number of non-NaN samples
    unsigned long long rank[2];// remaining rank inside the current prefix bucket, for the two middle ranks
    uint32_t prefix[2];        // key prefix (high bits) fixed so far
    float med, mad;
    int status;                // adb_status
    int _pad;
};

__device__ __forceinline__ int gsel_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ uint32_t gsel_mask(int pass) { return pass == 2 ? 1023u : 2047u; }

// stage: 0 = median of x, 1 = median of |x - med|
// hist layout: [minibatch][2 targets][GSEL_BINS]
__global__ void __launch_bounds__(256) gsel_hist_kernel(BatchDev B, int max_obs_trace, int stage, int pass,
                                                        const GselState *states, unsigned int *hist, const int *active) {
    __shared__ unsigned int sh[2][GSEL_BINS];
    const int mb = blockIdx.y;
    if (active && !active[mb]) return;  // this minibatch was settled by the sampled one-pass select (adb_gsample.cuh)
    const int r0 = mb * B.batch_size, r1 = min(r0 + B.batch_size, B.n_reads);
    const GselState st = states[mb];
    for (int b = threadIdx.x; b < 2 * GSEL_BINS; b += blockDim.x) (&sh[0][0])[b] = 0;
    __syncthreads();
    const int shift = gsel_shift(pass);
    const uint32_t mask = gsel_mask(pass);
    // prefix comparison: bits above (shift + width of this digit)
    const int pshift = (pass == 0) ? 32 : (pass == 1 ? 21 : 10);
    const bool same = (pass == 0) || (st.prefix[0] == st.prefix[1]);
    const float med = st.med;
    // run-length aggregation: consecutive samples mostly fall into the same bin, so a thread only touches the
    // shared histogram when its bin changes
    int cur0 = -1, cnt0 = 0, cur1 = -1, cnt1 = 0;
    auto consume = [&](float v) {
        if (!(v == v)) return;
        if (stage == 1) v = fabsf(__fsub_rn(v, med));
        const uint32_t k = f32_key(v);
        const uint32_t hi = (pshift >= 32) ? 0u : (k >> pshift);
        const int bin = (int)((k >> shift) & mask);
        if (pass == 0 || hi == st.prefix[0]) {
            if (bin == cur0) cnt0++;
            else { if (cnt0) atomicAdd(&sh[0][cur0], (unsigned)cnt0); cur0 = bin; cnt0 = 1; }
        }
        if (!same && hi == st.prefix[1]) {
            if (bin == cur1) cnt1++;
            else { if (cnt1) atomicAdd(&sh[1][cur1], (unsigned)cnt1); cur1 = bin; cnt1 = 1; }
        }
    };
    for (int r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
        const ReadSrc src = make_src(B, r);
        const int n = min(src.n, max_obs_trace);
        if (n <= 0) continue;
        if (src.i16) {
            // 16-byte vector loads (8 samples) over the aligned body, scalar head / tail
            const int16_t *p = src.i16;
            const int head = min(n, (int)(((16 - ((uintptr_t)p & 15)) & 15) >> 1));
            const int nvec = (n - head) >> 3;
            const uint4 *pv = (const uint4 *)(p + head);
            const float co = src.coff, cs = src.cscale;
            for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
                const uint4 q = __ldg(pv + v);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    consume(__fmul_rn(__fadd_rn((float)(int16_t)(w[t] & 0xffffu), co), cs));
                    consume(__fmul_rn(__fadd_rn((float)(int16_t)(w[t] >> 16), co), cs));
                }
            }
            const int tail0 = head + (nvec << 3);
            for (int j = threadIdx.x; j < head; j += blockDim.x) consume(src.pa(j));
            for (int j = tail0 + threadIdx.x; j < n; j += blockDim.x) consume(src.pa(j));
        } else {
            const float *p = src.f32;
            const int head = min(n, (int)(((16 - ((uintptr_t)p & 15)) & 15) >> 2));
            const int nvec = (n - head) >> 2;
            const float4 *pv = (const float4 *)(p + head);
            for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
                const float4 q = __ldg(pv + v);
                consume(q.x); consume(q.y); consume(q.z); consume(q.w);
            }
            const int tail0 = head + (nvec << 2);
            for (int j = threadIdx.x; j < head; j += blockDim.x) consume(p[j]);
            for (int j = tail0 + threadIdx.x; j < n; j += blockDim.x) consume(p[j]);
        }
    }
    if (cnt0) atomicAdd(&sh[0][cur0], (unsigned)cnt0);
    if (cnt1) atomicAdd(&sh[1][cur1], (unsigned)cnt1);
    __syncthreads();
    unsigned int *gh = hist + (size_t)mb * 2 * GSEL_BINS;
    for (int b = threadIdx.x; b < 2 * GSEL_BINS; b += blockDim.x) {
        unsigned v = (&sh[0][0])[b];
        if (v) atomicAdd(&gh[b], v);
    }
}

// one CTA (256 threads) per minibatch
__global__ void __launch_bounds__(256) gsel_scan_kernel(int stage, int pass, GselState *states, unsigned int *hist,
                                                        const int *active) {
    __shared__ unsigned long long part[256];
    __shared__ int found_bin[2];
    __shared__ unsigned long long found_before[2];
    const int mb = blockIdx.x, tid = threadIdx.x;
    if (active && !active[mb]) return;
    GselState *st = &states[mb];
    unsigned int *gh = hist + (size_t)mb * 2 * GSEL_BINS;
    const bool same = (pass == 0) || (st->prefix[0] == st->prefix[1]);
    if (pass == 0) {
        // total count = sum of the first histogram; ranks of the two middle elements
        unsigned long long s = 0;
        for (int b = tid; b < GSEL_BINS; b += 256) s += gh[b];
        part[tid] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (tid < o) part[tid] += part[tid + o]; __syncthreads(); }
        if (tid == 0) {
            unsigned long long n = part[0];
            st->count = n;
            st->rank[0] = n ? (n - 1) / 2 : 0;
            st->rank[1] = n / 2;
            st->prefix[0] = st->prefix[1] = 0;
        }
        __syncthreads();
    }
    const unsigned long long n = st->count;
    for (int t = 0; t < 2; t++) {
        const unsigned int *h = gh + ((same ? 0 : t) * GSEL_BINS);
        const unsigned long long rank = st->rank[t];
        // each thread owns 8 consecutive bins
        unsigned long long loc = 0;
        for (int b = tid * 8; b < tid * 8 + 8; b++) loc += h[b];
        part[tid] = loc;
        __syncthreads();
        if (tid == 0) {
            unsigned long long acc = 0;
            int ft = -1;
            for (int q = 0; q < 256; q++) {
                if (rank < acc + part[q]) { ft = q; break; }
                acc += part[q];
            }
            found_bin[t] = -1;
            if (ft >= 0) {
                for (int b = ft * 8; b < ft * 8 + 8; b++) {
                    if (rank < acc + h[b]) { found_bin[t] = b; found_before[t] = acc; break; }
                    acc += h[b];
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        if (n == 0 || found_bin[0] < 0 || found_bin[1] < 0) {
            // all-NaN minibatch: nanmedian is NaN, every read ends up with an empty trace
            if (stage == 0) st->med = CUDART_NAN_F; else st->mad = CUDART_NAN_F;
            if (st->status == ADB_OK) st->status = ADB_ERR_EMPTY_TRACE;
        } else {
            const int width = (pass == 2) ? 10 : 11;
            for (int t = 0; t < 2; t++) {
                st->prefix[t] = (st->prefix[t] << width) | (uint32_t)found_bin[t];
                st->rank[t] -= found_before[t];
            }
            if (pass == 2) {
                const float a = key_f32(st->prefix[0]), b = key_f32(st->prefix[1]);
                const float m = (n & 1ull) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
                if (stage == 0) st->med = m;
                else {
                    st->mad = m;
                    if (m == 0.0f && st->status == ADB_OK) st->status = ADB_ERR_MAD_ZERO;
                }
            }
        }
    }
    __syncthreads();
    for (int b = tid; b < 2 * GSEL_BINS; b += 256) gh[b] = 0;  // ready for the next pass
}
