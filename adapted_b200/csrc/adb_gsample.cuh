// Minibatch-global median / MAD in ONE pass over the raw samples (int16 sources).
//
// normalize_signal(batch[:, :max_obs_trace]) needs the exact median and the exact median of |x - med| of ~25 M
// samples per minibatch (adapted/detect/normalize.py:15-22,54; combined.py:128-132).  The multi-pass radix select of
// adb_global.cuh streams the data six times.  Here a 1/128 subsample fixes, per minibatch,
//     a value band  [m_lo, m_hi]   that holds the median            (with overwhelming probability), and
//     a deviation band [d_lo, d_hi] that holds the MAD              (for every median inside the value band),
// and a single streaming pass then
//     counts the samples below m_lo, the samples whose deviation is certainly below d_lo ("inside"),
//     and tallies -- per read and per ADC code -- the few samples that fall into the value band or into the two
//     bands [m_lo - d_hi, m_hi - d_lo) / (m_lo + d_lo, m_hi + d_hi] where the deviation may be the MAD.
// pA = (adc + offset) * scale is monotone in the ADC code, so every band is a short code interval per read and
// the tallies are exact (value, count) pairs.  A finishing CTA per minibatch selects the two middle ranks among
// them.  Every assumption is VERIFIED from the exact counts (rank inside the band, table not overflowing,
// calibration monotone); a minibatch that fails any check is handed to the exact multi-pass select, so the result
// is always the order statistic numpy returns -- the sample only decides how fast it is found.
#pragma once
#include "adb_common.cuh"
#include "adb_global.cuh"

#define GSB_BINS 16384          // sample histogram: 1/16 pA bins over [GSB_PA_MIN, GSB_PA_MIN + 1024)
#define GSB_BIN_PER_PA 16.0f
#define GSB_PA_MIN (-200.0f)
#define GSB_STRIDE 128          // one sample every GSB_STRIDE (> the dwell of a level: independent draws) ...
#define GSB_TARGET_SAMPLE 200000  // ... reduced for small minibatches so that the sample keeps about this size
#define GSB_MIN_SAMPLE 4096
#define GSB_Z 7.0f              // half-width of the rank bracket in standard deviations of the sample rank
#define GSB_TAB 128             // tally entries per read: [0,32) value band, [32,80) left band, [80,128) right band
#define GSB_TAB_M 32
#define GSB_TAB_L 48
#define GSB_TAB_R 48

struct GsbPlan {                // one per minibatch
    float m_lo, m_hi;           // value band (inclusive)
    float xL_lo, xI_lo, xI_hi, xR_hi;  // left band [xL_lo, xI_lo), inside [xI_lo, xI_hi], right band (xI_hi, xR_hi]
    float d_lo, d_hi;           // inside => dev < d_lo;  beyond the bands => dev > d_hi
    int use;                    // 1: sampled path planned
    int fallback;               // set when a check fails: the exact multi-pass select takes the minibatch
    unsigned long long n_total, c_below, c_inside;
};

__device__ __forceinline__ float gsb_pa(int c, float coff, float cscale) {
    return __fmul_rn(__fadd_rn((float)c, coff), cscale);
}

// smallest code c with pa(c) >= T (strict == false) or pa(c) > T (strict == true); pa is non-decreasing in c.
// The guess from the inverse map is corrected by evaluating pa exactly; *ok is cleared if that does not settle.
__device__ int gsb_code_at(float T, bool strict, float coff, float cscale, int *ok) {
    float g = ceilf(__fsub_rn(__fdiv_rn(T, cscale), coff));
    if (!(g == g)) { *ok = 0; return 0; }
    int c = (int)fminf(fmaxf(g, -40000.f), 40000.f);
    int it = 0;
    while (it < 64) {
        const float v = gsb_pa(c - 1, coff, cscale);
        if (strict ? (v > T) : (v >= T)) { c--; it++; } else break;
        if (c < -40000) break;
    }
    while (it < 64) {
        const float v = gsb_pa(c, coff, cscale);
        if (strict ? !(v > T) : !(v >= T)) { c++; it++; } else break;
        if (c > 40000) break;
    }
    if (it >= 64) *ok = 0;
    return c;
}

// ---- 1. sample histogram ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gsb_sample_kernel(BatchDev B, int max_obs_trace, int stride, unsigned *shist) {
    extern __shared__ unsigned gsb_sh[];
    const int mb = blockIdx.y;
    const int r0 = mb * B.batch_size, r1 = min(r0 + B.batch_size, B.n_reads);
    for (int b = threadIdx.x; b < GSB_BINS; b += blockDim.x) gsb_sh[b] = 0;
    __syncthreads();
    for (int r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
        const ReadSrc src = make_src(B, r);
        const int n = min(src.n, max_obs_trace);
        for (int j = ((r * 37) % stride) + threadIdx.x * stride; j < n; j += blockDim.x * stride) {
            const float v = src.pa(j);
            if (!(v == v)) continue;
            const float f = floorf((v - GSB_PA_MIN) * GSB_BIN_PER_PA);
            const int bin = (int)fminf(fmaxf(f, 0.f), (float)(GSB_BINS - 1));
            atomicAdd(&gsb_sh[bin], 1u);
        }
    }
    __syncthreads();
    unsigned *gh = shist + (size_t)mb * GSB_BINS;
    for (int b = threadIdx.x; b < GSB_BINS; b += blockDim.x) {
        const unsigned v = gsb_sh[b];
        if (v) atomicAdd(&gh[b], v);
    }
}

// ---- 2. plan: bands from the sample CDF (one CTA of 256 threads per minibatch) ----------------------------------
__global__ void __launch_bounds__(256) gsb_plan_kernel(const unsigned *shist, GsbPlan *plans) {
    extern __shared__ unsigned cum[];  // GSB_BINS inclusive prefix counts
    __shared__ unsigned wtot[8];
    const int mb = blockIdx.x, tid = threadIdx.x;
    const unsigned *gh = shist + (size_t)mb * GSB_BINS;
    const int per = GSB_BINS / 256;
    unsigned loc = 0;
    for (int q = tid * per; q < (tid + 1) * per; q++) { loc += gh[q]; cum[q] = loc; }
    unsigned incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(ADB_FULL, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) wtot[tid >> 5] = incl;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < (tid >> 5); w++) wbase += wtot[w];
    const unsigned add = wbase + incl - loc;
    for (int q = tid * per; q < (tid + 1) * per; q++) cum[q] += add;
    __syncthreads();
    if (tid != 0) return;
    GsbPlan P;
    memset(&P, 0, sizeof(P));
    const unsigned S = cum[GSB_BINS - 1];
    auto first_bin_above = [&](unsigned rank) {  // smallest b with cum[b] > rank
        int lo = 0, hi = GSB_BINS - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cum[mid] > rank) hi = mid; else lo = mid + 1; }
        return lo;
    };
    auto edge = [](int b) { return GSB_PA_MIN + (float)b / GSB_BIN_PER_PA; };  // exact in float32
    do {
        if (S < GSB_MIN_SAMPLE) break;
        const unsigned delta = (unsigned)ceilf(GSB_Z * 0.5f * sqrtf((float)S)) + 2u;
        if (S / 2 < delta + 1 || S / 2 + delta + 1 >= S) break;
        const unsigned r_lo = S / 2 - delta, r_hi = S / 2 + delta;
        const int b_lo = first_bin_above(r_lo), b_hi = first_bin_above(r_hi);
        if (b_lo <= 1 || b_hi >= GSB_BINS - 2) break;
        const float m_lo = edge(b_lo), m_hi = edge(b_hi + 1);
        // deviations around the bin edge cb: |x - edge(cb)| <= j bins  <=>  bins [cb - j, cb + j)
        const int cb = (b_lo + b_hi + 1) / 2;
        const int jmax = min(cb, GSB_BINS - cb) - 2;
        if (jmax < 8) break;
        auto count_r = [&](int j) -> unsigned { return j <= 0 ? 0u : cum[cb + j - 1] - cum[cb - j - 1]; };
        // j_lo: largest j with count(j + 1) <= r_lo ; j_hi: smallest j with count(j) > r_hi
        int lo = 0, hi = jmax;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (count_r(mid + 1) <= r_lo) lo = mid; else hi = mid - 1; }
        const int j_lo = lo;
        if (count_r(j_lo + 1) > r_lo) break;
        lo = 1; hi = jmax;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (count_r(mid) > r_hi) hi = mid; else lo = mid + 1; }
        const int j_hi = lo;
        if (!(count_r(j_hi) > r_hi) || j_hi + 2 >= jmax) break;
        const float c = edge(cb), bw = 1.0f / GSB_BIN_PER_PA;
        const float wc = fmaxf(c - m_lo, m_hi - c);          // |c - med| <= wc for every med in the value band
        float d_lo = (float)j_lo * bw - wc, d_hi = (float)(j_hi + 1) * bw + wc;
        d_lo *= (1.0f - 8e-6f);
        d_hi *= (1.0f + 8e-6f);
        if (!(d_lo > 2.0f * (m_hi - m_lo) + 4.0f * bw)) break;  // the value band must sit well inside "inside"
        const float mg = 8e-6f;
        P.m_lo = m_lo; P.m_hi = m_hi; P.d_lo = d_lo; P.d_hi = d_hi;
        // inside: |x - med| < d_lo for every med in [m_lo, m_hi]  (shrunk by a relative margin >> 1 ulp)
        P.xI_lo = (m_hi - d_lo) + fabsf(m_hi - d_lo) * mg + 1e-6f * d_lo;
        P.xI_hi = (m_lo + d_lo) - fabsf(m_lo + d_lo) * mg - 1e-6f * d_lo;
        // beyond the bands: |x - med| > d_hi for every med in [m_lo, m_hi]  (grown by the same margin)
        P.xL_lo = (m_lo - d_hi) - fabsf(m_lo - d_hi) * mg - 1e-6f * d_hi;
        P.xR_hi = (m_hi + d_hi) + fabsf(m_hi + d_hi) * mg + 1e-6f * d_hi;
        if (!(P.xL_lo < P.xI_lo && P.xI_lo < m_lo && m_hi < P.xI_hi && P.xI_hi < P.xR_hi)) break;
        P.use = 1;
    } while (0);
    plans[mb] = P;
}

// ---- 3. the streaming pass ---------------------------------------------------------------------------------------
struct GsbRead { int cM, cL, cR, ok; };  // first code of the three tally ranges of a read

__global__ void __launch_bounds__(256) gsb_pass_kernel(BatchDev B, int max_obs_trace, GsbPlan *plans, unsigned *tab,
                                                       GsbRead *bases) {
    __shared__ int th[8];
    __shared__ unsigned long long red[3][8];
    const int mb = blockIdx.y;
    GsbPlan *P = &plans[mb];
    if (!P->use) return;
    const int r0 = mb * B.batch_size, r1 = min(r0 + B.batch_size, B.n_reads);
    const float m_lo = P->m_lo, m_hi = P->m_hi, xL_lo = P->xL_lo, xI_lo = P->xI_lo, xI_hi = P->xI_hi, xR_hi = P->xR_hi;
    unsigned long long tot = 0;
    unsigned below = 0, inside = 0;
    for (int r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
        const ReadSrc src = make_src(B, r);
        const int n = min(src.n, max_obs_trace);
        __syncthreads();
        if (threadIdx.x < 6) {
            int ok = (src.i16 != nullptr) && (src.cscale > 0.0f) && isfinite(src.cscale) && isfinite(src.coff);
            int c = 0;
            if (ok) {
                switch (threadIdx.x) {
                    case 0: c = gsb_code_at(m_lo, false, src.coff, src.cscale, &ok); break;        // cM_lo
                    case 1: c = gsb_code_at(m_hi, true, src.coff, src.cscale, &ok) - 1; break;     // cM_hi
                    case 2: c = gsb_code_at(xI_lo, false, src.coff, src.cscale, &ok); break;       // cI_lo
                    case 3: c = gsb_code_at(xI_hi, true, src.coff, src.cscale, &ok) - 1; break;    // cI_hi
                    case 4: c = gsb_code_at(xL_lo, false, src.coff, src.cscale, &ok); break;       // cL_lo
                    default: c = gsb_code_at(xR_hi, true, src.coff, src.cscale, &ok) - 1; break;   // cR_hi
                }
            }
            th[threadIdx.x] = c;
            if (!ok) P->fallback = 1;
        }
        __syncthreads();
        // clip to the codes an int16 can hold (lower ends to [-32768, 32768], upper ends to [-32769, 32767])
        const int cM_lo = min(max(th[0], -32768), 32768), cM_hi = min(max(th[1], -32769), 32767);
        const int cI_lo = min(max(th[2], -32768), 32768), cI_hi = min(max(th[3], -32769), 32767);
        const int cL_lo = min(max(th[4], -32768), 32768), cR_hi = min(max(th[5], -32769), 32767);
        // sizes of the code ranges (an empty range has size 0); the tally ranges must fit the table
        const int nM = max(cM_hi - cM_lo + 1, 0), nI = max(cI_hi - cI_lo + 1, 0);
        const int nL = cI_lo - cL_lo, nR = cR_hi - cI_hi;
        const bool fits = (nM <= GSB_TAB_M) && (nL >= 0) && (nL <= GSB_TAB_L) && (nR >= 0) && (nR <= GSB_TAB_R) &&
                          (cI_lo <= cM_lo) && (cM_hi <= cI_hi) && (nI > 0);
        if (threadIdx.x == 0) {
            GsbRead br;
            br.cM = cM_lo; br.cL = cL_lo; br.cR = cI_hi + 1; br.ok = fits ? 1 : 0;
            bases[r] = br;
            if (!fits) P->fallback = 1;
            tot += (unsigned long long)max(n, 0);
        }
        if (!fits || n <= 0) continue;
        unsigned *row = tab + (size_t)r * GSB_TAB;
        const unsigned uI = (unsigned)nI, uA = (unsigned)(cR_hi - cL_lo + 1), uM = (unsigned)nM;
        auto consume = [&](int c) {
            below += (c < cM_lo);
            const bool inI = (unsigned)(c - cI_lo) < uI;
            inside += inI;
            const bool inM = (unsigned)(c - cM_lo) < uM;
            if ((unsigned)(c - cL_lo) < uA && (!inI || inM)) {
                int idx;
                if (inI) idx = c - cM_lo;
                else if (c < cI_lo) idx = GSB_TAB_M + (c - cL_lo);
                else idx = GSB_TAB_M + GSB_TAB_L + (c - cI_hi - 1);
                atomicAdd(&row[idx], 1u);
            }
        };
        const int16_t *p = src.i16;
        const int head = min(n, (int)(((16 - ((uintptr_t)p & 15)) & 15) >> 1));
        const int nvec = (n - head) >> 3;
        const uint4 *pv = (const uint4 *)(p + head);
        for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
            const uint4 q = __ldg(pv + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                consume((int)(int16_t)(w[t] & 0xffffu));
                consume((int)(int16_t)(w[t] >> 16));
            }
        }
        const int tail0 = head + (nvec << 3);
        for (int j = threadIdx.x; j < head; j += blockDim.x) consume((int)p[j]);
        for (int j = tail0 + threadIdx.x; j < n; j += blockDim.x) consume((int)p[j]);
    }
    // block totals -> plan counters
    unsigned long long b64 = below, i64 = inside;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b64 += __shfl_xor_sync(ADB_FULL, b64, o);
        i64 += __shfl_xor_sync(ADB_FULL, i64, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = b64; red[1][threadIdx.x >> 5] = i64; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long sb = 0, si = 0;
        for (int w = 0; w < 8; w++) { sb += red[0][w]; si += red[1][w]; }
        atomicAdd(&P->c_below, sb);
        atomicAdd(&P->c_inside, si);
        atomicAdd(&P->n_total, tot);
    }
}

// ---- 4. finish: exact weighted selection among the tallies (one CTA of 256 threads per minibatch) ---------------
// Two ranks (ascending) among items i in [0, n_items) with key(i) (32-bit, order preserving) and weight(i).
// 3-pass radix select (11 + 11 + 10 bits) with weighted shared-memory histograms.  Returns false if a rank is not
// covered by the total weight.  All threads call; out[] valid for all threads on return.
template <class ItemF>
__device__ bool cta_wselect2(ItemF item, int n_items, unsigned long long rank0, unsigned long long rank1, uint32_t out[2],
                             unsigned *hist /*[2][2048]*/, unsigned long long *part /*[256]*/, int *ibuf /*[8]*/) {
    const int tid = threadIdx.x, T = blockDim.x;
    uint32_t prefix[2] = {0, 0};
    unsigned long long rem[2] = {rank0, rank1};
    for (int pass = 0; pass < 3; pass++) {
        const int shift = gsel_shift(pass);
        const uint32_t mask = gsel_mask(pass);
        const int pshift = (pass == 0) ? 32 : (pass == 1 ? 21 : 10);
        const bool same = prefix[0] == prefix[1];
        __syncthreads();
        for (int b = tid; b < 2 * GSEL_BINS; b += T) hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < n_items; i += T) {
            uint32_t k; unsigned w;
            item(i, k, w);
            if (!w) continue;
            const uint32_t hi = (pshift >= 32) ? 0u : (k >> pshift);
            const int bin = (int)((k >> shift) & mask);
            if (hi == prefix[0]) atomicAdd(&hist[bin], w);
            if (!same && hi == prefix[1]) atomicAdd(&hist[GSEL_BINS + bin], w);
        }
        __syncthreads();
        for (int t = 0; t < 2; t++) {
            const unsigned *h = hist + ((same ? 0 : t) * GSEL_BINS);
            unsigned long long loc = 0;
            for (int b = tid * 8; b < tid * 8 + 8; b++) loc += h[b];
            part[tid] = loc;
            __syncthreads();
            if (tid == 0) {
                unsigned long long acc = 0;
                int fb = -1;
                for (int q = 0; q < 256 && fb < 0; q++) {
                    if (rem[t] < acc + part[q]) {
                        for (int b = q * 8; b < q * 8 + 8; b++) {
                            if (rem[t] < acc + h[b]) { fb = b; break; }
                            acc += h[b];
                        }
                    } else acc += part[q];
                }
                ibuf[2 * t] = fb;
                part[0] = acc;  // weight before the bin (read back below)
            }
            __syncthreads();
            const int fb = ibuf[2 * t];
            const unsigned long long before = part[0];
            __syncthreads();
            if (fb < 0) return false;
            prefix[t] = (prefix[t] << ((pass == 2) ? 10 : 11)) | (uint32_t)fb;
            rem[t] -= before;
        }
    }
    out[0] = prefix[0];
    out[1] = prefix[1];
    return true;
}

__global__ void __launch_bounds__(256) gsb_finish_kernel(BatchDev B, GsbPlan *plans, const unsigned *tab,
                                                         const GsbRead *bases, GselState *states, int *active) {
    __shared__ unsigned hist[2 * GSEL_BINS];
    __shared__ unsigned long long part[256];
    __shared__ int ibuf[8];
    __shared__ unsigned long long cnt[2];
    const int mb = blockIdx.x, tid = threadIdx.x;
    const GsbPlan P = plans[mb];
    const int r0 = mb * B.batch_size, r1 = min(r0 + B.batch_size, B.n_reads), nr = r1 - r0;
    bool ok = P.use && !P.fallback && P.n_total > 0;
    const unsigned long long n = P.n_total, k0 = n ? (n - 1) / 2 : 0, k1 = n / 2;
    if (ok) ok = (k0 >= P.c_below);
    float med = 0.f, mad = 0.f;
    if (ok) {
        auto itemM = [&](int i, uint32_t &k, unsigned &w) {
            const int r = r0 + i / GSB_TAB_M, e = i % GSB_TAB_M;
            w = tab[(size_t)r * GSB_TAB + e];
            k = w ? f32_key(gsb_pa(bases[r].cM + e, B.calib_offset[r], B.calib_scale[r])) : 0u;
        };
        uint32_t key[2];
        ok = cta_wselect2(itemM, nr * GSB_TAB_M, k0 - P.c_below, k1 - P.c_below, key, hist, part, ibuf);
        if (ok) {
            const float a = key_f32(key[0]), b = key_f32(key[1]);
            med = (n & 1ull) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
            ok = (med >= P.m_lo && med <= P.m_hi);
        }
    }
    if (ok) {
        // deviations of the tallied codes of the two outer bands; the "inside" samples all deviate less than d_lo
        auto itemD = [&](int i, uint32_t &k, unsigned &w) {
            const int r = r0 + i / (GSB_TAB_L + GSB_TAB_R), e = i % (GSB_TAB_L + GSB_TAB_R);
            w = tab[(size_t)r * GSB_TAB + GSB_TAB_M + e];
            if (!w) { k = 0; return; }
            const int c = (e < GSB_TAB_L) ? bases[r].cL + e : bases[r].cR + (e - GSB_TAB_L);
            k = __float_as_uint(fabsf(__fsub_rn(gsb_pa(c, B.calib_offset[r], B.calib_scale[r]), med)));
        };
        if (tid == 0) { cnt[0] = 0; cnt[1] = 0; }
        __syncthreads();
        unsigned long long c_le = 0, c_lt = 0;
        for (int i = tid; i < nr * (GSB_TAB_L + GSB_TAB_R); i += blockDim.x) {
            uint32_t k; unsigned w;
            itemD(i, k, w);
            if (!w) continue;
            const float d = __uint_as_float(k);
            if (d <= P.d_lo) c_le += w;
            if (d < P.d_hi) c_lt += w;
        }
        atomicAdd(&cnt[0], c_le);
        atomicAdd(&cnt[1], c_lt);
        __syncthreads();
        ok = (P.c_inside + cnt[0] <= k0) && (k1 < P.c_inside + cnt[1]);
        if (ok) {
            uint32_t key[2];
            ok = cta_wselect2(itemD, nr * (GSB_TAB_L + GSB_TAB_R), k0 - P.c_inside, k1 - P.c_inside, key, hist, part, ibuf);
            if (ok) {
                const float a = __uint_as_float(key[0]), b = __uint_as_float(key[1]);
                mad = (n & 1ull) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
                ok = (a > P.d_lo && b < P.d_hi);
            }
        }
    }
    if (tid == 0) {
        active[mb] = ok ? 0 : 1;
        if (ok) {
            GselState *st = &states[mb];
            st->count = n;
            st->med = med;
            st->mad = mad;
            st->status = (mad == 0.0f) ? ADB_ERR_MAD_ZERO : ADB_OK;
        }
    }
}
