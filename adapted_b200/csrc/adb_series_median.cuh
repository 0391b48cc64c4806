// Medians of the two moving-statistics series of every read (mvs.py:119-126: np.median of bottleneck's move_var /
// move_mean over the poly(A) candidate), one WARP per series.
//
// The series are the rows mvs_series_kernel wrote to the pools.  A warp stages the ordered keys of its row in its own
// slice of shared memory (one coalesced read of the row from HBM / L2), then bisects on the key value with warp-level
// counts only -- no CTA barrier, no atomics -- until at most SM_NCAND keys are left inside the bracket; those are
// compacted by ballot and ranked directly.  Rows longer than the slice (poly(A) candidates beyond ~3000 samples) are
// bracketed from a subsample first, so that one more pass over the row stages only the keys around the middle ranks.
//
// This used to be a CTA-wide phase of validate_fast_kernel (vf_series_medians: ~15 % of that kernel's warp samples,
// a barrier per pass, the keys parked in the window memory, and a second kernel instantiation for rows that did not fit
// there); as a kernel of its own every warp runs the same small loop.
#pragma once
#include "adb_common.cuh"
#include "adb_validate.cuh"

#define SM_WARPS 4
#define SM_CAP 3072      // keys per warp slice (12 KB)
#define SM_NCAND 64
#define SM_NSAMP 2048    // largest subsample of a row longer than the slice

struct SeriesMedianArgs {
    BatchDev B;
    const float *pre_var, *pre_mean;
    const long long *pre_off;   // [n_reads] row offset into both pools, -1: no row
    const int *pre_meta;        // [n_reads][2] = (adapter_end, polya_end) of the row
    const int *perm;            // optional: reads in length-sorted order (neighbouring warps finish together)
    float *out;                 // [n_reads][2] = median of the moving variance, of the moving mean (NaN: series empty)
    int n_reads;
};

__host__ __device__ inline size_t series_median_smem_bytes() { return (size_t)SM_WARPS * (SM_CAP + SM_NCAND + 4) * 4; }

// Keys of a series: staged in the warp's slice (ordered uint32) or formed on the fly from the pool row.
struct SmKeys {
    const uint4 *K4;       // staged keys (nullptr: read the row)
    const float4 *G4;
    __device__ __forceinline__ uint4 at(int v) const {
        if (K4) return K4[v];
        const float4 x = __ldg(G4 + v);
        return make_uint4(f32_key(x.x), f32_key(x.y), f32_key(x.z), f32_key(x.w));
    }
    __device__ __forceinline__ uint32_t at1(int j) const {
        if (K4) return reinterpret_cast<const uint32_t *>(K4)[j];
        return f32_key(__ldg(reinterpret_cast<const float *>(G4) + j));
    }
};

// 32 keys, one per lane -> ascending over the lanes (bitonic network, 15 exchange steps)
__device__ __forceinline__ uint32_t warp_sort32(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(ADB_FULL, v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? min(v, o) : max(v, o);
        }
    }
    return v;
}

// The keys at ranks `rank` and `rank + 1` (0-based) of n keys (n >= 1, rank < n): nv full vectors through K plus the
// lane's tail key kt (has_t).  Bisection on the key value from [mn, mx] with warp-level counts until at most SM_NCAND
// keys are left inside the bracket; those are compacted by ballot into `cand` and ranked directly.  Returns false in
// have_b if rank + 1 == n.
__device__ void warp_select2(const SmKeys &K, int nv, uint32_t kt, bool has_t, int n, unsigned rank, uint32_t mn,
                             uint32_t mx, uint32_t *cand, uint32_t &a, uint32_t &b, bool &have_b) {
    const int lane = threadIdx.x & 31;
    // keys in [lo, hi] hold the ranks cnt_lo .. cnt_hi - 1; the rank looked for stays inside
    uint32_t lo = mn, hi = mx;
    unsigned cnt_lo = 0, cnt_hi = (unsigned)n;
    if (n >= 512 && lo < hi) {
        // Opening bracket from 32 keys spread over the row: the sorted sample keys 8 positions either side of the rank's
        // quantile, both counted in one pass.  Bisecting the key VALUE from [min, max] spends its first passes on the
        // tails (moving variances span decades); the sample bracket starts where the keys are dense.  Whatever the
        // counts say is a valid bracket (the rank is inside, below or above), so nothing has to be redone.
        const int nk = (nv << 2);  // sampled from the full vectors
        const uint32_t sk = warp_sort32(K.at1((int)(((long long)(2 * lane + 1) * nk) >> 6)));
        const int c = (int)(((long long)rank * 32) / n);
        const int i_lo = c - 8, i_hi = c + 8;
        const uint32_t s_lo = __shfl_sync(ADB_FULL, sk, max(i_lo, 0)), s_hi = __shfl_sync(ADB_FULL, sk, min(i_hi, 31));
        const bool use_lo = i_lo >= 0 && s_lo > lo, use_hi = i_hi <= 31 && s_hi < hi;
        if (use_lo || use_hi) {
            const uint32_t tl = use_lo ? s_lo - 1 : 0u, th = use_hi ? s_hi : 0xffffffffu;  // count keys <= tl, keys <= th
            unsigned cl = 0, ch = 0;
            for (int v = lane; v < nv; v += 32) {
                const uint4 x = K.at(v);
                cl += (x.x <= tl) + (x.y <= tl) + (x.z <= tl) + (x.w <= tl);
                ch += (x.x <= th) + (x.y <= th) + (x.z <= th) + (x.w <= th);
            }
            if (has_t) { cl += (kt <= tl); ch += (kt <= th); }
            cl = use_lo ? __reduce_add_sync(ADB_FULL, cl) : 0u;
            ch = use_hi ? __reduce_add_sync(ADB_FULL, ch) : (unsigned)n;
            if (use_lo && rank < cl) { hi = tl; cnt_hi = cl; }
            else if (use_hi && rank >= ch) { lo = th + 1; cnt_lo = ch; }
            else {
                if (use_lo) { lo = s_lo; cnt_lo = cl; }
                if (use_hi) { hi = s_hi; cnt_hi = ch; }
            }
        }
    }
    while (lo < hi && cnt_hi - cnt_lo > SM_NCAND) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        unsigned c0 = 0, c1 = 0;
        int v = lane;
        for (; v + 32 < nv; v += 64) {
            const uint4 x = K.at(v), y = K.at(v + 32);
            c0 += (x.x <= mid) + (x.y <= mid) + (x.z <= mid) + (x.w <= mid);
            c1 += (y.x <= mid) + (y.y <= mid) + (y.z <= mid) + (y.w <= mid);
        }
        if (v < nv) { const uint4 x = K.at(v); c0 += (x.x <= mid) + (x.y <= mid) + (x.z <= mid) + (x.w <= mid); }
        if (has_t) c1 += (kt <= mid);
        const unsigned c = __reduce_add_sync(ADB_FULL, c0 + c1);
        if (c > rank) { hi = mid; cnt_hi = c; } else { lo = mid + 1; cnt_lo = c; }
    }
    b = 0xffffffffu;
    if (lo == hi) {  // every key left equals hi
        a = hi;
        b = hi;
        have_b = cnt_hi > rank + 1;
    } else {
        // compact the cnt_hi - cnt_lo (<= SM_NCAND) keys of the bracket, rank them directly
        const unsigned lt = (1u << lane) - 1u;
        int base = 0;
        auto put = [&](uint32_t k, bool valid) {
            const bool in = valid && k >= lo && k <= hi;
            const unsigned m = __ballot_sync(ADB_FULL, in);
            if (in) cand[base + __popc(m & lt)] = k;
            base += __popc(m);
        };
        for (int v0 = 0; v0 < nv; v0 += 32) {
            const int v = v0 + lane;
            const bool ok = v < nv;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (ok) x = K.at(v);
            put(x.x, ok); put(x.y, ok); put(x.z, ok); put(x.w, ok);
        }
        put(kt, has_t);
        __syncwarp();
        const int m = base;  // == cnt_hi - cnt_lo
        const int r = (int)(rank - cnt_lo);
        for (int i = lane; i < m; i += 32) {
            const uint32_t ki = cand[i];
            int pos = 0;
            for (int j = 0; j < m; j++) { const uint32_t kj = cand[j]; pos += (kj < ki) || (kj == ki && j < i); }
            if (pos == r) cand[SM_NCAND] = ki;
            if (pos == r + 1) cand[SM_NCAND + 1] = ki;
        }
        __syncwarp();
        a = cand[SM_NCAND];
        have_b = (r + 1 < m);
        if (have_b) b = cand[SM_NCAND + 1];
        __syncwarp();
    }
    if (!have_b && rank + 1 < (unsigned)n) {  // the upper neighbour lies beyond the bracket: the smallest key above hi (rare)
        uint32_t best = 0xffffffffu;
        for (int v = lane; v < nv; v += 32) {
            const uint4 x = K.at(v);
            if (x.x > hi) best = min(best, x.x);
            if (x.y > hi) best = min(best, x.y);
            if (x.z > hi) best = min(best, x.z);
            if (x.w > hi) best = min(best, x.w);
        }
        if (has_t && kt > hi) best = min(best, kt);
        b = __reduce_min_sync(ADB_FULL, best);
        have_b = true;
    }
}

// numpy median of the n NaN-free floats at g (16-byte aligned).  kbuf: SM_CAP + SM_NCAND + 4 words of this warp.
__device__ float warp_series_median(const float *__restrict__ g, int n, uint32_t *kbuf) {
    const int lane = threadIdx.x & 31;
    if (n <= 0) return CUDART_NAN_F;
    uint32_t *cand = kbuf + SM_CAP;
    const float4 *G4 = reinterpret_cast<const float4 *>(g);
    uint4 *K4 = reinterpret_cast<uint4 *>(kbuf);
    const unsigned rank = (unsigned)(n - 1) >> 1;
    const bool need_b = (n & 1) == 0;
    uint32_t a = 0, b = 0;
    bool have_b = false;
    if (n <= SM_CAP) {
        // stage the keys (one coalesced read of the row)
        const int nv = n >> 2;
        uint32_t mn = 0xffffffffu, mx = 0u;
        for (int v = lane; v < nv; v += 32) {
            const float4 x = __ldg(G4 + v);
            const uint4 k = make_uint4(f32_key(x.x), f32_key(x.y), f32_key(x.z), f32_key(x.w));
            K4[v] = k;
            mn = min(min(mn, k.x), min(k.y, min(k.z, k.w)));
            mx = max(max(mx, k.x), max(k.y, max(k.z, k.w)));
        }
        const int jt = (nv << 2) + lane;  // tail key of this lane (n & 3 of them)
        uint32_t kt = 0xffffffffu;
        const bool has_t = jt < n;
        if (has_t) { kt = f32_key(__ldg(g + jt)); mn = min(mn, kt); mx = max(mx, kt); }
        mn = __reduce_min_sync(ADB_FULL, mn);
        mx = __reduce_max_sync(ADB_FULL, mx);
        __syncwarp();
        const SmKeys K{K4, G4};
        warp_select2(K, nv, kt, has_t, n, rank, mn, mx, cand, a, b, have_b);
    } else {
        // Long row (poly(A) candidates beyond SM_CAP samples): a subsample of SM_NSAMP keys brackets the middle ranks,
        // ONE pass over the row counts the keys below the bracket and stages those inside, the slice finishes.
        // Verified from the exact counts; a bracket that misses (or overflows the slice) falls back to bisecting the
        // whole row from the pool.
        bool done = false;
        {
            const int nv_all = n >> 2;
            const int ns = min(nv_all, SM_NSAMP);
            uint32_t smn = 0xffffffffu, smx = 0u;
            if (nv_all <= SM_NSAMP) {
                // rows up to 4 SM_NSAMP keys: the first key of every vector, read in one COALESCED pass over the row (which
                // also brings it into L2 for the pass that follows)
                for (int v = lane; v < nv_all; v += 32) {
                    const uint32_t k = f32_key(__ldg(G4 + v).x);
                    kbuf[v] = k;
                    smn = min(smn, k); smx = max(smx, k);
                }
            } else {
                // longer rows: SM_NSAMP keys spread evenly (one sector each: less traffic than a whole pass)
                for (int i = lane; i < SM_NSAMP; i += 32) {
                    const uint32_t k = f32_key(__ldg(g + (size_t)(((long long)i * n) / SM_NSAMP)));
                    kbuf[i] = k;
                    smn = min(smn, k); smx = max(smx, k);
                }
            }
            const int nsp = (ns + 3) & ~3;                           // padded to full vectors with keys above every rank
            if (lane < nsp - ns) kbuf[ns + lane] = 0xffffffffu;
            smn = __reduce_min_sync(ADB_FULL, smn);
            smx = __reduce_max_sync(ADB_FULL, smx);
            __syncwarp();
            const SmKeys KS{K4, G4};
            // half-width of the bracket in sample ranks: 3.5 sigma of the middle rank of the subsample (+ 2); a miss is
            // caught by the exact counts below and the row redone
            const int spread = (int)(1.75f * sqrtf((float)ns)) + 2;
            const int rs = (int)(((long long)rank * ns) / n);
            const int r_lo = rs - spread, r_hi = rs + spread + 1;
            uint32_t s_lo = 0u, s_hi = 0xffffffffu, t0, t1;
            bool hb;
            if (r_lo >= 0) warp_select2(KS, nsp >> 2, 0, false, nsp, (unsigned)r_lo, smn, smx, cand, s_lo, t0, hb);
            if (r_hi < ns) warp_select2(KS, nsp >> 2, 0, false, nsp, (unsigned)r_hi, smn, smx, cand, s_hi, t1, hb);
            __syncwarp();
            // one pass: keys below s_lo are counted, keys in [s_lo, s_hi] staged
            const int nv = n >> 2;
            const unsigned lt = (1u << lane) - 1u;
            unsigned below = 0;
            int m = 0;
            bool overflow = false;
            auto put = [&](uint32_t k, bool valid) {
                below += (valid && k < s_lo);
                const bool in = valid && k >= s_lo && k <= s_hi;
                const unsigned bm = __ballot_sync(ADB_FULL, in);
                const int pos = m + __popc(bm & lt);
                if (in && pos < SM_CAP) kbuf[pos] = k;
                m += __popc(bm);
            };
            for (int v0 = 0; v0 < nv && !overflow; v0 += 32) {
                const int v = v0 + lane;
                const bool ok = v < nv;
                uint4 x = make_uint4(0, 0, 0, 0);
                if (ok) { const float4 f = __ldg(G4 + v); x = make_uint4(f32_key(f.x), f32_key(f.y), f32_key(f.z), f32_key(f.w)); }
                put(x.x, ok); put(x.y, ok); put(x.z, ok); put(x.w, ok);
                overflow = m > SM_CAP - 4;
            }
            const int jt = (nv << 2) + lane;
            put(jt < n ? f32_key(__ldg(g + jt)) : 0u, jt < n);
            below = __reduce_add_sync(ADB_FULL, below);
            __syncwarp();
            // the ranks wanted must lie inside the staged keys
            const unsigned top = need_b ? rank + 1 : rank;
            if (!overflow && m <= SM_CAP - 4 && below <= rank && top < below + (unsigned)m) {
                // pad to full vectors with keys above everything staged (they never enter a bracket below rank m)
                const int mp = (m + 3) & ~3;
                if (lane < mp - m) kbuf[m + lane] = 0xffffffffu;
                __syncwarp();
                const SmKeys KM{K4, G4};
                warp_select2(KM, mp >> 2, 0, false, mp, rank - below, s_lo, s_hi, cand, a, b, have_b);
                done = true;
            }
        }
        if (!done) {
            const int nv = n >> 2;
            uint32_t mn = 0xffffffffu, mx = 0u;
            const SmKeys KG{nullptr, G4};
            for (int v = lane; v < nv; v += 32) {
                const uint4 k = KG.at(v);
                mn = min(min(mn, k.x), min(k.y, min(k.z, k.w)));
                mx = max(max(mx, k.x), max(k.y, max(k.z, k.w)));
            }
            const int jt = (nv << 2) + lane;
            uint32_t kt = 0xffffffffu;
            const bool has_t = jt < n;
            if (has_t) { kt = f32_key(__ldg(g + jt)); mn = min(mn, kt); mx = max(mx, kt); }
            mn = __reduce_min_sync(ADB_FULL, mn);
            mx = __reduce_max_sync(ADB_FULL, mx);
            warp_select2(KG, nv, kt, has_t, n, rank, mn, mx, cand, a, b, have_b);
        }
    }
    if (!need_b) return key_f32(a);
    return __fdiv_rn(__fadd_rn(key_f32(a), key_f32(b)), 2.0f);
}

// lengths of the two series of a row (mvs.py:100-107 via mvs_plan: a row exists only if at least one window fits)
__device__ __forceinline__ void series_lengths(const adb_config &cfg, int size, int ae, int pe, int &nv, int &nm) {
    int a = ae, b = pe;
    clip_seg(a, b, size);
    const int L = b - a;
    const bool win_var = !(pe - ae <= cfg.pA_var_window + 2), win_mean = !(pe - ae <= cfg.pA_mean_window + 2);
    nv = win_var ? max(L - (cfg.pA_var_window - 1), 0) : 0;
    nm = win_mean ? max(L - (cfg.pA_mean_window - 1), 0) : 0;
}

__global__ void __launch_bounds__(SM_WARPS * 32) series_median_kernel(SeriesMedianArgs A, adb_config cfg) {
    extern __shared__ __align__(16) unsigned char sm_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *kbuf = reinterpret_cast<uint32_t *>(sm_smem) + (size_t)warp * (SM_CAP + SM_NCAND + 4);
    const int n_items = 2 * A.n_reads;
    for (int item = blockIdx.x * SM_WARPS + warp; item < n_items; item += gridDim.x * SM_WARPS) {
        const int q = item >> 1, which = item & 1;
        const int r = A.perm ? A.perm[q] : q;
        const long long po = A.pre_off[r];
        if (po < 0) continue;
        const ReadSrc src = make_src(A.B, r);
        int nv, nm;
        series_lengths(cfg, src.n, A.pre_meta[2 * r], A.pre_meta[2 * r + 1], nv, nm);
        const float med = which ? warp_series_median(A.pre_mean + po, nm, kbuf) : warp_series_median(A.pre_var + po, nv, kbuf);
        if (lane == 0) A.out[2 * r + which] = med;
        __syncwarp();
    }
}
