// validate_boundaries + per-segment statistics for int16 sources from HISTOGRAMS BUILT ON THE TENSOR CORES.
//
// Reference: adapted/detect/combined.py:358-631 (control flow and checks as in adb_vfast.cuh / adb_validate.cuh).
//
// Every order statistic the reference asks for (medians, MADs, p15 / p85 of up to a dozen sample ranges of a read) is
// a function of the histogram of the ADC codes of the range.  Shared-memory atomics build such a histogram at 2 cycles
// per sample and SM (profiles/r1_validate_kernel_v3.txt), counting passes need ~20 passes over the window
// (adb_vfast.cuh).  Here the tensor core does the scatter-add: for a batch of K = 32 consecutive samples a warp writes
// two ONE-HOT operand tiles in shared memory,
//     A[hi][k] = 1 iff (code_k - base) >> 4 == hi      (M = 128 rows)
//     B[lo][k] = 1 iff (code_k - base) & 15 == lo      (N = 16 rows)
// (one byte store per sample and tile, e4m3 1.0 = 0x38) and one tcgen05.mma.kind::f8f6f4 accumulates
//     D[hi][lo] += sum_k A[hi][k] * B[lo][k]  =  number of samples of the batch with code - base == 16 * hi + lo
// into a float32 accumulator in TMEM (exact: counts stay below 2^24) -- a 2048-bin histogram per accumulator, 32
// samples per MMA, no atomics, no conflicts between samples of equal value.  After the MMA has read the tiles the warp
// clears the bytes it set (the tiles are all-zero between batches).
//
// The ranges of a read overlap (adapter, its tail windows, poly(A), the windows around adapter_end, the rest), so the
// window is cut at every range boundary into at most VH_MAX_PIECES disjoint PIECES, one accumulator each (16 TMEM
// columns); a range is a run of consecutive pieces.  The samples are read ONCE from global memory (no staging of the
// window in shared memory).  The accumulators are then copied to shared memory as per-piece cumulative counts and
// every statistic is a handful of binary searches in them (lanes of warp 0, one query each): ranks by bisection on the
// code, MADs by the candidate halving of adb_vfast.cuh with the counts looked up instead of counted.
//
// Codes outside [base, base + 2047] (base = the code of -20 pA: the range covers -20 .. +339 pA at a typical
// calibration) fall into the end bins; a read is handed to validate_kernel (same results) whenever an answer or a probe
// touches an end bin that holds such codes.  Also handed over: what validate_fast_kernel hands over.
#pragma once
#include "adb_cnn_tc.cuh"
#include "adb_vfast.cuh"

#define VH_BINS 2048
#define VH_MAX_PIECES 8
#define VH_A_LBO 2112                      // A tile: 2 K-groups (16 samples each) of 128 rows x 16 B, 64 B apart in banks
#define VH_A_BYTES (VH_A_LBO + 2048)
#define VH_B_LBO 320                       // B tile: 2 K-groups of 16 rows x 16 B
#define VH_B_BYTES (VH_B_LBO + 256)
#define VH_TILE_BYTES 4864                 // A + B of one warp, rounded to 128 B
#define VH_WARPS (VF_THREADS / 32)
#define VH_ARENA (VH_WARPS * VH_TILE_BYTES)  // 38 912 B >= VH_MAX_PIECES * VH_BINS * 2 (cumulative counts, u16)
// instruction descriptor: D = f32, A = B = e4m3 (0), K-major, N = 16, M = 128
#define VH_IDESC ((1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24))
#define VH_NQ 12

__host__ __device__ inline size_t vhist_smem_bytes() { return (size_t)VH_ARENA + 128; }

__device__ __forceinline__ void vh_mma_f8(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(VH_IDESC)
        : "memory");
}

__device__ __forceinline__ void vh_tmem_zero32(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr),
        "r"(z)
        : "memory");
}

__device__ __forceinline__ void vh_tmem_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct VhShared {
    VfScratch S;                       // scratch of the helpers shared with adb_vfast.cuh
    uint64_t bar[VH_WARPS];            // "the MMA of this warp's batch has read the tiles"
    uint32_t tmem_slot;
    int cuts[VH_MAX_PIECES + 4];       // sorted cut points; piece p = samples [cuts[p], cuts[p + 1])
    int bstart[VH_MAX_PIECES + 2];     // first batch (32 samples) of piece p in the flattened batch space
    int np;
    int n_low, n_high;                 // samples of the read below base / above base + 2047
    int unsettled;                     // an answer or a probe touched an end bin holding such samples
    int qv0[VH_NQ], qv1[VH_NQ], qn[VH_NQ];  // rank queries: codes at rank k and k + 1, samples of the range
    float mad[4];
};

// samples of the pieces [p0, p1) whose clamped code offset is <= x
__device__ __forceinline__ int vh_cle(const uint16_t *cum, int p0, int p1, int x) {
    if (x < 0) return 0;
    x = min(x, VH_BINS - 1);
    int c = 0;
    for (int p = p0; p < p1; p++) c += (int)cum[p * VH_BINS + x];
    return c;
}
// smallest offset x with vh_cle(x) > k (0 <= k < samples of the range)
__device__ __forceinline__ int vh_select(const uint16_t *cum, int p0, int p1, int k) {
    int lo = 0, hi = VH_BINS - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (vh_cle(cum, p0, p1, mid) > k) hi = mid; else lo = mid + 1;
    }
    return lo;
}

struct VhRange { int p0, p1; };  // pieces of a sample range
__device__ __forceinline__ VhRange vh_range(const VhShared &H, int a, int b, int size) {
    clip_seg(a, b, size);
    VhRange q{0, 0};
    if (b <= a) return q;
    const int np = H.np;
    for (int i = 0; i <= np; i++) {
        if (H.cuts[i] == a) q.p0 = i;
        if (H.cuts[i] == b) q.p1 = i;
    }
    return q;
}

// median(|x - med|) of a range from the cumulative counts: the candidate halving of vf_run (adb_vfast.cuh) with the
// counts looked up.  One thread.
__device__ float vh_mad(const VfRead &R, VhShared &H, const uint16_t *cum, VhRange q, int base, int n, float med) {
    if (n <= 0 || !(med == med)) return CUDART_NAN_F;
    const int n_low = H.n_low, n_high = H.n_high;
    auto touch = [&](int x) { if ((x <= 0 && n_low > 0) || (x >= VH_BINS - 1 && n_high > 0)) H.unsettled = 1; };
    const int xmin = vh_select(cum, q.p0, q.p1, 0), xmax = vh_select(cum, q.p0, q.p1, n - 1);
    touch(xmin); touch(xmax);
    const int smin = xmin + base, smax = xmax + base;
    const int k = (n - 1) / 2;
    int ok = 1;
    int pv = gsb_code_at(med, false, R.coff, R.cscale, &ok);
    pv = min(max(pv, smin), smax + 1);
    const int nR = smax + 1 - pv, nL = pv - smin;
    // samples with a code in [pA, pB]
    auto count = [&](int pA, int pB) -> int {
        if (pB < pA) return 0;
        const int xa = pA - base, xb = pB - base;
        touch(xb);
        touch(xa - 1 < 0 ? 0 : xa);
        return vh_cle(cum, q.p0, q.p1, xb) - vh_cle(cum, q.p0, q.p1, xa - 1);
    };
    auto sdev = [&](bool right, int i) { return vf_dev(R, right ? pv + i : pv - 1 - i, med); };
    auto sfirst = [&](bool right, int nn, float thr, bool strict) -> int {
        float gf = thr / R.cscale;
        int g = (gf == gf && gf < 1e9f) ? (int)gf : nn;
        g = min(max(g, 0), nn);
        int guard = 0;
        while (g > 0 && guard++ < 100000) {
            const float d = sdev(right, g - 1);
            if (strict ? (d > thr) : (d >= thr)) g--; else break;
        }
        while (g < nn && guard++ < 100000) {
            const float d = sdev(right, g);
            if (strict ? !(d > thr) : !(d >= thr)) g++; else break;
        }
        return g;
    };
    int rlo = 0, rhi = nR, llo = 0, lhi = nL, c_best = -1;
    float t_best = CUDART_INF_F;
    while (rlo < rhi || llo < lhi) {
        // the middle candidate of the longer side: by symmetry of the two sides it halves the other one as well
        const bool right = (rhi - rlo) >= (lhi - llo);
        const int i = right ? (rlo + rhi) >> 1 : (llo + lhi) >> 1;
        const float thr = sdev(right, i);
        const int jR = sfirst(true, nR, thr, true), jL = sfirst(false, nL, thr, true);
        const int c = count(pv - jL, pv + jR - 1);
        if (c > k) {  // every candidate deviating at least thr passes
            rhi = min(rhi, sfirst(true, nR, thr, false));
            lhi = min(lhi, sfirst(false, nL, thr, false));
            rlo = min(rlo, rhi);
            llo = min(llo, lhi);
            t_best = thr;
            c_best = c;
        } else {      // every candidate deviating at most thr fails
            rlo = min(max(rlo, jR), rhi);
            llo = min(max(llo, jL), lhi);
        }
    }
    const float d0 = t_best;
    if (n & 1) return d0;
    float d1 = d0;
    if (!(c_best > k + 1)) {
        // the next larger deviation: the first occupied code on either side of the interval [l, r] deviating <= d0
        int lo = smin, hi = pv;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (vf_dev(R, m, med) <= d0) hi = m; else lo = m + 1; }
        const int l = lo;
        lo = pv; hi = smax + 1;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (vf_dev(R, m, med) > d0) hi = m; else lo = m + 1; }
        const int r = lo - 1;
        d1 = CUDART_INF_F;
        const int c_r = vh_cle(cum, q.p0, q.p1, r - base);          // samples with a code <= r
        if (c_r < n) { const int x = vh_select(cum, q.p0, q.p1, c_r); touch(x); d1 = fminf(d1, vf_dev(R, x + base, med)); }
        const int c_l = vh_cle(cum, q.p0, q.p1, l - 1 - base);      // samples with a code < l
        if (c_l > 0) { const int x = vh_select(cum, q.p0, q.p1, c_l - 1); touch(x); d1 = fminf(d1, vf_dev(R, x + base, med)); }
    }
    return __fdiv_rn(__fadd_rn(d0, d1), 2.0f);
}

__global__ void __launch_bounds__(VF_THREADS, 4) validate_hist_kernel(VfastArgs A, adb_config cfg) {
    extern __shared__ __align__(128) unsigned char vh_smem[];
    unsigned char *arena = vh_smem;                                   // operand tiles, later the cumulative counts
    uint16_t *cum = reinterpret_cast<uint16_t *>(arena);
    __shared__ VhShared H;
    VfScratch &S = H.S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int w = 0; w < VH_WARPS; w++) mbar_init(&H.bar[w], 1);
    }
    for (int i = tid; i < VH_ARENA / 16; i += VF_THREADS) reinterpret_cast<uint4 *>(arena)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&H.tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = H.tmem_slot;
    unsigned char *tileA = arena + (size_t)warp * VH_TILE_BYTES, *tileB = tileA + VH_A_BYTES;
    const uint32_t a_lo = tc_desc_lo(smem_u32(tileA), VH_A_LBO), b_lo = tc_desc_lo(smem_u32(tileB), VH_B_LBO);
    const uint32_t ab_hi = tc_desc_hi(128);
    uint32_t phase = 0;  // parity of this warp's barrier

    for (int r = blockIdx.x; r < A.B.n_reads; r += gridDim.x) {
        const int mb = r / A.B.batch_size;
        adb_record *rec = A.out + r;
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
        if (size <= 0) continue;
        // the window stays in global memory: the helpers take a 16-byte aligned base + the index of sample 0
        VfRead R;
        R.W16 = reinterpret_cast<const uint16_t *>((uintptr_t)gsrc.i16 & ~(uintptr_t)15);
        R.s0 = (int)(((uintptr_t)gsrc.i16 & 15) >> 1);
        R.n = size; R.coff = gsrc.coff; R.cscale = gsrc.cscale;
        R.kmin = 0; R.kmax = 0;
        const int16_t *W = gsrc.i16;
        int cok = 1;
        const int base = gsb_code_at(-20.0f, false, R.coff, R.cscale, &cok);
        if (!cok) continue;
        for (int w = tid; w < (int)(sizeof(adb_record) / 4); w += blockDim.x) ((uint32_t *)rec)[w] = 0;

        // ---- speculative inputs of the checks ----
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
        const int nLR = lrw;
        const double vLR85 = __dmul_rn((double)(nLR - 1), 0.85), vLR15 = __dmul_rn((double)(nLR - 1), 0.15);
        int pa_ = a_end, pb_ = pe0;
        clip_seg(pa_, pb_, size);
        const int nP = pb_ - pa_;
        const double vP85 = __dmul_rn((double)(nP - 1), 0.85), vP15 = __dmul_rn((double)(nP - 1), 0.15);
        const bool needP = mvs_geom || (pe_best > a_end);
        const bool ms_geom = cfg.detect_med_shift && haveA;

        // ---- the sample ranges of the statistics (query q works on [qa[q], qb[q]); empty: not wanted) ----
        int qa[VH_NQ], qb[VH_NQ];
#pragma unroll
        for (int q = 0; q < VH_NQ; q++) { qa[q] = 0; qb[q] = 0; }
        if (haveA) { qa[0] = 0; qb[0] = a_end; }                                              // adapter from 0
        if (haveA && a_start1 != 0) { qa[1] = a_start1; qb[1] = a_end; }                      // adapter behind the open pore
        if (rr_geom) { qa[2] = rb - lrw; qb[2] = rb; qa[3] = rb - lrw; qb[3] = rb; }          // local range: p15, p85
        if (needP) { qa[4] = a_end; qb[4] = pe_best; }                                        // poly(A)
        if (mvs_geom && nP > 0) { qa[5] = a_end; qb[5] = pe0; qa[6] = a_end; qb[6] = pe0; }   // its p15, p85
        if (mvs_geom) { qa[7] = a_end; qb[7] = min(a_end + msw, size); qa[8] = max(a_end - msw, 0); qb[8] = a_end; }
        if (size > pe_best) { qa[9] = pe_best; qb[9] = size; }                                // the rest
        if (ms_geom) {
            qa[10] = a_end; qb[10] = min(a_end + cfg.med_shift_window, full_len);
            qa[11] = max(a_end - cfg.med_shift_window, 0); qb[11] = a_end;
        }
#pragma unroll
        for (int q = 0; q < VH_NQ; q++) clip_seg(qa[q], qb[q], size);
        __syncthreads();
        if (tid == 0) {
            // cut points: sorted, distinct
            int c[2 * VH_NQ + 2], nc = 0;
            c[nc++] = 0; c[nc++] = size;
#pragma unroll
            for (int q = 0; q < VH_NQ; q++) if (qb[q] > qa[q]) { c[nc++] = qa[q]; c[nc++] = qb[q]; }
            for (int i = 1; i < nc; i++) { const int v = c[i]; int j = i - 1; while (j >= 0 && c[j] > v) { c[j + 1] = c[j]; j--; } c[j + 1] = v; }
            int m = 0;
            for (int i = 0; i < nc; i++) if (m == 0 || c[i] != c[m - 1]) c[m++] = c[i];
            const int np = m - 1;
            H.np = np;
            if (np <= VH_MAX_PIECES) {
                int bs = 0;
                for (int p = 0; p < np; p++) { H.cuts[p] = c[p]; H.bstart[p] = bs; bs += (c[p + 1] - c[p] + 31) >> 5; }
                H.cuts[np] = c[np];
                H.bstart[np] = bs;
            }
            H.n_low = 0; H.n_high = 0; H.unsettled = 0;
        }
        // ---- zero the accumulators ----
        {
            const uint32_t t0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64;
            vh_tmem_zero32(t0);
            vh_tmem_zero32(t0 + 32);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int np = H.np;
        if (np > VH_MAX_PIECES) continue;  // (uniform) more ranges than accumulators: validate_kernel
        // ---- one pass over the samples: one-hot tiles -> MMA -> clear ----
        {
            const int nb_total = H.bstart[np];
            int gb = warp, p = 0;
            int code = 0;
            bool valid = false;
            auto fetch = [&](int gq, int &pp, int &cd, bool &vd) {
                while (gq >= H.bstart[pp + 1]) pp++;
                const int j = H.cuts[pp] + ((gq - H.bstart[pp]) << 5) + lane;
                vd = j < H.cuts[pp + 1];
                cd = vd ? (int)W[j] : 0;
            };
            if (gb < nb_total) fetch(gb, p, code, valid);
            int n_low = 0, n_high = 0;
            while (gb < nb_total) {
                const int g2 = gb + VH_WARPS;
                int p2 = p, code2 = 0;
                bool valid2 = false;
                if (g2 < nb_total) fetch(g2, p2, code2, valid2);
                int cp = code - base;
                n_low += __popc(__ballot_sync(ADB_FULL, valid && cp < 0));
                n_high += __popc(__ballot_sync(ADB_FULL, valid && cp > VH_BINS - 1));
                cp = min(max(cp, 0), VH_BINS - 1);
                unsigned char *pA8 = tileA + (lane >> 4) * VH_A_LBO + (cp >> 4) * 16 + (lane & 15);
                unsigned char *pB8 = tileB + (lane >> 4) * VH_B_LBO + (cp & 15) * 16 + (lane & 15);
                if (valid) { *pA8 = 0x38; *pB8 = 0x38; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (tc_elect_one()) {
                    vh_mma_f8(tmem + (uint32_t)p * 16, a_lo, ab_hi, b_lo, ab_hi);
                    tc_commit(&H.bar[warp]);
                }
                __syncwarp();
                mbar_wait(&H.bar[warp], phase);
                phase ^= 1;
                if (valid) { *pA8 = 0; *pB8 = 0; }
                gb = g2; p = p2; code = code2; valid = valid2;
            }
            if (lane == 0 && (n_low | n_high)) { atomicAdd(&H.n_low, n_low); atomicAdd(&H.n_high, n_high); }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- accumulators -> counts (u16) in the arena (the tiles are idle and all-zero) ----
        {
            const int quad = warp & 3;
            for (int pp = 0; pp < 4; pp++) {
                const int p = (warp >> 2) * 4 + pp;
                if (p >= np) break;
                uint32_t v[16];
                vh_tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)p * 16, v);
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const unsigned c0 = (unsigned)__float2int_rn(__uint_as_float(v[2 * i]));
                    const unsigned c1 = (unsigned)__float2int_rn(__uint_as_float(v[2 * i + 1]));
                    w[i] = (c0 & 0xffffu) | (c1 << 16);
                }
                uint4 *dst = reinterpret_cast<uint4 *>(cum + (size_t)p * VH_BINS + (quad * 32 + lane) * 16);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- counts -> cumulative counts per piece (warp p: 8 rounds of 256 bins, 8 bins per lane) ----
        if (warp < np) {
            uint16_t *cp = cum + (size_t)warp * VH_BINS;
            unsigned carry = 0;
            for (int rd = 0; rd < VH_BINS / 256; rd++) {
                uint4 *q4 = reinterpret_cast<uint4 *>(cp + rd * 256 + lane * 8);
                const uint4 x = *q4;
                unsigned e[8] = {x.x & 0xffffu, x.x >> 16, x.y & 0xffffu, x.y >> 16, x.z & 0xffffu, x.z >> 16, x.w & 0xffffu, x.w >> 16};
#pragma unroll
                for (int i = 1; i < 8; i++) e[i] += e[i - 1];
                unsigned incl = e[7];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(ADB_FULL, incl, o); if (lane >= o) incl += t; }
                const unsigned off = carry + incl - e[7];
#pragma unroll
                for (int i = 0; i < 8; i++) e[i] += off;
                *q4 = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
                carry += __shfl_sync(ADB_FULL, incl, 31);
            }
        }
        __syncthreads();
        // ---- the statistics: lane q of warp 0 answers rank query q, then lanes 0..3 the four MADs ----
        if (warp == 0) {
            const int n_low = H.n_low, n_high = H.n_high;
            if (lane < VH_NQ) {
                int a = 0, b = 0;
#pragma unroll
                for (int q = 0; q < VH_NQ; q++) if (lane == q) { a = qa[q]; b = qb[q]; }
                int v0 = 0, v1 = 0;
                const int n = b - a;
                if (n > 0) {
                    const VhRange q = vh_range(H, a, b, size);
                    int k = (n - 1) / 2;
                    bool two = (n & 1) == 0;
                    if (lane == 2 || lane == 3 || lane == 5 || lane == 6) {  // np.percentile: floor of the virtual index
                        const double vv = (lane == 2) ? vLR15 : (lane == 3) ? vLR85 : (lane == 5) ? vP15 : vP85;
                        k = (int)floor(vv);
                        two = min(k + 1, n - 1) != k;
                    }
                    const int x0 = vh_select(cum, q.p0, q.p1, k);
                    const int x1 = two ? vh_select(cum, q.p0, q.p1, k + 1) : x0;
                    if (((x0 <= 0 || x1 <= 0) && n_low > 0) || ((x0 >= VH_BINS - 1 || x1 >= VH_BINS - 1) && n_high > 0)) H.unsettled = 1;
                    v0 = x0 + base;
                    v1 = x1 + base;
                }
                H.qv0[lane] = v0; H.qv1[lane] = v1; H.qn[lane] = n;
            }
            __syncwarp();
            if (lane < 4) {
                const int src = (lane == 0) ? 0 : (lane == 1) ? 1 : (lane == 2) ? 4 : 9;
                int a = 0, b = 0;
#pragma unroll
                for (int q = 0; q < VH_NQ; q++) if (src == q) { a = qa[q]; b = qb[q]; }
                const int n = b - a;
                float mad = CUDART_NAN_F;
                if (n > 0) {
                    const float x0 = vf_pa(R, H.qv0[src]);
                    const float med = (n & 1) ? x0 : __fdiv_rn(__fadd_rn(x0, vf_pa(R, H.qv1[src])), 2.0f);
                    mad = vh_mad(R, H, cum, vh_range(H, a, b, size), base, n, med);
                }
                H.mad[lane] = mad;
            }
        }
        __syncthreads();
        const bool unsettled = H.unsettled != 0;
        auto median_of = [&](int q) -> float {
            const int n = H.qn[q];
            if (n <= 0) return CUDART_NAN_F;
            const float x0 = vf_pa(R, H.qv0[q]);
            if (n & 1) return x0;
            return __fdiv_rn(__fadd_rn(x0, vf_pa(R, H.qv1[q])), 2.0f);
        };
        const float medA0 = median_of(0), medA1 = median_of(1), medP = median_of(4), medR = median_of(9);
        const float medAF = median_of(7), medBF = median_of(8), medMA = median_of(10), medMB = median_of(11);
        auto local_range = [&](int q15, int q85, double v15, double v85) -> double {
            if (H.qn[q15] <= 0 || H.qn[q85] <= 0) return CUDART_NAN;
            const int l15 = (int)floor(v15), l85 = (int)floor(v85);
            const double p85 = np_lerp_f32(vf_pa(R, H.qv0[q85]), vf_pa(R, H.qv1[q85]), __dsub_rn(v85, (double)l85));
            const double p15 = np_lerp_f32(vf_pa(R, H.qv0[q15]), vf_pa(R, H.qv1[q15]), __dsub_rn(v15, (double)l15));
            return __dsub_rn(p85, p15);
        };
        const double lrA = local_range(2, 3, vLR15, vLR85);
        const double lrP = local_range(5, 6, vP15, vP85);
        const float madA0 = H.mad[0], madA1 = H.mad[1], madP = H.mad[2], madR = H.mad[3];
        __syncthreads();
        // the arena goes back to all-zero operand tiles for the next read
        for (int i = tid; i < VH_ARENA / 16; i += VF_THREADS) reinterpret_cast<uint4 *>(arena)[i] = make_uint4(0, 0, 0, 0);
        if (unsettled) continue;  // (uniform) codes outside the histogram range matter: validate_kernel

        // ---- the checks (combined.py:394-580), as in validate_fast_kernel ----
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
            valid |= ADB_V_MVS;
            bool ok = false;
            if (mvs_geom) {
                const int L = nP;
                __syncthreads();
                if (!win_var || !win_mean) {
                    // exact numpy mean / variance of a short segment (one thread, pairwise order)
                    if (tid == 0) {
                        const int16_t *p = W + pa_;
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
                if (A.cand_followup != 0) followup = true; else defer = true;
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
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}
