// Per-read kernel: one CTA per read, persistent over the reads of a call.
//
//   LLR primary (combined_detect_llr2, combined.py:122-227):
//     clip / normalise with the minibatch-global med/MAD -> float32 mean-pool downscale (numpy's pairwise order)
//     -> sequential float64 prefix sums -> adapter LLR trace -> first surviving peak (+ plateau / split fixes)
//     -> poly(A) LLR trace -> spike rule -> validate_boundaries + partition statistics -> record.
//   GIVEN primary (CNN / start-peak): boundaries come from a device array; validation (+ the CNN path's
//     "hail mary" LLR fallback, combined.py:251-301) is the same code.
//
// Shared memory (dynamic): [trace f64 nds][c f64 nds][c2 f64 nds][peak scratch][small]; the downscaled float32
// row aliases the trace, the select histogram and the staging buffer alias c/c2 once the traces are done.
#pragma once
#include "adb_common.cuh"
#include "adb_global.cuh"
#include "adb_llr.cuh"
#include "adb_select.cuh"
#include "adb_validate.cuh"

#define ADB_READ_THREADS 128

struct ReadKernelArgs {
    BatchDev B;
    const GselState *gstates;   // LLR: per-minibatch med / mad
    const int *given;           // GIVEN: [n_reads][given_stride] = adapter_end, polya_end, topk...
    int given_stride;
    int given_ntopk;            // number of top-k entries per read (-1: polya_end_topk is None)
    int mode;                   // ADB_METHOD_*
    int nds_max;                // capacity of the trace buffers
    int peak_cap;
    adb_record *out;
    float *series;              // [gridDim.x][2][m]
    int *batch_status;          // [n_batches]
};

__host__ __device__ inline size_t read_kernel_smem_bytes(int nds_max, int peak_cap) {
    size_t trace = (size_t)nds_max * 8;
    size_t cc2 = (size_t)nds_max * 16;
    size_t need = ADB_SEL_SMEM_BYTES + 64 + (ADB_STAGE_HIST + ADB_STAGE_CHUNK) * 4;
    if (cc2 < need) cc2 = need;
    cc2 = (cc2 + 15) & ~(size_t)15;
    size_t peaks = (size_t)peak_cap * 6;
    peaks = (peaks + 15) & ~(size_t)15;
    return trace + cc2 + peaks + 256;
}

// float32 mean of one downscale block in numpy's pairwise order.  f(k) = k-th sample of the block.
template <class F>
__device__ __forceinline__ float block_mean_f32(F f, int factor) {
    return __fdiv_rn(np_sum_f32_leaf(f, factor), (float)factor);
}

// LLR boundaries on the downscaled, normalised row ds[0..nds) (already in shared memory at `ds`).
// Outputs (uniform): ae_ds (-1: no candidate), pe_ds (0: none).  CTA-wide.
__device__ void llr_boundaries_cta(const float *ds, int nds, double *trace, double *c, double *c2,
                                   const PeakScratch &PS, int *itmp, double *dtmp, const adb_config &cfg,
                                   bool adapter_stage, int &ae_ds, int &pe_ds) {
    // sequential float64 prefix sums (_c_llr.pyx:216-217): two independent chains on two warps
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nds; i++) { s = __dadd_rn(s, (double)ds[i]); c[i] = s; }
    } else if (threadIdx.x == 32) {
        double s2 = 0.0;
        for (int i = 0; i < nds; i++) { double x = (double)ds[i]; s2 = __dadd_rn(s2, __dmul_rn(x, x)); c2[i] = s2; }
    }
    __syncthreads();
    ae_ds = 0;
    pe_ds = 0;
    if (adapter_stage) {
        // adapter trace: start 0, end nds-1, head 5, tail 5 (combined.py:155-170)
        cta_llr_gains(c, c2, nds, 0, nds - 1, 5, 5, 1, trace);
        int s0, e0;
        cta_trace_support(trace, nds, s0, e0, itmp);
        if (threadIdx.x == 0) dtmp[0] = lane_nanstd(trace, s0, e0);
        __syncthreads();
        const double pmin = __dmul_rn(cfg.adapter_peak_prominence, dtmp[0]);
        const double wmin = (double)(cfg.adapter_peak_width / cfg.downscale_factor);
        if (threadIdx.x < 32) {
            int r = warp_adapter_end(trace, nds, s0, e0, pmin, wmin, cfg.adapter_peak_rel_height, PS);
            if (threadIdx.x == 0) itmp[4] = r;
        }
        __syncthreads();
        ae_ds = itmp[4];
        __syncthreads();
        if (ae_ds < 0) return;
        // poly(A) trace: start = adapter_end, head 1, tail 1, same prefix sums (combined.py:189-204)
        cta_llr_gains(c, c2, nds, ae_ds, nds - 1, 1, 1, 1, trace);
    } else {
        // hail-mary variant: a single trace (head 5, tail 5) feeds the spike rule directly (combined.py:277-292)
        cta_llr_gains(c, c2, nds, 0, nds - 1, 5, 5, 1, trace);
    }
    if (threadIdx.x < 32) {
        int err = 0;
        int r = warp_polya_end(trace, nds, PS, &err);
        if (threadIdx.x == 0) itmp[5] = r;
    }
    __syncthreads();
    pe_ds = itmp[5];
    __syncthreads();
}

__global__ void __launch_bounds__(ADB_READ_THREADS) read_kernel(ReadKernelArgs A, adb_config cfg) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nds_max = A.nds_max;
    double *trace = (double *)smem;
    double *c = trace + nds_max;
    size_t cc2 = (size_t)nds_max * 16;
    {
        size_t need = ADB_SEL_SMEM_BYTES + 64 + (ADB_STAGE_HIST + ADB_STAGE_CHUNK) * 4;
        if (cc2 < need) cc2 = need;
        cc2 = (cc2 + 15) & ~(size_t)15;
    }
    double *c2 = c + nds_max;
    unsigned char *after = (unsigned char *)c + cc2;
    PeakScratch PS;
    PS.cap = A.peak_cap;
    PS.pk = (unsigned short *)after;
    PS.stack = PS.pk + A.peak_cap;
    PS.status = (unsigned char *)(PS.stack + A.peak_cap);
    PS.flags = PS.status + A.peak_cap;
    unsigned char *small = after + (((size_t)A.peak_cap * 6 + 15) & ~(size_t)15);
    uint32_t *kbuf = (uint32_t *)small;          // 8 words
    int *itmp = (int *)(small + 32);             // 8 ints
    double *dtmp = (double *)(small + 64);       // 8 doubles
    int *topk_sh = (int *)(small + 128);         // ADB_MAX_CAND ints
    float *ds = (float *)trace;

    ValCtx C;
    C.cfg = &cfg;
    C.S = sel_scratch_from((unsigned char *)c);
    C.stage = (float *)((unsigned char *)c + ((ADB_SEL_SMEM_BYTES + 63) & ~63));
    C.kbuf = kbuf;
    C.itmp = itmp;
    C.dtmp = dtmp;
    C.series_a = A.series + (size_t)blockIdx.x * 2 * A.B.m;
    C.series_b = C.series_a + A.B.m;

    for (int r = blockIdx.x; r < A.B.n_reads; r += gridDim.x) {
        const int mb = r / A.B.batch_size;
        adb_record *rec = A.out + r;
        const ReadSrc src = make_src(A.B, r);
        const int full_len = A.B.full_lens[r];
        C.src = src;
        C.int_keys = (src.i16 != nullptr) && (src.cscale > 0.0f);
        __syncthreads();
        // zero the record (so unset groups read as zeros) -- 512 B = 128 words
        for (int w = threadIdx.x; w < (int)(sizeof(adb_record) / 4); w += blockDim.x) ((uint32_t *)rec)[w] = 0;
        __syncthreads();

        PrimaryBounds PB;
        PB.adapter_start = 0; PB.adapter_end = 0; PB.polya_end = 0; PB.n_topk = -1; PB.topk = topk_sh;

        if (A.mode == ADB_METHOD_LLR) {
            const GselState gs = A.gstates[mb];
            if (gs.status != ADB_OK) continue;  // minibatch lost (host raises)
            const int T = cfg.max_obs_trace, A0 = cfg.min_obs_adapter, f = cfg.downscale_factor;
            // number of non-NaN downscaled bins (combined.py:133-154)
            int nds;
            {
                const int L = min(src.n, T) - A0;
                if (src.n >= T) nds = (T - A0 + f - 1) / f;
                else nds = (L > 0) ? L / f : 0;
                if (T - A0 <= 0) nds = 0;
            }
            if (nds <= 0) {
                if (threadIdx.x == 0) atomicMin(&A.batch_status[mb], (int)ADB_ERR_EMPTY_TRACE);
                continue;
            }
            const float med = gs.med, mad = gs.mad;
            const float lo = (float)((double)med - (double)mad * cfg.sig_norm_outlier_thresh);
            const float hi = (float)((double)med + (double)mad * cfg.sig_norm_outlier_thresh);
            for (int b = threadIdx.x; b < nds; b += blockDim.x) {
                const int j0 = A0 + b * f;
                ds[b] = block_mean_f32(
                    [&](int k) {
                        const int j = j0 + k;
                        if (j >= T) return 0.0f;  // np.pad zero padding of a ragged last block
                        float v = src.pa(j);
                        v = (v < lo) ? lo : ((v > hi) ? hi : v);  // np.clip keeps NaN
                        return __fdiv_rn(__fsub_rn(v, med), mad);
                    },
                    f);
            }
            __syncthreads();
            int ae_ds, pe_ds;
            llr_boundaries_cta(ds, nds, trace, c, c2, PS, itmp, dtmp, cfg, true, ae_ds, pe_ds);
            if (ae_ds > 0) PB.adapter_end = ae_ds * f + A0;
            if (ae_ds >= 0 && pe_ds > 0) {
                PB.polya_end = pe_ds * f + A0;
                PB.n_topk = 1;
                if (threadIdx.x == 0) topk_sh[0] = PB.polya_end;
            }
            __syncthreads();
        } else {
            const int *g = A.given + (size_t)r * A.given_stride;
            PB.adapter_end = g[0];
            PB.polya_end = g[1];
            PB.n_topk = A.given_ntopk;
            if (threadIdx.x < ADB_MAX_CAND) topk_sh[threadIdx.x] = (threadIdx.x < A.given_ntopk) ? g[1 + threadIdx.x] : 0;
            __syncthreads();
        }

        val_window_bounds(C);
        validate_boundaries_cta(C, PB, full_len, rec);

        if (A.mode == ADB_METHOD_CNN) {
            // "hail mary" LLR fallback for short reads (combined.py:251-301)
            __syncthreads();
            const bool failed = (rec->success == 0) && (rec->valid & ADB_V_FIELDS);
            const int ae = PB.adapter_end, pe = PB.polya_end;
            if (failed && ae > 0 && pe > 0 && pe - ae > 1000 && full_len < 2 * cfg.max_obs_adapter &&
                cfg.fallback_to_llr_short_reads) {
                // per-read normalisation over signal[:min(max_obs_trace, full_len)] (NaN padding dropped)
                const int nn = min(min(cfg.max_obs_trace, full_len), A.B.m);
                const int nvalid = min(nn, src.n);
                const float med = seg_median(C, 0, nvalid);
                const float mad = seg_mad(C, 0, nvalid, med);
                if (nvalid > 0 && mad == 0.0f) {
                    // normalize_signal raises ValueError("MAD normalization failed: scale is 0") inside the try
                    if (threadIdx.x == 0) { rec->valid = 0; rec->success = 0; rec->fail_code = ADB_FAIL_EXC_MAD_ZERO; }
                    __syncthreads();
                    continue;
                }
                const float lo = (float)((double)med - (double)mad * cfg.sig_norm_outlier_thresh);
                const float hi = (float)((double)med + (double)mad * cfg.sig_norm_outlier_thresh);
                const int f = cfg.downscale_factor;
                // norm_signal[ae:pe] (clipped to nn), zero padded to a multiple of f, NaN where the read has ended
                int sa = min(ae, nn), sb = min(pe, nn);
                const int seglen = sb - sa;
                const int nblk = (seglen + f - 1) / f;
                // leading NaN-free blocks: a block is NaN iff it reaches past the read's last sample
                int nds;
                if (sb <= nvalid) nds = nblk;
                else nds = (nvalid > sa) ? (nvalid - sa) / f : 0;
                if (nds > nds_max) nds = nds_max;  // cannot exceed (pe-ae)/f <= window/f
                if (nds <= 0) {
                    if (threadIdx.x == 0) { rec->valid = 0; rec->success = 0; rec->fail_code = ADB_FAIL_EXC_EMPTY_TRACE; }
                    __syncthreads();
                    continue;
                }
                __syncthreads();
                for (int b = threadIdx.x; b < nds; b += blockDim.x) {
                    const int j0 = sa + b * f;
                    ds[b] = block_mean_f32(
                        [&](int k) {
                            const int j = j0 + k;
                            if (j >= sb) return 0.0f;
                            float v = src.pa(j);
                            v = (v < lo) ? lo : ((v > hi) ? hi : v);
                            return __fdiv_rn(__fsub_rn(v, med), mad);
                        },
                        f);
                }
                __syncthreads();
                int ae_ds, pe_ds;
                llr_boundaries_cta(ds, nds, trace, c, c2, PS, itmp, dtmp, cfg, false, ae_ds, pe_ds);
                if (pe_ds > 0) {
                    PB.polya_end = pe_ds * f + ae;
                    PB.n_topk = 1;
                    __syncthreads();
                    if (threadIdx.x == 0) topk_sh[0] = PB.polya_end;
                    __syncthreads();
                    for (int w = threadIdx.x; w < (int)(sizeof(adb_record) / 4); w += blockDim.x) ((uint32_t *)rec)[w] = 0;
                    __syncthreads();
                    validate_boundaries_cta(C, PB, full_len, rec);
                }
            }
        }
        __syncthreads();
    }
}
