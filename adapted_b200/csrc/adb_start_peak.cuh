// start-peak primary (detect_rna_start_peak, adapted/detect/start_peak.py:7-119; seam combined.py:312-355).
//
// Literal semantics, SURVEY.md A.9: the raw signal is searched for the first sample above open_pore_pa inside
// raw[:end_idx] where end_idx is the DOWNSCALED bound (sic); the start peak is the maximum of the downscaled row in
// [offset1, start_peak_max_idx); the next-greater index is the first downscaled bin in
// [start_peak_max_idx + offset2, end_idx) strictly above it (or the slice start when there is none).
// A read whose slices are empty yields a row of Nones in the reference; pandas then turns the integer columns of the
// whole minibatch frame into floats and every read of that minibatch dies on a float slice index -- the finish kernel
// reproduces that ("poisoned" minibatch).
#pragma once
#include "adb_common.cuh"
#include "adb_ctx.cuh"

struct SpRow {
    int ok;             // 0: row of Nones
    int idx, next_idx, open_idx, flag;
    float pa, next_pa;
    int _pad;
};

template <class F>
__device__ __forceinline__ float sp_block_mean(F f, int factor) {
    return __fdiv_rn(np_sum_f32_leaf(f, factor), (float)factor);
}

// one CTA (128 threads) per read
__global__ void __launch_bounds__(128) start_peak_kernel(BatchDev B, adb_config cfg, SpRow *rows, int *given, int *poison,
                                                         int *batch_status) {
    __shared__ int s_first;
    __shared__ float s_maxv;
    __shared__ int s_anynan, s_maxidx, s_next;
    __shared__ float s_nextv;
    const int r = blockIdx.x, mb = r / B.batch_size;
    const ReadSrc src = make_src(B, r);
    const int f = cfg.sp_downscale_factor, o1 = cfg.sp_offset1, mx = cfg.start_peak_max_idx, o2 = cfg.sp_offset2;
    const int Lds = (B.m + f - 1) / f;
    const int full_len = B.full_lens[r];
    const int end_idx = min(full_len, B.m) / f;
    SpRow row;
    row.ok = 0; row.idx = row.next_idx = row.open_idx = row.flag = 0; row.pa = row.next_pa = 0.f; row._pad = 0;
    if (threadIdx.x == 0) { s_first = 0x7fffffff; s_anynan = 0; s_maxidx = 0x7fffffff; s_next = 0x7fffffff; }
    __syncthreads();
    if (end_idx <= 0) {
        // np.argmax over an empty slice raises outside the try (start_peak.py:25-29): the whole call fails
        if (threadIdx.x == 0) {
            atomicMin(&batch_status[mb], (int)ADB_ERR_EMPTY_TRACE);
            rows[r] = row;
            given[2 * r] = 0; given[2 * r + 1] = 0;
        }
        return;
    }
    auto ds = [&](int b) -> float {  // downscaled bin b of the NaN-padded row (zero padded past m)
        const int j0 = b * f;
        if (j0 + f > src.n && src.n < B.m) {
            // the block reaches past the read: NaN if any of its in-window samples is padding
            for (int k = 0; k < f; k++) { const int j = j0 + k; if (j < B.m && j >= src.n) return CUDART_NAN_F; }
        }
        return sp_block_mean([&](int k) { const int j = j0 + k; return (j < B.m && j < src.n) ? src.pa(j) : 0.0f; }, f);
    };
    // first raw sample above open_pore_pa within raw[:end_idx]
    {
        int first = 0x7fffffff;
        const float thr = (float)cfg.open_pore_pa;
        for (int j = threadIdx.x; j < min(end_idx, src.n); j += blockDim.x)
            if (src.pa(j) > thr) { first = j; break; }
        if (first != 0x7fffffff) atomicMin(&s_first, first);
    }
    // max over ds[o1:mx]
    const int a0 = min(max(o1, 0), Lds), a1 = min(max(mx, 0), Lds);
    float mymax = -CUDART_INF_F;
    bool mynan = false;
    for (int b = a0 + threadIdx.x; b < a1; b += blockDim.x) {
        const float v = ds(b);
        if (!(v == v)) mynan = true; else mymax = fmaxf(mymax, v);
    }
    if (mynan) s_anynan = 1;
    if (threadIdx.x == 0) s_maxv = -CUDART_INF_F;
    __syncthreads();
    // block max (float bits trick is not order preserving for negatives; reduce through shuffles + shared)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(ADB_FULL, mymax, o));
    __shared__ float wmax[4];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mymax;
    __syncthreads();
    const float maxv_clean = fmaxf(fmaxf(wmax[0], wmax[1]), fmaxf(wmax[2], wmax[3]));
    const bool empty1 = (a1 <= a0);
    const float max_ = s_anynan ? CUDART_NAN_F : maxv_clean;
    // first index equal to max_ (NaN: none -> argmax of all-False = 0)
    for (int b = a0 + threadIdx.x; b < a1; b += blockDim.x)
        if (ds(b) == max_) { atomicMin(&s_maxidx, b); break; }
    // next greater in ds[mx+o2 : end_idx)
    const int n0 = min(max(mx + o2, 0), Lds), n1 = min(end_idx, Lds);
    const bool empty2 = (n1 <= n0);
    for (int b = n0 + threadIdx.x; b < n1; b += blockDim.x)
        if (ds(b) > max_) { atomicMin(&s_next, b); break; }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (empty1 || empty2) {
            atomicExch(&poison[mb], 1);
        } else {
            const int max_idx = (s_maxidx == 0x7fffffff) ? a0 : s_maxidx;
            const int next_idx = (s_next == 0x7fffffff) ? n0 : s_next;
            const float next_v = ds(next_idx);
            int op = (s_first == 0x7fffffff) ? 0 : s_first / f;
            const bool have_op = op > 0;
            row.ok = 1;
            row.idx = max_idx * f;
            row.pa = max_;
            row.next_idx = next_idx * f;
            row.next_pa = next_v;
            if (have_op && fabs((double)next_idx - (double)op) <= 2.0 + 0.01 * fabs((double)op)) {
                row.flag = 1; row.open_idx = op * f;
            } else if (have_op && max_idx < op && op < next_idx) {
                row.flag = 2; row.open_idx = op * f;
            }
        }
        rows[r] = row;
        given[2 * r] = row.next_idx;
        given[2 * r + 1] = row.next_idx;
    }
}

__global__ void start_peak_finish_kernel(int n_reads, int batch_size, const SpRow *rows, const int *poison, adb_record *recs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    adb_record *rec = recs + r;
    if (poison[r / batch_size]) {
        rec->valid = 0; rec->success = 0; rec->fail_code = ADB_FAIL_EXC_SLICE_INDEX; rec->mvs_fail_mask = 0;
        return;
    }
    if (!(rec->valid & ADB_V_FIELDS)) return;  // validate_boundaries raised: DetectResults(success=False, fail_reason=str(e))
    const SpRow row = rows[r];
    rec->valid |= ADB_V_START_PEAK;
    rec->sp_idx = row.idx; rec->sp_pa = row.pa; rec->sp_next_idx = row.next_idx; rec->sp_next_pa = row.next_pa;
    rec->sp_flag = row.flag;
    if (row.flag) {
        rec->valid |= ADB_V_SP_OPEN_PORE;
        rec->sp_open_pore_idx = row.open_idx;
        rec->success = 0;  // success and not flagged; the host appends "+<type>" when it had already failed (fail_code != 0)
    }
}

static int start_peak_primary(adb_ctx *ctx, const BatchDev &B, const adb_config &cfg, int *given, adb_record *,
                              int *status, cudaStream_t st) {
    const int n_batches = (B.n_reads + B.batch_size - 1) / B.batch_size;
    if (ctx->sp_rows.ensure(sizeof(SpRow) * (size_t)B.n_reads + sizeof(int) * (size_t)n_batches + 64)) {
        set_err("cudaMalloc start-peak rows");
        return ADB_ERR_CUDA;
    }
    SpRow *rows = (SpRow *)ctx->sp_rows.p;
    int *poison = (int *)(rows + B.n_reads);
    CUDA_TRY(cudaMemsetAsync(poison, 0, sizeof(int) * (size_t)n_batches, st));
    {
        KernelTimer t(ctx, 6, st);
        start_peak_kernel<<<B.n_reads, 128, 0, st>>>(B, cfg, rows, given, poison, status);
    }
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return ADB_OK;
}

static int start_peak_finish(adb_ctx *ctx, const BatchDev &B, const adb_config &, adb_record *recs, cudaStream_t st) {
    const int n_batches = (B.n_reads + B.batch_size - 1) / B.batch_size;
    (void)n_batches;
    SpRow *rows = (SpRow *)ctx->sp_rows.p;
    int *poison = (int *)(rows + B.n_reads);
    {
        KernelTimer t(ctx, 6, st);
        start_peak_finish_kernel<<<(B.n_reads + 127) / 128, 128, 0, st>>>(B.n_reads, B.batch_size, rows, poison, recs);
    }
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return ADB_OK;
}
