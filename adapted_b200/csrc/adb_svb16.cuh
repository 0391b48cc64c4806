// Device-side decode of VBZ-style compressed reads: svb16 + zig-zag + delta, i.e. pod5's signal
// compression minus its zstd stage (pod5 `c++/pod5_format/svb16`: StreamVByte with one key BIT per 16-bit value --
// 0: one data byte, 1: two data bytes little endian, key bits LSB first, all keys in front of the data -- applied to
// zigzag(sample[i] - sample[i-1]) with sample[-1] = 0).  pod5 is absent from the image: the layout is restated from
// its published sources, PARITY UNPINNED; what is tested is the round trip against the encoder in
// adapted_b200/svb16.py and tests/.
//
// Per read the stream is [ceil(n / 8) key bytes, zero padded to a multiple of 4][data bytes]; streams start 16-byte
// aligned in the blob, which carries 16 bytes of slack behind its last stream.
//
// One WARP per read, 128 values per iteration: lane l owns values 4l .. 4l+3.  Its byte offset is the running offset
// + 4l + popc(key bits before its nibble) (control-bit prefix scan = four popcounts and a mask); it fetches the 8
// bytes there with three aligned word loads and two funnel shifts, peels its four values (1 or 2 bytes each by its key
// nibble), undoes the zig-zag, and the delta prefix scan is a lane-local running sum + one warp scan of the lane
// totals.  The running byte offset and the last sample stay in registers, so a read is decoded in one pass with no
// barrier and no shared memory; the reads of a launch keep ~48 warps per SM in flight to cover the load latency.
// Samples leave as int16 in the ragged layout the detection kernels take (ADB_SIG_I16).
#pragma once
#include "adb_common.cuh"

struct SvbArgs {
    const uint8_t *comp;        // compressed blob
    const int64_t *comp_off;    // [n_reads] byte offset of every stream (16-byte aligned)
    const int32_t *n_samples;   // [n_reads] values per stream
    const int64_t *out_off;     // [n_reads] element offset of the read in `out`
    int16_t *out;
    int n_reads;
};

#define ADB_SVB_THREADS 256

__global__ void __launch_bounds__(ADB_SVB_THREADS) svb16_decode_kernel(SvbArgs A) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int kword = lane >> 3, kshift = (lane & 7) * 4;
    for (int r = warp; r < A.n_reads; r += nwarps) {
        const int n = A.n_samples[r];
        const uint8_t *base = A.comp + A.comp_off[r];
        const uint4 *keys4 = reinterpret_cast<const uint4 *>(base);
        const uint32_t *keys = reinterpret_cast<const uint32_t *>(base);
        const int key_words = ((n + 7) / 8 + 3) >> 2;
        const uint32_t *dw = keys + key_words;
        uint16_t *out = reinterpret_cast<uint16_t *>(A.out + A.out_off[r]);
        uint32_t bpos = 0, val = 0;
        for (int i0 = 0; i0 < n; i0 += 128) {
            const int w0i = i0 >> 5;
            uint32_t k[4];
            if (w0i + 4 <= key_words) {
                const uint4 q = keys4[w0i >> 2];  // i0 is a multiple of 128: four words, 16-byte aligned
                k[0] = q.x; k[1] = q.y; k[2] = q.z; k[3] = q.w;
            } else {
#pragma unroll
                for (int t = 0; t < 4; t++) k[t] = (w0i + t < key_words) ? keys[w0i + t] : 0u;
            }
            // key bits of values >= n do not count (an encoder leaves them zero; do not rely on it)
            const int left = n - i0;
            if (left < 128) {
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int l = left - 32 * t;
                    if (l <= 0) k[t] = 0u;
                    else if (l < 32) k[t] &= (1u << l) - 1u;
                }
            }
            const uint32_t p0 = __popc(k[0]), p1 = __popc(k[1]), p2 = __popc(k[2]), p3 = __popc(k[3]);
            const uint32_t kw = kword == 0 ? k[0] : (kword == 1 ? k[1] : (kword == 2 ? k[2] : k[3]));
            const uint32_t words_before = kword == 0 ? 0u : (kword == 1 ? p0 : (kword == 2 ? p0 + p1 : p0 + p1 + p2));
            const uint32_t nib = (kw >> kshift) & 0xFu;
            const uint32_t my_b = bpos + 4u * lane + words_before + __popc(kw & ((1u << kshift) - 1u));
            const uint32_t a = my_b >> 2, sh = (my_b & 3u) * 8u;
            const int mine = left - 4 * lane;  // values of this lane that exist
            uint32_t w0 = 0u, w1 = 0u, w2 = 0u;
            if (mine > 0) { w0 = dw[a]; w1 = dw[a + 1]; w2 = dw[a + 2]; }  // at most 11 bytes past the stream: the slack
            uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
            int pre[4];
            int acc = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool two = (nib >> e) & 1u;
                const uint32_t u = two ? (lo & 0xffffu) : (lo & 0xffu);
                const uint32_t s = two ? 16u : 8u;
                lo = __funnelshift_r(lo, hi, s);
                hi >>= s;
                const int d = (int)((u >> 1) ^ (0u - (u & 1u)));
                acc += (e < mine) ? d : 0;
                pre[e] = acc;
            }
            int incl = acc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(ADB_FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t lane_base = val + (uint32_t)(incl - acc);
            const int j = i0 + 4 * lane;
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (e < mine) out[j + e] = (uint16_t)(lane_base + (uint32_t)pre[e]);
            val += (uint32_t)__shfl_sync(ADB_FULL, incl, 31);
            bpos += 128u + p0 + p1 + p2 + p3;
        }
    }
}
