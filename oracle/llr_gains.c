/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's LLR changepoint trace,
 * adapted/detect/_c_llr.pyx:22-236.  Compiled with `gcc -O2 -ffp-contract=off` (no -march), so every
 * floating-point operation is an individually rounded IEEE double op, like the stock build of the
 * Cython module, and `log` is glibc's.
 *
 * Pinned against oracle/_ref/_c_llr*.so (the reference's own pyx compiled from /root/reference) by
 * tests/test_oracle_llr.py: bit-identical gains for all three dispatch branches.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* numpy's pairwise summation for float64 (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum);
 * used by np.mean in the early-stop predicates (_c_llr.pyx:115-116,159-160,164-165). */
static double pairwise_sum_f64(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8) {
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum_f64(a, n2) + pairwise_sum_f64(a + n2, n - n2);
    }
}

double adb_oracle_pairwise_sum_f64(const double *a, int64_t n) { return pairwise_sum_f64(a, n); }

/* _c_llr.pyx:216-217: c = cumsum(x), c2 = cumsum(x*x), strictly sequential float64 */
void adb_oracle_cumsum(const double *x, int64_t n, double *c, double *c2) {
    double s = 0., s2 = 0.;
    for (int64_t i = 0; i < n; i++) {
        s += x[i];
        s2 += x[i] * x[i];
        c[i] = s;
        c2[i] = s2;
    }
}

/* _c_llr.pyx:22-37 */
static inline double var_c(int64_t start, int64_t end, const double *c, const double *c2) {
    if (start == end) return 0.;
    if (start == 0) {
        double m = c[end - 1] / (double)end;
        return c2[end - 1] / (double)end - m * m;
    }
    double n = (double)(end - start);
    double m = (c[end - 1] - c[start - 1]) / n;
    return (c2[end - 1] - c2[start - 1]) / n - m * m;
}

static inline double gain_at(int64_t s, int64_t e, int64_t i, double var_summed, const double *c,
                             const double *c2) {
    double head = (double)(i - s) * log(var_c(s, i, c, c2));
    double tail = (double)(e - i) * log(var_c(i, e, c, c2));
    return var_summed - (head + tail);
}

/* mean(diff(gains[lo:hi:stride])) as numpy computes it (pairwise float64 sum / count).
 * An empty diff gives nan (numpy: mean of empty slice) -> both predicates false. */
static double mean_diff(const double *gains, int64_t n, int64_t lo, int64_t hi, int64_t stride) {
    double buf[4096];
    /* python slice semantics: negative lo wraps once, then clamps */
    if (lo < 0) { lo += n; if (lo < 0) lo = 0; }
    if (hi > n) hi = n;
    int64_t cnt = 0;
    double prev = 0.;
    int64_t k = 0;
    for (int64_t j = lo; j < hi; j += stride, k++) {
        if (k > 0 && cnt < 4096) buf[cnt++] = gains[j] - prev;
        prev = gains[j];
    }
    if (cnt == 0) return NAN;
    return pairwise_sum_f64(buf, cnt) / (double)cnt;
}

/*
 * mode 0: _gains (67-88); mode 1: _gains_w_early_stop (91-123); mode 2: _gains_w_polya_early_stop (126-173).
 * gains must hold n doubles; it is zero-filled here (np.zeros_like(c), :80).
 * Returns 0, or -1 if an early-stop stride is not a multiple of stride (the reference asserts, :102,139-140).
 */
int adb_oracle_llr_gains(const double *c, const double *c2, int64_t n, int64_t start, int64_t end,
                         int64_t offset_head, int64_t offset_tail, int64_t stride, int mode,
                         int64_t a_window, int64_t a_stride, int64_t p_window, int64_t p_stride,
                         double *gains) {
    memset(gains, 0, (size_t)n * sizeof(double));
    if (mode == 1 && (a_stride % stride) != 0) return -1;
    if (mode == 2 && ((a_stride % stride) != 0 || (p_stride % stride) != 0)) return -1;
    double var_summed = (double)(end - start) * log(var_c(start, end, c, c2));
    int adapter_found = 0;
    for (int64_t i = start + offset_head; i < end - offset_tail; i += stride) {
        if (mode == 1) {
            if (i >= start + offset_head + a_window && ((i - (start + offset_head)) % a_stride) == 0) {
                if (mean_diff(gains, n, i - a_window, i, stride) < 0) break;
            }
        } else if (mode == 2) {
            if (!adapter_found && i >= start + offset_head + a_window &&
                ((i - (start + offset_head)) % a_stride) == 0) {
                if (mean_diff(gains, n, i - a_window, i, stride) < 0) adapter_found = 1;
            }
            if (adapter_found) {
                if (mean_diff(gains, n, i - p_window, i, stride) > 0) break;
            }
        }
        gains[i] = gain_at(start, end, i, var_summed, c, c2);
    }
    return 0;
}
