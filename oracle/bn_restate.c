/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Restatement of bottleneck.move_mean / bottleneck.move_var (ddof=0) for float32 input, following the
 * recurrences of bottleneck's src/move_template.c (v1.3.x; the reference leaves bottleneck unpinned,
 * setup.py:39 / environment.yml:15, py3.8 => 1.3.x).  bottleneck is NOT installed in this image and
 * there is no network, so this cannot be checked against the real library: PARITY UNPINNED.
 * Call sites in the reference: adapted/detect/mvs.py:93-96,103-106.
 *
 * All arithmetic is C `float` (the template instantiates its accumulators in the input dtype);
 * `1.0 / count` is a double division rounded to float, `x / count` is a float division by (float)count.
 * gcc -O2 -ffp-contract=off, no -march: no FMA.
 */
#include <math.h>
#include <stdint.h>

void adb_oracle_move_mean_f32(const float *a, int64_t n, int64_t window, float *y) {
    int64_t min_count = window, count = 0, i = 0;
    float asum = 0.f, ai, aold, count_inv;
    for (; i < min_count - 1 && i < n; i++) {
        ai = a[i];
        if (ai == ai) { asum += ai; count += 1; }
        y[i] = NAN;
    }
    for (; i < window && i < n; i++) {
        ai = a[i];
        if (ai == ai) { asum += ai; count += 1; }
        y[i] = count >= min_count ? asum / count : NAN;
    }
    count_inv = 1.0 / count;
    for (; i < n; i++) {
        ai = a[i];
        aold = a[i - window];
        if (ai == ai) {
            if (aold == aold) {
                asum += ai - aold;
            } else {
                asum += ai;
                count++;
                count_inv = 1.0 / count;
            }
        } else {
            if (aold == aold) {
                asum -= aold;
                count--;
                count_inv = 1.0 / count;
            }
        }
        y[i] = count >= min_count ? asum * count_inv : NAN;
    }
}

void adb_oracle_move_var_f32(const float *a, int64_t n, int64_t window, float *y) {
    const int ddof = 0;
    int64_t min_count = window, count = 0, i = 0;
    float delta, amean = 0.f, assqdm = 0.f, ai, aold, yi, count_inv, ddof_inv;
    for (; i < min_count - 1 && i < n; i++) {
        ai = a[i];
        if (ai == ai) {
            count += 1;
            delta = ai - amean;
            amean += delta / count;
            assqdm += delta * (ai - amean);
        }
        y[i] = NAN;
    }
    for (; i < window && i < n; i++) {
        ai = a[i];
        if (ai == ai) {
            count += 1;
            delta = ai - amean;
            amean += delta / count;
            assqdm += delta * (ai - amean);
        }
        if (count >= min_count) {
            if (assqdm < 0) assqdm = 0;
            yi = assqdm / (count - ddof);
        } else {
            yi = NAN;
        }
        y[i] = yi;
    }
    count_inv = 1.0 / count;
    ddof_inv = 1.0 / (count - ddof);
    for (; i < n; i++) {
        ai = a[i];
        aold = a[i - window];
        if (ai == ai) {
            if (aold == aold) {
                delta = ai - aold;
                aold -= amean;
                amean += delta * count_inv;
                ai -= amean;
                assqdm += (ai + aold) * delta;
            } else {
                count++;
                count_inv = 1.0 / count;
                ddof_inv = 1.0 / (count - ddof);
                delta = ai - amean;
                amean += delta * count_inv;
                assqdm += delta * (ai - amean);
            }
        } else {
            if (aold == aold) {
                count--;
                count_inv = 1.0 / count;
                ddof_inv = 1.0 / (count - ddof);
                if (count > 0) {
                    delta = aold - amean;
                    amean -= delta * count_inv;
                    assqdm -= delta * (aold - amean);
                } else {
                    amean = 0;
                    assqdm = 0;
                }
            }
        }
        if (count >= min_count) {
            if (assqdm < 0) assqdm = 0;
            yi = assqdm * ddof_inv;
        } else {
            yi = NAN;
        }
        y[i] = yi;
    }
}
