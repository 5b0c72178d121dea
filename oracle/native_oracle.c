/*
 * oracle/native_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar, single-threaded) of the numeric leaves the
 * reference's detection hot path calls through its Rust `_native` module.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (audio_pattern_detector_b200/) never does.
 *
 * Each function cites the reference location whose behaviour it restates
 * (paths are relative to /root/reference/native-helper/src/).  The reference
 * is Rust; this is an independent C formulation of the same published
 * algorithms (ITU-R BS.1770-4 gating, scipy-style find_peaks, window-max
 * resampling, two-pass Pearson r).
 *
 * Pinned by tests/test_oracle_native.py against the reference's own known
 * answers (lib.rs:683-1173, native-helper/tests/test_python_bindings.py).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* BS.1770 K-weighting: lib.rs:13-53                                   */
/* coef layout: [b_shelf(3) | a_shelf(3) | b_hpass(3) | a_hpass(3)]    */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_k_weighting(double rate, double coef[12])
{
    /* stage 1: high shelf, +4 dB, Q = 1/sqrt(2), 1500 Hz */
    const double A = pow(10.0, 4.0 / 40.0);
    const double w = 2.0 * M_PI * (1500.0 / rate);
    const double alpha = sin(w) / (2.0 * M_SQRT1_2);
    const double c = cos(w);
    const double k = 2.0 * sqrt(A) * alpha;

    const double sb0 = A * ((A + 1.0) + (A - 1.0) * c + k);
    const double sb1 = -2.0 * A * ((A - 1.0) + (A + 1.0) * c);
    const double sb2 = A * ((A + 1.0) + (A - 1.0) * c - k);
    const double sa0 = (A + 1.0) - (A - 1.0) * c + k;
    const double sa1 = 2.0 * ((A - 1.0) - (A + 1.0) * c);
    const double sa2 = (A + 1.0) - (A - 1.0) * c - k;
    coef[0] = sb0 / sa0; coef[1] = sb1 / sa0; coef[2] = sb2 / sa0;
    coef[3] = 1.0;       coef[4] = sa1 / sa0; coef[5] = sa2 / sa0;

    /* stage 2: high pass, Q = 0.5, 38 Hz */
    const double w2 = 2.0 * M_PI * (38.0 / rate);
    const double alpha2 = sin(w2) / (2.0 * 0.5);
    const double c2 = cos(w2);
    const double ha0 = 1.0 + alpha2;
    coef[6] = ((1.0 + c2) / 2.0) / ha0;
    coef[7] = (-(1.0 + c2)) / ha0;
    coef[8] = ((1.0 + c2) / 2.0) / ha0;
    coef[9] = 1.0;
    coef[10] = (-2.0 * c2) / ha0;
    coef[11] = (1.0 - alpha2) / ha0;
}

/* Prefix sum of the squared K-weighted signal: lib.rs:75-110.
 * Two cascaded transposed-direct-form-II biquads, zero initial state, f64. */
static double *kweighted_energy_prefix(const float *x, size_t n, const double *cf)
{
    double *P = (double *)malloc((n + 1) * sizeof(double));
    if (!P) return NULL;
    double s1 = 0.0, s2 = 0.0, h1 = 0.0, h2 = 0.0;
    P[0] = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double in = (double)x[i];
        const double u = cf[0] * in + s1;          /* shelf output */
        s1 = cf[1] * in - cf[4] * u + s2;
        s2 = cf[2] * in - cf[5] * u;
        const double v = cf[6] * u + h1;           /* high-pass output */
        h1 = cf[7] * u - cf[10] * v + h2;
        h2 = cf[8] * u - cf[11] * v;
        P[i + 1] = P[i] + v * v;
    }
    return P;
}

/* Gated integrated loudness: lib.rs:113-122 (block bounds, truncating casts),
 * lib.rs:128-214 (absolute gate -70 LUFS, relative gate -10 LU). */
ORACLE_API double oracle_integrated_loudness(const float *x, size_t n, uint32_t sample_rate,
                                             double block_size)
{
    const double OFFSET = -0.691, ABS_GATE = -70.0;
    if (n == 0) return -INFINITY;
    const double rate = (double)sample_rate;
    double cf[12];
    oracle_k_weighting(rate, cf);
    double *P = kweighted_energy_prefix(x, n, cf);
    if (!P) return NAN;

    const double step = 1.0 - 0.75;
    const double win = block_size * rate;
    const double hop = win * step;
    const double T = (double)n / rate;
    /* Rust f64::round == C round(): half away from zero (lib.rs:149) */
    const long long nb = (long long)round((T - block_size) / (block_size * step)) + 1;
    double result;
    if (nb <= 0) {
        const double ms = P[n] / (double)n;
        result = (ms <= 0.0) ? -INFINITY : OFFSET + 10.0 * log10(ms);
        free(P);
        return result;
    }

    double gate = -INFINITY;  /* pass 0: absolute gate only; pass 1: + relative gate */
    result = -INFINITY;
    for (int pass = 0; pass < 2; ++pass) {
        double sum = 0.0;
        size_t cnt = 0;
        for (long long j = 0; j < nb; ++j) {
            size_t lo = (size_t)((double)j * hop);
            size_t hi = (size_t)((double)j * hop + win);
            if (hi > n) hi = n;
            if (lo >= hi) continue;
            const double ms = (P[hi] - P[lo]) / (double)(hi - lo);
            if (ms <= 0.0) continue;
            const double l = OFFSET + 10.0 * log10(ms);
            const int keep = (pass == 0) ? (l >= ABS_GATE) : (l > gate && l >= ABS_GATE);
            if (keep) { sum += ms; ++cnt; }
        }
        if (cnt == 0) { result = -INFINITY; break; }
        const double mean = sum / (double)cnt;
        if (pass == 0) gate = OFFSET + 10.0 * log10(mean) - 10.0;
        else result = OFFSET + 10.0 * log10(mean);
    }
    free(P);
    return result;
}

/* Gain + hard clip: lib.rs:220-227.  Rust f64::clamp keeps NaN (0*inf for
 * silence); reproduce that explicitly, fmin/fmax would drop it. */
ORACLE_API void oracle_loudness_normalize(const float *x, size_t n, double current_lufs,
                                          double target_lufs, float *out)
{
    const double gain = pow(10.0, (target_lufs - current_lufs) / 20.0);
    for (size_t i = 0; i < n; ++i) {
        double v = (double)x[i] * gain;
        if (v == v) {           /* not NaN */
            if (v < -1.0) v = -1.0;
            if (v > 1.0) v = 1.0;
        }
        out[i] = (float)v;
    }
}

/* ------------------------------------------------------------------ */
/* Window-max resampling: lib.rs:283-318                               */
/* returns number of outputs written (0 if n == 0 or target == 0)       */
/* ------------------------------------------------------------------ */
ORACLE_API size_t oracle_resample_preserve_maxima(const float *x, size_t n, size_t target,
                                                  float *out)
{
    if (n == 0 || target == 0) return 0;
    const double step = (double)n / (double)target;
    for (size_t i = 0; i < target; ++i) {
        size_t a = (size_t)((double)i * step);
        size_t b = (size_t)((double)(i + 1) * step);
        if (b <= a) b = a + 1;
        if (a >= n) a = n - 1;
        if (b > n) b = n;
        float m = x[a];
        for (size_t j = a + 1; j < b; ++j) {
            /* Rust f32::max: if one operand is NaN return the other */
            const float v = x[j];
            if (m != m) m = v;
            else if (v == v && v > m) m = v;
        }
        out[i] = m;
    }
    return target;
}

/* ------------------------------------------------------------------ */
/* find_peaks(height, distance): lib.rs:380-396, 404-485               */
/* ------------------------------------------------------------------ */
typedef struct { float v; int64_t pos; int64_t rank; } pk_t;

static int by_height_then_index(const void *pa, const void *pb)
{
    const pk_t *a = (const pk_t *)pa, *b = (const pk_t *)pb;
    if (a->v > b->v) return -1;       /* taller first */
    if (a->v < b->v) return 1;
    return (a->rank > b->rank) - (a->rank < b->rank);   /* ties: lower index first (lib.rs:446-451) */
}

/* Strict local maxima with plateau midpoint (lib.rs:404-428).
 * Writes up to cap indices, returns the count found. */
ORACLE_API int64_t oracle_local_maxima(const float *x, int64_t n, int64_t *out, int64_t cap)
{
    int64_t cnt = 0;
    if (n < 3) return 0;
    int64_t i = 1;
    while (i < n - 1) {
        if (x[i - 1] < x[i]) {
            const int64_t left = i;
            while (i + 1 < n && x[i] == x[i + 1]) ++i;
            if (i + 1 < n && x[i] > x[i + 1]) {
                if (cnt < cap) out[cnt] = (left + i) / 2;
                ++cnt;
            }
        }
        ++i;
    }
    return cnt;
}

/* use_height / use_distance select the optional filters.  `out` must hold at
 * least n/2+1 entries.  Returns the number of peaks (sorted ascending). */
ORACLE_API int64_t oracle_find_peaks(const float *x, int64_t n, int use_height, float height,
                                     int use_distance, int64_t distance, int64_t *out)
{
    const int64_t cap = n / 2 + 1;
    int64_t m = oracle_local_maxima(x, n, out, cap);
    if (use_height) {                           /* lib.rs:431-433 */
        int64_t w = 0;
        for (int64_t r = 0; r < m; ++r)
            if (x[out[r]] >= height) out[w++] = out[r];
        m = w;
    }
    if (use_distance && m > 0 && distance > 0) {   /* lib.rs:437-485 */
        pk_t *order = (pk_t *)malloc((size_t)m * sizeof(pk_t));
        unsigned char *keep = (unsigned char *)malloc((size_t)m);
        if (!order || !keep) { free(order); free(keep); return -1; }
        for (int64_t r = 0; r < m; ++r) {
            order[r].v = x[out[r]]; order[r].pos = out[r]; order[r].rank = r;
            keep[r] = 1;
        }
        qsort(order, (size_t)m, sizeof(pk_t), by_height_then_index);
        for (int64_t q = 0; q < m; ++q) {
            const int64_t r = order[q].rank;
            if (!keep[r]) continue;
            for (int64_t j = r - 1; j >= 0 && out[r] - out[j] < distance; --j) keep[j] = 0;
            for (int64_t j = r + 1; j < m && out[j] - out[r] < distance; ++j) keep[j] = 0;
        }
        int64_t w = 0;
        for (int64_t r = 0; r < m; ++r)
            if (keep[r]) out[w++] = out[r];
        m = w;
        free(order);
        free(keep);
    }
    return m;
}

/* ------------------------------------------------------------------ */
/* Pearson r, two pass, f64: lib.rs:651-675                            */
/* ------------------------------------------------------------------ */
ORACLE_API double oracle_pearson(const float *x, const float *y, size_t n)
{
    if (n == 0) return 0.0;
    double sx = 0.0, sy = 0.0;
    for (size_t i = 0; i < n; ++i) sx += (double)x[i];
    for (size_t i = 0; i < n; ++i) sy += (double)y[i];
    const double mx = sx / (double)n, my = sy / (double)n;
    double cov = 0.0, vx = 0.0, vy = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double dx = (double)x[i] - mx, dy = (double)y[i] - my;
        cov += dx * dy; vx += dx * dx; vy += dy * dy;
    }
    const double den = sqrt(vx * vy);
    return den == 0.0 ? 0.0 : cov / den;
}
