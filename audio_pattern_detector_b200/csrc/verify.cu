// Step 2: gather-and-verify kernels for the candidate peaks of the selected units.
//
//  * normal / short-clip verifier: reference audio_pattern_detector.py:622-623 (centred slice,
//    renormalise), :773-804 (10-partition MSE), :808-846 + lib.rs:283-318, 651-675 (window-max
//    down-sampling + Pearson r, decision on the centre window).
//  * marker-tone verifier: reference audio_pattern_detector.py:642-750 and
//    detection_utils.py:41-142 (Hann-windowed full-length spectrum and 25 ms STFT frames of the
//    matched segment and both flanks, all float64).  The arbitrary-length DFT (L = 1223, 1827, ...)
//    is evaluated exactly as a Bluestein chirp-z transform over a power-of-two float64 FFT.
#include <algorithm>
#include <cmath>

#include "internal.h"
#include "peaks.h"

namespace apd {

__device__ __forceinline__ double block_sum(double v, double* red)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    return t;
}

__device__ __forceinline__ float block_max(float v, double* red)
{
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)v;
    __syncthreads();
    float t = 0.0f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t = fmaxf(t, (float)red[i]);
    return t;
}

__device__ __forceinline__ void init_record(apd_candidate& r, int chunk, int clip, int peak, int kind, float h)
{
    r.chunk = chunk; r.clip = clip; r.peak = peak; r.flags = kind << APD_KIND_SHIFT;
    r.height = h; r.similarity_whole = 0; r.similarity_middle = 0; r.reserved = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < 3; ++i) r.pearson[i] = nan;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 5; ++j) r.tone[i][j] = 0.0;
}

// ---------------------------------------------------------------------------
// normal + short clip
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_verify_normal(VerifyArgs A, int nslots)
{
    __shared__ double red[32];
    __shared__ double part[10];
    __shared__ float dsv[512];
    const int nsel = *A.pk.sel_count;
    for (int slot = blockIdx.y; slot < nslots && A.pk.slot0 + slot < nsel; slot += gridDim.y) {
    const int2 unit = A.pk.sel[A.pk.slot0 + slot];
    const int clip = unit.y;
    if (A.cv.strategy[clip] == APD_STRATEGY_MARKER_TONE && A.cv.tone_hz[clip] == A.cv.tone_hz[clip]) continue;   // not NaN
    long long start;
    int nsec;
    section_bounds(A.pk.geom[A.pk.clip_group[clip]], unit.x, start, nsec);
    const int L = A.pk.clip_len[clip];
    const int n = nsec + L - 1;
    const int W = 2 * L - 1;
    const float* __restrict__ q = A.pk.corr + (long long)slot * A.pk.corr_stride;
    const float* __restrict__ cc = A.cv.self_corr[clip];
    const int np = A.pk.n_peaks[slot];
    const int kind = A.cv.is_short[clip] ? 1 : 0;
    for (int p = blockIdx.x; p < np; p += gridDim.x) {
        const int pk = A.pk.peaks[(long long)slot * A.pk.peak_stride + p];
        apd_candidate* rec = A.slot_cands + (long long)slot * A.pk.peak_stride + p;
        const int half = W / 2;                                            // apd.py:534-535
        if (pk + half > n + 5 || pk - half < -5) {                         // apd.py:541-546
            if (threadIdx.x == 0) {
                init_record(*rec, A.chunk0 + unit.x, clip, pk, kind, A.pk.peak_height[(long long)slot * A.pk.peak_stride + p]);
                rec->flags |= APD_FLAG_SKIPPED;
            }
            continue;
        }
        const int beg = pk - (L - 1);                                      // au.py:177-191
        // max of the zero-padded slice
        // (loads are made unconditional through a clamped index so that several are in flight per thread: these
        //  loops run one CTA per peak and are otherwise bound by one L2 round trip per iteration)
        float m = 0.0f;
#pragma unroll 8
        for (int t = threadIdx.x; t < W; t += blockDim.x) {
            const int k = beg + t;
            const float raw = q[min(max(k, 0), n - 1)];
            if (k >= 0 && k < n) m = fmaxf(m, raw);
        }
        const float smax = block_max(m, red);
        // partition MSEs (float32 element ops as numpy, float64 accumulation), one partition at a time
        const int ps = W / 10;
        for (int i = 0; i < 10; ++i) {
            double acc = 0.0;
#pragma unroll 8
            for (int t = i * ps + threadIdx.x; t < (i + 1) * ps; t += blockDim.x) {
                const int k = beg + t;
                const float raw = q[min(max(k, 0), n - 1)];
                const float s = (k >= 0 && k < n) ? raw / smax : 0.0f / smax;
                const float d = cc[t] - s;
                acc += (double)(d * d);
            }
            const double s = block_sum(acc, red);
            if (threadIdx.x == 0) part[i] = s;
        }
        __syncthreads();
        float parts[10];
        float whole = 0.0f;
        for (int i = 0; i < 10; ++i) {
            parts[i] = (float)(part[i] / (double)ps);
            whole += parts[i];
        }
        whole = whole / 10.0f;                                             // apd.py:786
        const float middle = (parts[4] + parts[5]) / 2.0f;                 // apd.py:785
        const float sim = kind ? whole : fminf(whole, middle);             // apd.py:788-791
        double r[3] = {__longlong_as_double(0x7ff8000000000000LL), __longlong_as_double(0x7ff8000000000000LL),
                       __longlong_as_double(0x7ff8000000000000LL)};
        bool accept = false;
        if (!(sim > kMseLimit)) {                                          // apd.py:796
            const float* __restrict__ cache = A.cv.win_cache[clip];
            int cache_off = 0;
            for (int w = 0; w < 3; ++w) {
                const int ds = A.cv.win_ds[clip * 3 + w];
                if (ds <= 0) continue;
                const int lo = A.cv.win_lo[clip * 3 + w], hi = A.cv.win_hi[clip * 3 + w];
                const int nw = hi - lo;
                const double step = (double)nw / (double)ds;               // lib.rs:289
                __syncthreads();
                // one warp per output point, lanes striding through its window (coalesced)
                for (int i = threadIdx.x >> 5; i < ds; i += (int)(blockDim.x >> 5)) {
                    int a = (int)((double)i * step);
                    int b = (int)((double)(i + 1) * step);
                    if (b <= a) b = a + 1;
                    if (a >= nw) a = nw - 1;
                    if (b > nw) b = nw;
                    float mx = -INFINITY;
#pragma unroll 4
                    for (int t = a + (int)(threadIdx.x & 31); t < b; t += 32) {
                        const int k = beg + lo + t;
                        const float raw = q[min(max(k, 0), n - 1)];
                        mx = fmaxf(mx, (k >= 0 && k < n) ? raw : 0.0f);
                    }
                    mx = warp_max(mx);
                    if ((threadIdx.x & 31) == 0) dsv[i] = mx / smax;
                }
                __syncthreads();
                double sx = 0, sy = 0;
                for (int i = threadIdx.x; i < ds; i += blockDim.x) { sx += (double)cache[cache_off + i]; sy += (double)dsv[i]; }
                const double mx_ = block_sum(sx, red) / (double)ds;
                const double my_ = block_sum(sy, red) / (double)ds;
                double cov = 0, vx = 0, vy = 0;
                for (int i = threadIdx.x; i < ds; i += blockDim.x) {
                    const double dx = (double)cache[cache_off + i] - mx_, dy = (double)dsv[i] - my_;
                    cov += dx * dy; vx += dx * dx; vy += dy * dy;
                }
                cov = block_sum(cov, red);
                vx = block_sum(vx, red);
                vy = block_sum(vy, red);
                const double den = sqrt(vx * vy);
                r[w] = den == 0.0 ? 0.0 : cov / den;                       // lib.rs:670-674
                cache_off += ds;
            }
            const int center = kind ? 0 : 1;                               // apd.py:813,820
            accept = r[center] >= kPearsonMin;                             // apd.py:897
        }
        if (threadIdx.x == 0) {
            init_record(*rec, A.chunk0 + unit.x, clip, pk, kind, A.pk.peak_height[(long long)slot * A.pk.peak_stride + p]);
            rec->similarity_whole = whole;
            rec->similarity_middle = middle;
            for (int w = 0; w < 3; ++w) rec->pearson[w] = r[w];
            if (accept) rec->flags |= APD_FLAG_ACCEPT;
        }
        __syncthreads();
    }
    }
}

// ---------------------------------------------------------------------------
// marker tone
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 zmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// float64 FFT of length P (power of two) by one CTA: Stockham radix-4 passes (plus one radix-2 pass when
// log2 P is odd), ping-ponging between x and y in global memory (L2 resident); returns the buffer
// holding the result.  tw[t] = e^{-2 pi i t / P}, t < P/2.  INV conjugates the twiddles (unnormalised).
__device__ __forceinline__ double2 tw_at(const double2* __restrict__ tw, int idx, int half, bool inv)
{
    double2 w = idx < half ? tw[idx] : tw[idx - half];
    if (idx >= half) { w.x = -w.x; w.y = -w.y; }
    if (inv) w.y = -w.y;
    return w;
}

template <bool INV>
__device__ double2* fft64(double2* x, double2* y, const double2* __restrict__ tw, int P)
{
    const int half = P >> 1, quarter = P >> 2;
    int Ns = 1;
    while (Ns * 4 <= P) {
        const int tstep = P / (4 * Ns);
#pragma unroll 2
        for (int j = threadIdx.x; j < quarter; j += blockDim.x) {
            const int k = j & (Ns - 1);
            const double2 a0 = x[j];
            double2 a1 = x[j + quarter], a2 = x[j + 2 * quarter], a3 = x[j + 3 * quarter];
            if (Ns > 1) {
                a1 = zmul(a1, tw_at(tw, k * tstep, half, INV));
                a2 = zmul(a2, tw_at(tw, 2 * k * tstep, half, INV));
                a3 = zmul(a3, tw_at(tw, 3 * k * tstep, half, INV));
            }
            const double2 s02 = make_double2(a0.x + a2.x, a0.y + a2.y), d02 = make_double2(a0.x - a2.x, a0.y - a2.y);
            const double2 s13 = make_double2(a1.x + a3.x, a1.y + a3.y), d13 = make_double2(a1.x - a3.x, a1.y - a3.y);
            // (+-i) * d13 : forward multiplies by -i, inverse by +i
            const double2 r = INV ? make_double2(-d13.y, d13.x) : make_double2(d13.y, -d13.x);
            const int o = (j / Ns) * 4 * Ns + k;
            y[o] = make_double2(s02.x + s13.x, s02.y + s13.y);
            y[o + Ns] = make_double2(d02.x + r.x, d02.y + r.y);
            y[o + 2 * Ns] = make_double2(s02.x - s13.x, s02.y - s13.y);
            y[o + 3 * Ns] = make_double2(d02.x - r.x, d02.y - r.y);
        }
        __syncthreads();
        double2* t = x; x = y; y = t;
        Ns *= 4;
    }
    if (Ns < P) {                                   // one radix-2 pass left (Ns == P/2)
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            const double2 a = x[j];
            const double2 b = zmul(x[j + half], tw_at(tw, j, half, INV));     // k = j, tstep = 1
            y[j] = make_double2(a.x + b.x, a.y + b.y);
            y[j + Ns] = make_double2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
        double2* t = x; x = y; y = t;
    }
    return x;
}

__device__ __forceinline__ double2 chirp(long long n, int L, double sign)   // e^{sign * i pi n^2 / L}
{
    const long long r = (n * n) % (2LL * L);
    double s, c;
    sincospi(sign * (double)r / (double)L, &s, &c);
    return make_double2(c, s);
}

// Init-time tables of one tone clip: FFT_P of the wrapped chirp b[m] = e^{+i pi m^2 / L} (-L < m <= L/2),
// pre[n] = hann_L[n] * e^{-i pi n^2 / L} and post[k] = e^{-i pi k^2 / L} (k <= L/2).
__global__ void __launch_bounds__(1024)
k_tone_tables(int L, int P, const double2* __restrict__ tw, double2* buf0, double2* buf1, double2* chirp_fft,
              double2* pre, double2* post)
{
    for (int m = threadIdx.x; m < P; m += blockDim.x) {
        double2 v = make_double2(0, 0);
        if (m <= L / 2) v = chirp(m, L, +1.0);                     // b[m], m in [0, K), K = L/2 + 1 output bins
        else if (P - m < L) v = chirp(P - m, L, +1.0);             // b[-m'], m' in [1, L)
        buf0[m] = v;
    }
    for (int n = threadIdx.x; n < L; n += blockDim.x) {
        const double h = L > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)n / (double)(L - 1)) : 1.0;    // np.hanning
        const double2 c = chirp(n, L, -1.0);
        pre[n] = make_double2(h * c.x, h * c.y);
        if (n <= L / 2) post[n] = c;
    }
    __syncthreads();
    const double2* r = fft64<false>(buf0, buf1, tw, P);
    for (int m = threadIdx.x; m < P; m += blockDim.x) chirp_fft[m] = r[m];
}

struct ToneItem { int ci; int clip; int peak; float height; };

// Appends this round's tone-clip peaks to the batch-level work list (tone verification only needs
// the raw audio + gain, so it is deferred to one launch per batch).
__global__ void k_tone_collect(VerifyArgs A, int nslots, ToneItem* items, int* n_items, int capacity)
{
    if (blockIdx.x || threadIdx.x) return;
    int cnt = *n_items;
    for (int s = 0; s < nslots; ++s) {
        if (A.pk.slot0 + s >= *A.pk.sel_count) break;
        const int2 u = A.pk.sel[A.pk.slot0 + s];
        if (!(A.cv.strategy[u.y] == APD_STRATEGY_MARKER_TONE && A.cv.tone_hz[u.y] == A.cv.tone_hz[u.y])) continue;
        for (int p = 0; p < A.pk.n_peaks[s]; ++p) {
            if (cnt < capacity)
                items[cnt] = ToneItem{u.x, u.y, A.pk.peaks[(long long)s * A.pk.peak_stride + p],
                                      A.pk.peak_height[(long long)s * A.pk.peak_stride + p]};
            ++cnt;
        }
    }
    if (cnt > capacity) { atomicOr(A.pk.overflow, 2); cnt = capacity; }
    *n_items = cnt;
}

// ---------------------------------------------------------------------------
// Tone metrics for a round of up to R work items x 3 segments, as whole-GPU kernels:
//   prep -> forward FFT passes -> multiply by the chirp spectrum -> inverse FFT passes ->
//   spectrum statistics -> STFT frame bins -> per-frame statistics -> run-length summary.
// Each (item, segment) owns two double2 buffers A, B of `stride` elements (ping-pong).
// ---------------------------------------------------------------------------
struct ToneRound {
    const ToneItem* items;
    int i0, n_round;              // items [i0, i0 + n_round)
    double2* scratch;
    long long stride;
    double* stats;                // [(r*3+seg)*4] = {tot, band sum, detected Hz, -}
    double* metrics;              // [item][seg][5]
    int wl, hop;
    int seg0;                     // this pass handles segments seg0 + blockIdx.{y|x}
    const unsigned char* alive;   // flank passes: items whose matched segment failed are skipped (nullptr: none is)
    int frames;                   // 0: this pass computes no STFT frame metrics (they read as zero)
};

__device__ __forceinline__ bool tone_skip(const ToneRound& T, int r, int seg)
{
    return r >= T.n_round || (seg > 0 && T.alive && !T.alive[T.i0 + r]);
}

__device__ __forceinline__ double2* tone_buf(const ToneRound& T, int r, int seg, int which)
{
    return T.scratch + (((long long)r * 3 + seg) * 2 + which) * T.stride;
}

// Pass schedule of the batched power-of-two float64 FFT: log2 P = 4 a + b  ->  a radix-16 passes, then one radix-4
// pass if b >= 2, then one radix-2 pass if b is odd (P = 2^17: 16 16 16 16 2, five trips through memory instead of
// the nine of a radix-4 schedule; the passes are bound by the traffic of the float64 ping-pong buffers).
__host__ __device__ __forceinline__ int tone_npass_lg(int lg) { return (lg >> 2) + ((lg & 3) >= 2 ? 1 : 0) + (lg & 1); }
__device__ __forceinline__ int tone_npass(int P) { return tone_npass_lg(31 - __clz(P)); }

struct ToneSeg {
    const float* x;               // section start
    int nsec, L, P, ms, clip;
    double gain, f0;
};

__device__ __forceinline__ ToneSeg tone_seg(const VerifyArgs& A, const ToneItem& item, int seg)
{
    ToneSeg s;
    s.clip = item.clip;
    const int g = A.pk.clip_group[item.clip];
    long long start;
    section_bounds(A.pk.geom[g], item.ci, start, s.nsec);
    s.x = A.pk.geom[g].audio + (start - A.pk.geom[g].base);
    s.gain = A.gains[(long long)item.ci * A.n_groups + g];
    s.L = A.pk.clip_len[item.clip];
    s.P = A.cv.tone_P[item.clip];
    s.f0 = A.cv.tone_hz[item.clip];
    s.ms = item.peak - s.L + 1 + (seg == 1 ? -s.L : (seg == 2 ? s.L : 0));       // apd.py:650-653
    return s;
}

__device__ __forceinline__ double tone_sample(const ToneSeg& s, int k)
{
    return (k >= 0 && k < s.nsec) ? (double)normalize_sample(s.x[k], s.gain) : 0.0;
}

__global__ void __launch_bounds__(256)
k_tone_prep(VerifyArgs A, ToneRound T)
{
    const int r = blockIdx.z, seg = T.seg0 + blockIdx.y;
    if (tone_skip(T, r, seg)) return;
    const ToneSeg s = tone_seg(A, T.items[T.i0 + r], seg);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < s.P; n += gridDim.x * blockDim.x) {
        double2 v = make_double2(0, 0);
        if (n < s.L) {
            const double xv = tone_sample(s, s.ms + n);
            const double2 c = A.cv.tone_pre[s.clip][n];             // hann * e^{-i pi n^2 / L}
            v = make_double2(xv * c.x, xv * c.y);
        }
        tone_buf(T, r, seg, 0)[n] = v;
    }
}

// 16-point DFT in registers (4 x 4, constant inner twiddles), natural order in and out; INV conjugates.
template <bool INV>
__device__ __forceinline__ void zdft4(double2& a0, double2& a1, double2& a2, double2& a3)
{
    const double2 s02 = make_double2(a0.x + a2.x, a0.y + a2.y), d02 = make_double2(a0.x - a2.x, a0.y - a2.y);
    const double2 s13 = make_double2(a1.x + a3.x, a1.y + a3.y), d13 = make_double2(a1.x - a3.x, a1.y - a3.y);
    const double2 rr = INV ? make_double2(-d13.y, d13.x) : make_double2(d13.y, -d13.x);
    a0 = make_double2(s02.x + s13.x, s02.y + s13.y);
    a1 = make_double2(d02.x + rr.x, d02.y + rr.y);
    a2 = make_double2(s02.x - s13.x, s02.y - s13.y);
    a3 = make_double2(d02.x - rr.x, d02.y - rr.y);
}
template <bool INV>
__device__ __forceinline__ double2 zmul_w16(double2 a, double c, double sn)     // a * (c - i sn), conjugated for INV
{
    const double s_ = INV ? sn : -sn;
    return make_double2(a.x * c - a.y * s_, a.x * s_ + a.y * c);
}
template <bool INV>
__device__ __forceinline__ void zdft16(double2* v)
{
    constexpr double c1 = 0.9238795325112867, s1 = 0.3826834323650898, h = 0.7071067811865476;
    // r = 4 r1 + r2: DFT4 over r1 for every r2  ->  v[4 q1 + r2]
#pragma unroll
    for (int r2 = 0; r2 < 4; ++r2) zdft4<INV>(v[r2], v[4 + r2], v[8 + r2], v[12 + r2]);
    // inner twiddles w16^{q1 r2}
    v[5] = zmul_w16<INV>(v[5], c1, s1);
    v[6] = zmul_w16<INV>(v[6], h, h);
    v[7] = zmul_w16<INV>(v[7], s1, c1);
    v[9] = zmul_w16<INV>(v[9], h, h);
    v[10] = zmul_w16<INV>(v[10], 0.0, 1.0);
    v[11] = zmul_w16<INV>(v[11], -h, h);
    v[13] = zmul_w16<INV>(v[13], s1, c1);
    v[14] = zmul_w16<INV>(v[14], -h, h);
    v[15] = zmul_w16<INV>(v[15], -c1, -s1);
    // DFT4 over r2 for every q1  ->  X[q1 + 4 q2] at v[4 q1 + q2]
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) zdft4<INV>(v[4 * q1], v[4 * q1 + 1], v[4 * q1 + 2], v[4 * q1 + 3]);
}

// One pass slot of the batched float64 FFT (see fft64 for the single-CTA variant used at init).
template <bool INV>
__global__ void __launch_bounds__(256)
k_tone_fft_pass(VerifyArgs A, ToneRound T, int slot)
{
    const int r = blockIdx.z, seg = T.seg0 + blockIdx.y;
    if (tone_skip(T, r, seg)) return;
    const int clip = T.items[T.i0 + r].clip;
    const int P = A.cv.tone_P[clip];
    const int lg = 31 - __clz(P);
    const int n16 = lg >> 2, has4 = (lg & 3) >= 2 ? 1 : 0, np = n16 + has4 + (lg & 1);
    if (slot >= np) return;
    const int start_b = INV ? (np & 1) : 0;                        // the inverse starts where the forward ended
    const double2* __restrict__ x = tone_buf(T, r, seg, (start_b + slot) & 1);
    double2* __restrict__ y = tone_buf(T, r, seg, (start_b + slot + 1) & 1);
    const double2* __restrict__ tw = A.cv.tone_tw[clip];
    const int half = P >> 1, quarter = P >> 2;
    const int stride = gridDim.x * blockDim.x;
    if (slot < n16) {                                               // radix 16, Ns = 16^slot
        const int Ns = 1 << (4 * slot), T16 = P >> 4;
        const int tstep = P / (16 * Ns);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T16; j += stride) {
            const int k = j & (Ns - 1);
            double2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = x[j + q * T16];
            if (Ns > 1) {
                const double2 w1 = tw_at(tw, k * tstep, half, INV);
                double2 w = w1;
#pragma unroll
                for (int q = 1; q < 16; ++q) {
                    v[q] = zmul(v[q], w);
                    if (q < 15) w = zmul(w, w1);
                }
            }
            zdft16<INV>(v);
            const int o = (j / Ns) * 16 * Ns + k;
#pragma unroll
            for (int q1 = 0; q1 < 4; ++q1)
#pragma unroll
                for (int q2 = 0; q2 < 4; ++q2) y[o + (q1 + 4 * q2) * Ns] = v[4 * q1 + q2];
        }
    } else if (has4 && slot == n16) {                               // radix 4, Ns = 16^n16
        const int Ns = 1 << (4 * n16);
        const int tstep = P / (4 * Ns);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < quarter; j += stride) {
            const int k = j & (Ns - 1);
            double2 a0 = x[j], a1 = x[j + quarter], a2 = x[j + 2 * quarter], a3 = x[j + 3 * quarter];
            if (Ns > 1) {
                a1 = zmul(a1, tw_at(tw, k * tstep, half, INV));
                a2 = zmul(a2, tw_at(tw, 2 * k * tstep, half, INV));
                a3 = zmul(a3, tw_at(tw, 3 * k * tstep, half, INV));
            }
            zdft4<INV>(a0, a1, a2, a3);
            const int o = (j / Ns) * 4 * Ns + k;
            y[o] = a0;
            y[o + Ns] = a1;
            y[o + 2 * Ns] = a2;
            y[o + 3 * Ns] = a3;
        }
    } else {                                                        // the single radix-2 pass (Ns = P/2)
        for (int jj = blockIdx.x * blockDim.x + threadIdx.x; jj < half; jj += stride) {
            const double2 a = x[jj];
            const double2 b = zmul(x[jj + half], tw_at(tw, jj, half, INV));
            y[jj] = make_double2(a.x + b.x, a.y + b.y);
            y[jj + half] = make_double2(a.x - b.x, a.y - b.y);
        }
    }
}

__global__ void __launch_bounds__(256)
k_tone_mul(VerifyArgs A, ToneRound T)
{
    const int r = blockIdx.z, seg = T.seg0 + blockIdx.y;
    if (tone_skip(T, r, seg)) return;
    const int clip = T.items[T.i0 + r].clip;
    const int P = A.cv.tone_P[clip];
    double2* f = tone_buf(T, r, seg, tone_npass(P) & 1);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < P; n += gridDim.x * blockDim.x)
        f[n] = zmul(f[n], A.cv.tone_chirp_fft[clip][n]);
}

// One CTA per (segment, item): X_k = post[k] * conv[k] / P for k <= L/2 (conv ends in buffer A).
__global__ void __launch_bounds__(1024)
k_tone_stats(VerifyArgs A, ToneRound T)
{
    __shared__ double red[32];
    __shared__ int s_arg;
    const int r = blockIdx.y, seg = T.seg0 + blockIdx.x;
    if (tone_skip(T, r, seg)) return;
    const ToneSeg s = tone_seg(A, T.items[T.i0 + r], seg);
    const double2* __restrict__ c = tone_buf(T, r, seg, 0);
    const double2* __restrict__ post = A.cv.tone_post[s.clip];
    const double band = fmax(40.0, s.f0 * 0.08);                               // du.py:56
    const double d = __ddiv_rn(1.0, (double)A.sample_rate);
    const double fval = __ddiv_rn(1.0, __dmul_rn((double)s.L, d));            // np.fft.rfftfreq
    const int nb = s.L / 2 + 1;
    const double invP = 1.0 / (double)s.P;
    double tot = 0, bsum = 0, best = -1.0;
    int arg = 0x7fffffff;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        const double2 z = zmul(c[k], post[k]);
        const double re = z.x * invP, im = z.y * invP;
        const double m2 = re * re + im * im;
        tot += m2;
        if (fabs(__dmul_rn((double)k, fval) - s.f0) <= band) bsum += m2;      // du.py:74
        if (m2 > best) { best = m2; arg = k; }
    }
    tot = block_sum(tot, red);
    bsum = block_sum(bsum, red);
    double bm = best;                                                         // np.argmax: first maximum
    for (int o2 = 16; o2 > 0; o2 >>= 1) bm = fmax(bm, __shfl_xor_sync(0xffffffffu, bm, o2));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bm;
    if (threadIdx.x == 0) s_arg = 0x7fffffff;
    __syncthreads();
    double gm = -1.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) gm = fmax(gm, red[i]);
    if (best == gm && arg != 0x7fffffff) atomicMin(&s_arg, arg);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int argk = s_arg == 0x7fffffff ? 0 : s_arg;
        double* st = T.stats + ((long long)r * 3 + seg) * 4;
        st[0] = tot;
        st[1] = bsum;
        st[2] = __dmul_rn((double)argk, fval);                                // du.py:62
    }
}

// STFT frame bins (du.py:77-100): a CTA owns kFramesPerCta consecutive frames of one (item, segment), stages their
// windowed samples once in shared memory, then thread = (frame, bin) accumulates its DFT bin; |X|^2 into buffer B.
// For an even frame length the frame is folded first (one radix-2 step: even bins see x[n] + x[n + wl/2], odd bins
// x[n] - x[n + wl/2]), which halves the multiply-adds of every bin.
constexpr int kFramesPerCta = 8;

__global__ void __launch_bounds__(256)
k_tone_frames(VerifyArgs A, ToneRound T)
{
    extern __shared__ double2 ftw[];            // e^{-2 pi i t / wl} (t < wl), the frame Hann window (.x), then samples
    const int r = blockIdx.z, seg = T.seg0 + blockIdx.y;
    if (tone_skip(T, r, seg)) return;
    if (T.stats[((long long)r * 3 + seg) * 4] == 0.0) return;                 // du.py:65-72
    const ToneSeg s = tone_seg(A, T.items[T.i0 + r], seg);
    const int wl = T.wl, hop = T.hop;
    const int nf = s.L - wl > 0 ? (s.L - wl + hop - 1) / hop : 0;             // range(0, L - wl, hop)
    const int nbw = wl / 2 + 1;
    const int f0 = blockIdx.x * kFramesPerCta;
    if (f0 >= nf) return;
    double* xw = reinterpret_cast<double*>(ftw + 2 * wl);                     // [kFramesPerCta][wl]
    for (int t = threadIdx.x; t < wl; t += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)t / (double)wl, &sn, &cs);
        ftw[t] = make_double2(cs, sn);
        ftw[wl + t] = make_double2(wl > 1 ? 0.5 - 0.5 * cospi(2.0 * (double)t / (double)(wl - 1)) : 1.0, 0.0);
    }
    __syncthreads();
    const bool fold = (wl & 1) == 0;
    const int h = fold ? wl / 2 : wl;                                         // samples per bin after the fold
    if (fold) {
        for (int t = threadIdx.x; t < kFramesPerCta * h; t += blockDim.x) {
            const int fl = t / h, n = t - fl * h;
            if (f0 + fl >= nf) continue;
            const int k0 = s.ms + (f0 + fl) * hop + n;
            const double lo = tone_sample(s, k0) * ftw[wl + n].x, hi = tone_sample(s, k0 + h) * ftw[wl + n + h].x;
            xw[fl * wl + n] = lo + hi;
            xw[fl * wl + n + h] = lo - hi;
        }
    } else {
        for (int t = threadIdx.x; t < kFramesPerCta * wl; t += blockDim.x) {
            const int fl = t / wl, n = t - fl * wl;
            if (f0 + fl < nf) xw[t] = tone_sample(s, s.ms + (f0 + fl) * hop + n) * ftw[wl + n].x;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kFramesPerCta * nbw; t += blockDim.x) {
        const int fl = t / nbw, kb = t - fl * nbw;
        if (f0 + fl >= nf) continue;
        const double* __restrict__ x = xw + fl * wl + ((fold && (kb & 1)) ? h : 0);
        double re = 0, im = 0;
        int ph = 0;
        for (int n = 0; n < h; ++n) {
            const double2 e = ftw[ph];
            re += x[n] * e.x;
            im += x[n] * e.y;
            ph += kb;
            if (ph >= wl) ph -= wl;
        }
        ((double*)tone_buf(T, r, seg, 1))[(long long)(f0 + fl) * nbw + kb] = re * re + im * im;
    }
}

// Thread per frame: energy, dominant bin, band purity, "active" flag (du.py:89-105) into buffer A.
__global__ void __launch_bounds__(128)
k_tone_frame_stats(VerifyArgs A, ToneRound T)
{
    const int r = blockIdx.z, seg = T.seg0 + blockIdx.y;
    if (tone_skip(T, r, seg)) return;
    if (T.stats[((long long)r * 3 + seg) * 4] == 0.0) return;
    const ToneSeg s = tone_seg(A, T.items[T.i0 + r], seg);
    const int wl = T.wl, hop = T.hop;
    const int nf = s.L - wl > 0 ? (s.L - wl + hop - 1) / hop : 0;
    const int nbw = wl / 2 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const double* __restrict__ fm2 = (const double*)tone_buf(T, r, seg, 1);
    double* fpur = (double*)tone_buf(T, r, seg, 0);
    unsigned char* fact = (unsigned char*)(fpur + nf);
    const double band = fmax(40.0, s.f0 * 0.08), lock = fmax(20.0, s.f0 * 0.04);   // du.py:56-57
    const double d = __ddiv_rn(1.0, (double)A.sample_rate);
    const double fvalw = __ddiv_rn(1.0, __dmul_rn((double)wl, d));
    double e = 0, bs = 0, bm = -1.0;
    int ak = 0;
    for (int kb = 0; kb < nbw; ++kb) {
        const double m2 = fm2[(long long)f * nbw + kb];
        e += m2;
        if (fabs(__dmul_rn((double)kb, fvalw) - s.f0) <= band) bs += m2;
        if (m2 > bm) { bm = m2; ak = kb; }
    }
    if (e == 0.0) { fpur[f] = -1.0; fact[f] = 0; return; }
    const double fdom = __dmul_rn((double)ak, fvalw);
    const double pur = bs / e;
    const double tol = fmax(1e-9 * fmax(fabs(fdom), fabs(s.f0)), lock);       // math.isclose(abs_tol=lock)
    fpur[f] = pur;
    fact[f] = (fabs(fdom - s.f0) <= tol && pur >= 0.55) ? 1 : 0;              // du.py:102-105
}

// Warp per (item, segment): run-length summary over frames (du.py:106-117) and the 5 metrics.  Each lane summarises a
// contiguous slice of the frames (count, active count, purity sum, longest run, leading / trailing run, all-active
// flag); lane 0 chains the 32 slices.
__global__ void __launch_bounds__(128)
k_tone_final(VerifyArgs A, ToneRound T, int nseg)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= T.n_round * nseg) return;
    const int r = w / nseg, seg = T.seg0 + w % nseg;
    if (tone_skip(T, r, seg)) return;
    const int t = r * 3 + seg;
    const ToneSeg s = tone_seg(A, T.items[T.i0 + r], seg);
    const double* st = T.stats + (long long)t * 4;
    const double tot = st[0];
    const int wl = T.wl, hop = T.hop;
    const int nf = s.L - wl > 0 ? (s.L - wl + hop - 1) / hop : 0;
    const double* fpur = (const double*)tone_buf(T, r, seg, 0);
    const unsigned char* fact = (const unsigned char*)(fpur + nf);
    double ratio = 0, meanp = 0;
    int longest = 0;
    if (tot != 0.0 && T.frames) {                                              // warp-uniform
        const int per = (nf + 31) / 32, f0 = lane * per, f1 = min(nf, f0 + per);
        int frames = 0, active = 0, run = 0, best = 0, lead = 0;
        bool all_active = true;                                                // every frame of the slice extends a run
        double psum = 0;
        for (int f = f0; f < f1; ++f) {
            const double pu = fpur[f];
            const bool on = pu >= 0.0 && fact[f];
            if (pu >= 0.0) ++frames;
            if (on) { ++active; ++run; best = max(best, run); psum += pu; }
            else { if (all_active) lead = run; all_active = false; run = 0; }
        }
        if (all_active) lead = run;
        const int trail = run;
        const int len = max(f1 - f0, 0);
        // totals
        for (int o = 16; o > 0; o >>= 1) {
            frames += __shfl_xor_sync(0xffffffffu, frames, o);
            active += __shfl_xor_sync(0xffffffffu, active, o);
            psum += __shfl_xor_sync(0xffffffffu, psum, o);
        }
        // longest run across slice boundaries: lane 0 walks the slices in order
        int carry = 0;
        longest = 0;
        for (int l = 0; l < 32; ++l) {
            const int b = __shfl_sync(0xffffffffu, best, l), ld = __shfl_sync(0xffffffffu, lead, l);
            const int tr = __shfl_sync(0xffffffffu, trail, l), ln = __shfl_sync(0xffffffffu, len, l);
            const int aa = __shfl_sync(0xffffffffu, all_active ? 1 : 0, l);
            if (ln == 0) continue;
            longest = max(longest, max(b, carry + ld));
            carry = aa ? carry + ln : tr;
        }
        ratio = frames > 0 ? (double)active / (double)frames : 0.0;
        meanp = active > 0 ? psum / (double)active : 0.0;
    }
    if (lane == 0) {
        double* out = T.metrics + ((long long)(T.i0 + r) * 3 + seg) * 5;
        out[0] = st[2];
        out[1] = tot != 0.0 ? st[1] / tot : 0.0;                               // du.py:64-75
        out[2] = ratio;
        out[3] = (double)longest;
        out[4] = meanp;
    }
}

// The conditions of apd.py:707-724 that only look at the matched segment (frequency within 5 %, band purity, active
// frame ratio, longest run, mean active purity); the flank purities complete the decision.
__device__ __forceinline__ bool tone_match_ok(const double* m, const double* th, double f0)
{
    bool ok = fabs(m[0] - f0) <= fmax(0.05 * fmax(fabs(m[0]), fabs(f0)), 0.0);   // apd.py:707
    return ok && m[1] >= th[0] && m[2] >= th[1] && (long long)m[3] >= (long long)th[2] && m[4] >= th[3];
}

// After the matched-segment pass of a round: which of its items still need their flanks.  A candidate whose matched
// segment fails is rejected whatever its flanks are (the decision is a conjunction), so unless every metric was asked
// for (trace / single-candidate calls) the two flank transforms of such items are not computed and read as zero.
__global__ void k_tone_gate(VerifyArgs A, ToneRound T, unsigned char* alive)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= T.n_round) return;
    const int clip = T.items[T.i0 + r].clip;
    alive[T.i0 + r] = tone_match_ok(T.metrics + (long long)(T.i0 + r) * 15, A.cv.tone_thr + clip * 6, A.cv.tone_hz[clip]) ? 1 : 0;
}

// Decision (apd.py:707-724) and record emission for the deferred tone candidates.
__global__ void k_tone_decide(VerifyArgs A, const ToneItem* __restrict__ items, const int* __restrict__ n_items,
                              const double* __restrict__ metrics)
{
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_items) return;
    const ToneItem item = items[it];
    const int clip = item.clip;
    long long start;
    int nsec;
    section_bounds(A.pk.geom[A.pk.clip_group[clip]], item.ci, start, nsec);
    const int L = A.pk.clip_len[clip];
    const int n = nsec + L - 1, half = (2 * L - 1) / 2;
    apd_candidate rec;
    init_record(rec, A.chunk0 + item.ci, clip, item.peak, 2, item.height);
    if (item.peak + half > n + 5 || item.peak - half < -5) {
        rec.flags |= APD_FLAG_SKIPPED;
    } else {
        const double* m = metrics + (long long)it * 15;
        for (int s = 0; s < 3; ++s) for (int j = 0; j < 5; ++j) rec.tone[s][j] = m[s * 5 + j];
        const double* th = A.cv.tone_thr + clip * 6;
        const double f0 = A.cv.tone_hz[clip];
        const double lo = fmin(m[5 + 1], m[10 + 1]), hi = fmax(m[5 + 1], m[10 + 1]);
        const bool ok = tone_match_ok(m, th, f0) && lo <= th[4] && hi <= th[5];       // apd.py:717-724
        if (ok) rec.flags |= APD_FLAG_ACCEPT;
    }
    const int o = atomicAdd(A.out_count, 1);
    if (o < A.out_capacity) A.out[o] = rec;
    else atomicOr(A.pk.overflow, 4);
}

__device__ __forceinline__ bool is_tone_slot(const VerifyArgs& A, int s)
{
    const int clip = A.pk.sel[A.pk.slot0 + s].y;
    return A.cv.strategy[clip] == APD_STRATEGY_MARKER_TONE && A.cv.tone_hz[clip] == A.cv.tone_hz[clip];
}

// Ordered gather of the per-slot records (normal / short clips) into the output list: one thread per slot
// reads its count, block-wide exclusive scan, then one warp per slot copies its records.
__global__ void __launch_bounds__(1024)
k_emit(VerifyArgs A, int nslots)
{
    __shared__ int s_off[kMaxSlots + 1];
    __shared__ int s_warp[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nsel = *A.pk.sel_count;
    const int base = *A.out_count;
    int cnt = 0;
    if (t < nslots && A.pk.slot0 + t < nsel && !is_tone_slot(A, t)) cnt = A.pk.n_peaks[t];
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += s_warp[k];
    if (t < nslots) s_off[t] = base + woff + inc - cnt;
    if (t == nslots - 1) s_off[nslots] = base + woff + inc;
    __syncthreads();
    for (int s = w; s < nslots; s += (int)(blockDim.x >> 5)) {
        const int np = s_off[s + 1] - s_off[s];
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(A.slot_cands + (long long)s * A.pk.peak_stride);
        constexpr int Q = (int)(sizeof(apd_candidate) / sizeof(uint4));
        for (int i = lane; i < np * Q; i += 32) {
            const int o = s_off[s] + i / Q;
            if (o < A.out_capacity) reinterpret_cast<uint4*>(A.out + o)[i % Q] = src[i];
        }
    }
    if (t == 0) {
        if (s_off[nslots] > A.out_capacity) atomicOr(A.pk.overflow, 4);
        *A.out_count = s_off[nslots] < A.out_capacity ? s_off[nslots] : A.out_capacity;
    }
}

void launch_verify(const VerifyArgs& A, int nslots, cudaStream_t st, long long* launches)
{
    if (nslots <= 0) return;
    // x: peaks of a unit in parallel (short clips can keep dozens; the kernel strides); y: slots, strided (most
    // slots are unused).  1024 threads: a 10 s clip's slice is 160 k samples per pass.  Smaller CTA footprints
    // (1024 x 32 registers, 512 x 40, 256 x 63) and a 4x smaller grid were measured: no change of the step time
    // (DESIGN.md section 6).
    dim3 gn(32, std::min(nslots, 128));
    k_verify_normal<<<gn, 1024, 0, st>>>(A, nslots);
    ++*launches;
}

void launch_tone_collect(const VerifyArgs& A, int nslots, void* items, int* n_items, int item_capacity,
                         cudaStream_t st, long long* launches)
{
    if (nslots <= 0) return;
    k_tone_collect<<<1, 32, 0, st>>>(A, nslots, (ToneItem*)items, n_items, item_capacity);
    ++*launches;
}

void launch_tone_batch(const VerifyArgs& A, void* items, int* n_items_dev, int n_items, double* metrics,
                       double* stats, int round_items, int max_P, int max_L, int wl, bool all_segments,
                       unsigned char* alive, cudaStream_t st, long long* launches)
{
    if (n_items <= 0 || round_items <= 0) return;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tone_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr = true;
    }
    int lg = 0;
    while ((1 << lg) < max_P) ++lg;
    int max_pass = 0;                                              // a shorter transform may need more passes
    for (int l = 1; l <= lg; ++l) max_pass = std::max(max_pass, tone_npass_lg(l));
    const int hop = wl / 2 > 1 ? wl / 2 : 1;
    const int nf_max = max_L - wl > 0 ? (max_L - wl + hop - 1) / hop : 0;
    const long long work_max = (long long)nf_max * (wl / 2 + 1);
    // skipped flank segments read as zero
    cudaMemsetAsync(metrics, 0, sizeof(double) * 15 * (size_t)n_items, st);
    for (int i0 = 0; i0 < n_items; i0 += round_items) {
        ToneRound T{(const ToneItem*)items, i0, std::min(round_items, n_items - i0), A.tone_scratch,
                    A.tone_scratch_stride, stats, metrics, wl, hop, 0, nullptr, 1};
        const unsigned R = (unsigned)T.n_round;
        // grid-stride in x (capping the grid at 32 CTAs per transform was measured: no gain for the correlate stage
        // that shares the GPU, longer phase 2)
        const unsigned gx = (unsigned)((max_P / 4 + 255) / 256);
        // pass 0: the matched segment of every item; pass 1: the two flanks of the items that still need them
        // timing experiments only (results are wrong with parts switched off): bit 0 spectrum, 1 frames, 2 summary
        static const int parts = getenv("APD_B200_TONE_PARTS") ? atoi(getenv("APD_B200_TONE_PARTS")) : 7;
        for (int pass = 0; pass < 2; ++pass) {
            const unsigned ns = pass == 0 ? 1u : 2u;
            T.seg0 = pass;
            T.alive = (pass == 1 && !all_segments) ? alive : nullptr;
            // the decision reads only the band purity of the flanks (apd.py:716-724): their frame metrics are
            // computed only when every metric was asked for
            T.frames = (pass == 0 || all_segments) ? 1 : 0;
            if (parts & 1) {
            k_tone_prep<<<dim3(gx, ns, R), 256, 0, st>>>(A, T);
            for (int s = 0; s < max_pass; ++s)
                k_tone_fft_pass<false><<<dim3(gx, ns, R), 256, 0, st>>>(A, T, s);
            k_tone_mul<<<dim3(gx, ns, R), 256, 0, st>>>(A, T);
            for (int s = 0; s < max_pass; ++s)
                k_tone_fft_pass<true><<<dim3(gx, ns, R), 256, 0, st>>>(A, T, s);
            k_tone_stats<<<dim3(ns, R), 1024, 0, st>>>(A, T);
            }
            if (work_max > 0 && T.frames && (parts & 2)) {
                k_tone_frames<<<dim3((unsigned)((nf_max + kFramesPerCta - 1) / kFramesPerCta), ns, R), 256,
                                (size_t)2 * wl * sizeof(double2) + (size_t)kFramesPerCta * wl * sizeof(double), st>>>(A, T);
                k_tone_frame_stats<<<dim3((nf_max + 127) / 128, ns, R), 128, 0, st>>>(A, T);
            }
            if (parts & 4) k_tone_final<<<(T.n_round * (int)ns + 3) / 4, 128, 0, st>>>(A, T, (int)ns);
            *launches += 5 + 2 * max_pass + (work_max > 0 && T.frames ? 2 : 0);
            if (pass == 0 && !all_segments) {
                k_tone_gate<<<(T.n_round + 63) / 64, 64, 0, st>>>(A, T, alive);
                ++*launches;
            }
        }
    }
    k_tone_decide<<<(n_items + 127) / 128, 128, 0, st>>>(A, (const ToneItem*)items, n_items_dev, metrics);
    ++*launches;
}

void launch_emit(const VerifyArgs& A, int nslots, cudaStream_t st, long long* launches)
{
    if (nslots <= 0) return;
    k_emit<<<1, 1024, 0, st>>>(A, nslots);     // nslots <= kMaxSlots == 1024
    ++*launches;
}

size_t tone_item_bytes() { return sizeof(ToneItem); }

void launch_tone_tables(int L, int P, const double2* tw, double2* buf0, double2* buf1, double2* chirp_fft,
                        double2* pre, double2* post, cudaStream_t st)
{
    k_tone_tables<<<1, 1024, 0, st>>>(L, P, tw, buf0, buf1, chirp_fft, pre, post);
}

}  // namespace apd
