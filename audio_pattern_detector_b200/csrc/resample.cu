// FFT resampler on the device.
//
// Replaces _native.resample (native-helper/src/python.rs:106-116 -> resample_1d, native-helper/src/lib.rs:235-275),
// which the reference calls once per chunk read in _WavFileStreamWrapper.read (match.py:395-425) through
// audio_utils.resample_audio (audio_utils.py:154-171): complex float64 FFT of the n input samples, copy of the
// (N+1)/2 lowest positive and (N-1)/2 lowest negative bins (N = min(n, m); the Nyquist bin of an even N is
// dropped, unlike scipy), inverse FFT of length m, scale 1/n, round to float32.
//
// Arbitrary n and m: 7-smooth lengths run as mixed-radix Stockham passes (one butterfly of radix
// 25/16/8/7/5/4/3/2 per thread, float64, ping-pong through HBM, natural order out); any other length goes
// through Bluestein's chirp-z with a 7-smooth inner length >= 2*len-1.  The first forward pass reads the float32
// input, the first inverse pass reads the forward spectrum through the bin map and the last inverse pass writes
// the scaled float32 output, so no stand-alone convert / copy kernels run on the smooth path.  `batch`
// independent transforms of the same (n, m) run in the same launches (grid.y).
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/apd_b200.h"

namespace {

typedef double2 cpx;

__constant__ double kCos3[3] = {1.0, -0.5, -0.5};
__constant__ double kSin3[3] = {0.0, 0.8660254037844386, -0.8660254037844386};
__constant__ double kCos5[5] = {1.0, 0.30901699437494745, -0.8090169943749475, -0.8090169943749475, 0.30901699437494745};
__constant__ double kSin5[5] = {0.0, 0.9510565162951535, 0.5877852522924731, -0.5877852522924731, -0.9510565162951535};
__constant__ double kCos7[7] = {1.0, 0.6234898018587335, -0.2225209339563144, -0.9009688679024191,
                                -0.9009688679024191, -0.2225209339563144, 0.6234898018587335};
__constant__ double kSin7[7] = {0.0, 0.7818314824680298, 0.9749279121818236, 0.4338837391175581,
                                -0.4338837391175581, -0.9749279121818236, -0.7818314824680298};
__constant__ double kCos16[16] = {1.0, 0.9238795325112867, 0.7071067811865476, 0.3826834323650898, 0.0,
                                  -0.3826834323650898, -0.7071067811865476, -0.9238795325112867, -1.0,
                                  -0.9238795325112867, -0.7071067811865476, -0.3826834323650898, 0.0,
                                  0.3826834323650898, 0.7071067811865476, 0.9238795325112867};
__constant__ double kSin16[16] = {0.0, 0.3826834323650898, 0.7071067811865476, 0.9238795325112867, 1.0,
                                  0.9238795325112867, 0.7071067811865476, 0.3826834323650898, 0.0,
                                  -0.3826834323650898, -0.7071067811865476, -0.9238795325112867, -1.0,
                                  -0.9238795325112867, -0.7071067811865476, -0.3826834323650898};
__constant__ double kCos25[25] = {1.0, 0.9685831611286311, 0.8763066800438636, 0.7289686274214116, 0.5358267949789967,
                                  0.30901699437494745, 0.06279051952931337, -0.18738131458572463, -0.42577929156507266,
                                  -0.6374239897486897, -0.8090169943749475, -0.9297764858882515, -0.9921147013144779,
                                  -0.9921147013144779, -0.9297764858882515, -0.8090169943749475, -0.6374239897486897,
                                  -0.42577929156507266, -0.18738131458572463, 0.06279051952931337, 0.30901699437494745,
                                  0.5358267949789967, 0.7289686274214116, 0.8763066800438636, 0.9685831611286311};
__constant__ double kSin25[25] = {0.0, 0.2486898871648548, 0.48175367410171527, 0.6845471059286887, 0.8443279255020151,
                                  0.9510565162951535, 0.9980267284282716, 0.9822872507286887, 0.9048270524660196,
                                  0.7705132427757893, 0.5877852522924731, 0.368124552684678, 0.12533323356430426,
                                  -0.12533323356430426, -0.368124552684678, -0.5877852522924731, -0.7705132427757893,
                                  -0.9048270524660196, -0.9822872507286887, -0.9980267284282716, -0.9510565162951535,
                                  -0.8443279255020151, -0.6845471059286887, -0.48175367410171527, -0.2486898871648548};

// e^{2*pi*i*k/R} for the table radices, k a compile-time constant after unrolling
template <int R>
__device__ __forceinline__ cpx root(int k)
{
    if constexpr (R == 3) return make_double2(kCos3[k], kSin3[k]);
    else if constexpr (R == 5) return make_double2(kCos5[k], kSin5[k]);
    else if constexpr (R == 7) return make_double2(kCos7[k], kSin7[k]);
    else if constexpr (R == 16) return make_double2(kCos16[k], kSin16[k]);
    else return make_double2(kCos25[k], kSin25[k]);
}

__device__ __forceinline__ cpx operator+(cpx a, cpx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx operator-(cpx a, cpx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by DIR * i  (DIR = -1: the forward transform's -i)
template <int DIR>
__device__ __forceinline__ cpx mul_i(cpx a) { return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }

// In-register DFTs, natural order in and out, stride S between the elements of v:
//   v[q*S] <- sum_r v[r*S] * e^{DIR * 2*pi*i*q*r/R}
template <int R, int DIR, int S>
struct Dft;

template <int DIR, int S>
struct Dft<1, DIR, S> {
    static __device__ __forceinline__ void run(cpx*) {}
};
template <int DIR, int S>
struct Dft<2, DIR, S> {
    static __device__ __forceinline__ void run(cpx* v)
    {
        const cpx a = v[0], b = v[S];
        v[0] = a + b;
        v[S] = a - b;
    }
};
template <int DIR, int S>
struct Dft<4, DIR, S> {
    static __device__ __forceinline__ void run(cpx* v)
    {
        const cpx b0 = v[0] + v[2 * S], b1 = v[0] - v[2 * S];
        const cpx b2 = v[S] + v[3 * S], b3 = mul_i<DIR>(v[S] - v[3 * S]);
        v[0] = b0 + b2;
        v[S] = b1 + b3;
        v[2 * S] = b0 - b2;
        v[3 * S] = b1 - b3;
    }
};
template <int DIR, int S>
struct Dft<8, DIR, S> {
    static __device__ __forceinline__ void run(cpx* v)
    {
        constexpr double h = 0.7071067811865476;
        cpx e[4], o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            e[r] = v[r * S] + v[(r + 4) * S];
            o[r] = v[r * S] - v[(r + 4) * S];
        }
        // odd half times e^{DIR*2*pi*i*r/8}
        o[1] = DIR < 0 ? make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x))
                       : make_double2(h * (o[1].x - o[1].y), h * (o[1].y + o[1].x));
        o[2] = mul_i<DIR>(o[2]);
        o[3] = DIR < 0 ? make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y))
                       : make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y));
        Dft<4, DIR, 1>::run(e);
        Dft<4, DIR, 1>::run(o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[2 * q * S] = e[q];
            v[(2 * q + 1) * S] = o[q];
        }
    }
};
// odd primes: pairs (r, R-r) share cos and differ in the sign of sin
template <int R, int DIR, int S>
struct DftPrime {
    static __device__ __forceinline__ void run(cpx* v)
    {
        constexpr int H = (R - 1) / 2;
        cpx a[H], b[H];
        cpx sum = v[0];
#pragma unroll
        for (int r = 1; r <= H; ++r) {
            a[r - 1] = v[r * S] + v[(R - r) * S];
            b[r - 1] = v[r * S] - v[(R - r) * S];
            sum = sum + a[r - 1];
        }
        const cpx v0 = v[0];
        v[0] = sum;
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            cpx p = v0, t = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 1; r <= H; ++r) {
                const cpx w = root<R>((q * r) % R);
                p.x += w.x * a[r - 1].x;
                p.y += w.x * a[r - 1].y;
                t.x += w.y * b[r - 1].x;
                t.y += w.y * b[r - 1].y;
            }
            // forward: p - i*t for q, p + i*t for R-q; inverse the other way round
            const cpx it = mul_i<DIR>(t);
            v[q * S] = p + it;
            v[(R - q) * S] = p - it;
        }
    }
};
template <int DIR, int S> struct Dft<3, DIR, S> : DftPrime<3, DIR, S> {};
template <int DIR, int S> struct Dft<5, DIR, S> : DftPrime<5, DIR, S> {};
template <int DIR, int S> struct Dft<7, DIR, S> : DftPrime<7, DIR, S> {};

// R = A*A in registers: A DFTs of size A over r1 (r = r1*A + r2), the inner twiddles e^{DIR*2*pi*i*r2*q1/R},
// A DFTs of size A over r2; result q = q1 + A*q2 sits at y[q1*A + q2] and is put back in natural order
template <int A, int DIR>
struct DftSquare {
    static __device__ __forceinline__ void run(cpx* v)
    {
        constexpr int R = A * A;
#pragma unroll
        for (int r2 = 0; r2 < A; ++r2) Dft<A, DIR, A>::run(v + r2);        // over r1, stride A: v[q1*A + r2]
#pragma unroll
        for (int q1 = 1; q1 < A; ++q1)
#pragma unroll
            for (int r2 = 1; r2 < A; ++r2) {
                cpx w = root<R>((q1 * r2) % R);
                if (DIR < 0) w.y = -w.y;
                v[q1 * A + r2] = cmul(v[q1 * A + r2], w);
            }
#pragma unroll
        for (int q1 = 0; q1 < A; ++q1) Dft<A, DIR, 1>::run(v + q1 * A);    // over r2: v[q1*A + q2]
        // transpose to natural order: out[q1 + A*q2] = v[q1*A + q2]
#pragma unroll
        for (int i = 0; i < A; ++i)
#pragma unroll
            for (int j = i + 1; j < A; ++j) {
                const cpx t = v[i * A + j];
                v[i * A + j] = v[j * A + i];
                v[j * A + i] = t;
            }
    }
};
template <int DIR> struct Dft<16, DIR, 1> : DftSquare<4, DIR> {};
template <int DIR> struct Dft<25, DIR, 1> : DftSquare<5, DIR> {};

// REAL2: element i is the packed pair (x[2i], x[2i+1]) of a real signal (half-length transforms of even lengths)
enum { LOAD_CPX = 0, LOAD_REAL = 1, LOAD_BINS = 2, LOAD_REAL2 = 3 };
enum { STORE_CPX = 0, STORE_REAL = 1, STORE_REAL2 = 2 };

struct Io {
    const void* src;           // cpx (LOAD_CPX / LOAD_BINS) or float (LOAD_REAL)
    void* dst;                 // cpx (STORE_CPX) or float (STORE_REAL)
    long long src_stride;      // elements between the transforms of a batch
    long long dst_stride;
    int len;                   // transform length
    int n_src, pos, neg;       // LOAD_BINS: element i of the length-len spectrum comes from the length-n_src one
    double scale;              // STORE_REAL
};

template <int LOAD>
__device__ __forceinline__ cpx load_elem(const Io& io, int b, int i)
{
    if constexpr (LOAD == LOAD_REAL) {
        return make_double2((double)((const float*)io.src)[(long long)b * io.src_stride + i], 0.0);
    } else if constexpr (LOAD == LOAD_REAL2) {
        const float* p = (const float*)io.src + (long long)b * io.src_stride + 2LL * i;
        return make_double2((double)p[0], (double)p[1]);
    } else if constexpr (LOAD == LOAD_BINS) {                                // lib.rs:256-266
        const cpx* s = (const cpx*)io.src + (long long)b * io.src_stride;
        if (i < io.pos) return s[i];
        if (i >= io.len - io.neg) return s[io.n_src - (io.len - i)];
        return make_double2(0.0, 0.0);
    } else {
        return ((const cpx*)io.src)[(long long)b * io.src_stride + i];
    }
}
template <int STORE>
__device__ __forceinline__ void store_elem(const Io& io, int b, int i, cpx v)
{
    if constexpr (STORE == STORE_REAL) {
        ((float*)io.dst)[(long long)b * io.dst_stride + i] = (float)(v.x * io.scale);     // lib.rs:273-274
    } else if constexpr (STORE == STORE_REAL2) {
        float* p = (float*)io.dst + (long long)b * io.dst_stride + 2LL * i;
        p[0] = (float)(v.x * io.scale);
        p[1] = (float)(v.y * io.scale);
    } else
        ((cpx*)io.dst)[(long long)b * io.dst_stride + i] = v;
}

// One Stockham pass of radix R: butterfly j reads elements j + r*T (T = len/R), applies the twiddles
// e^{DIR*2*pi*i*r*k/(Ns*R)} (k = j mod Ns, Ns = product of the radices already done) and writes
// (j/Ns)*Ns*R + k + q*Ns.
template <int R, int DIR, int LOAD, int STORE>
__global__ void __launch_bounds__(R >= 16 ? 128 : 256)
k_pass(Io io, int T, int Ns)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= T) return;
    const int b = blockIdx.y;
    cpx v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = load_elem<LOAD>(io, b, j + r * T);
    const int k = j % Ns;
    if (R > 1 && k != 0) {
        double s, c;
        sincospi((DIR < 0 ? -2.0 : 2.0) * (double)k / ((double)Ns * (double)R), &s, &c);
        const cpx w1 = make_double2(c, s);
        cpx w = w1;
#pragma unroll
        for (int r = 1; r < R; ++r) {
            v[r] = cmul(v[r], w);
            if (r + 1 < R) w = cmul(w, w1);
        }
    }
    Dft<R, DIR, 1>::run(v);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) store_elem<STORE>(io, b, j0 + q * Ns, v[q]);
}

// ---- Bluestein: X[q] = w[q] * sum_k (x[k] w[k]) conj(w)[q-k],  w[k] = e^{DIR*i*pi*k^2/len}
template <int DIR>
__device__ __forceinline__ cpx chirp(long long k, int len)
{
    const long long r = (k * k) % (2LL * len);
    double s, c;
    sincospi((double)r / (double)len, &s, &c);
    return make_double2(c, DIR < 0 ? -s : s);
}
// a[k] = x[k] * w[k] for k < len, 0 up to P
template <int DIR, int LOAD>
__global__ void __launch_bounds__(256) k_bs_pre(Io io, cpx* a, long long a_stride, int P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int b = blockIdx.y;
    cpx v = make_double2(0.0, 0.0);
    if (k < io.len) v = cmul(load_elem<LOAD>(io, b, k), chirp<DIR>(k, io.len));
    a[(long long)b * a_stride + k] = v;
}
// b[d] = conj(w[|d|]) at d and P-d
template <int DIR>
__global__ void __launch_bounds__(256) k_bs_kernel(cpx* bq, int len, int P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int d = k < len ? k : (P - k < len ? P - k : -1);
    cpx v = make_double2(0.0, 0.0);
    if (d >= 0) {
        v = chirp<DIR>(d, len);
        v.y = -v.y;
    }
    bq[k] = v;
}
__global__ void __launch_bounds__(256) k_bs_mul(cpx* a, long long a_stride, const cpx* __restrict__ bf, int P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    cpx* p = a + (long long)blockIdx.y * a_stride + k;
    *p = cmul(*p, bf[k]);
}
// X[q] = c[q] * w[q] / P
template <int DIR, int STORE>
__global__ void __launch_bounds__(256) k_bs_post(const cpx* __restrict__ c, long long c_stride, Io io, int P)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= io.len) return;
    const int b = blockIdx.y;
    cpx v = cmul(c[(long long)b * c_stride + q], chirp<DIR>(q, io.len));
    const double inv = 1.0 / (double)P;
    v.x *= inv;
    v.y *= inv;
    store_elem<STORE>(io, b, q, v);
}

// ---- even lengths: real transforms through half-length complex ones
// Z = FFT_{n/2} of the packed input z[j] = x[2j] + i x[2j+1].  The real signal's spectrum is
//   X[q] = ((Z[q] + conj Z[h-q]) - i e^{-2 pi i q/n} (Z[q] - conj Z[h-q])) / 2,  h = n/2, Z[h] = Z[0],
// the resampled spectrum Y keeps X[q] for q < pos and is zero up to m/2 (lib.rs:256-266; Hermitian, so the
// negative bins need no storage), and the packed output y[2j] + i y[2j+1] is the inverse FFT_{m/2} of
//   Z'[k] = (Y[k] + conj Y[g-k]) + i e^{2 pi i k/m} (Y[k] - conj Y[g-k]),  g = m/2.
// This kernel goes from Z straight to Z'.
__device__ __forceinline__ cpx real_bin(const cpx* __restrict__ Z, int q, int h, int n, int pos)
{
    if (q >= pos) return make_double2(0.0, 0.0);
    const cpx a = Z[q], p = Z[q == 0 ? 0 : h - q];
    const double ax = a.x + p.x, ay = a.y - p.y, bx = a.x - p.x, by = a.y + p.y;
    double s, c;
    sincospi(2.0 * (double)q / (double)n, &s, &c);
    return make_double2(0.5 * (ax + (c * by - s * bx)), 0.5 * (ay - (c * bx + s * by)));
}
__global__ void __launch_bounds__(256)
k_bridge(const cpx* __restrict__ Z, long long z_stride, cpx* __restrict__ out, long long o_stride, int n, int m, int pos)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = m / 2;
    if (k >= g) return;
    const cpx* z = Z + (long long)blockIdx.y * z_stride;
    const cpx yk = real_bin(z, k, n / 2, n, pos), yq = real_bin(z, g - k, n / 2, n, pos);
    const double ax = yk.x + yq.x, ay = yk.y - yq.y, bx = yk.x - yq.x, by = yk.y + yq.y;
    double s, c;
    sincospi(2.0 * (double)k / (double)m, &s, &c);
    out[(long long)blockIdx.y * o_stride + k] = make_double2(ax - (c * by + s * bx), ay + (c * bx - s * by));
}

// ---- host side
bool factorize(long long n, std::vector<int>& radices)
{
    radices.clear();
    if (n == 1) {
        radices.push_back(1);
        return true;
    }
    static const int order[] = {25, 16, 8, 7, 5, 4, 3, 2};
    for (int r : order)
        while (n % r == 0) {
            radices.push_back(r);
            n /= r;
        }
    return n == 1;
}
bool smooth(long long n)
{
    for (int p : {2, 3, 5, 7})
        while (n % p == 0) n /= p;
    return n == 1;
}
long long bluestein_len(long long len)
{
    long long p = 2 * len - 1;
    while (!smooth(p)) ++p;
    return p;
}
// complex elements one transform of `len` needs in each ping-pong buffer
long long span(long long len) { return smooth(len) ? len : bluestein_len(len); }

template <int R, int DIR>
cudaError_t launch_pass(const Io& io, int T, int Ns, int batch, int load, int store, cudaStream_t st)
{
    const int threads = R >= 16 ? 128 : 256;
    const dim3 grid((unsigned)((T + threads - 1) / threads), (unsigned)batch);
#define APD_RS_CASE(L, S)                                              \
    if (load == L && store == S) {                                    \
        k_pass<R, DIR, L, S><<<grid, threads, 0, st>>>(io, T, Ns);    \
        return cudaGetLastError();                                    \
    }
    if constexpr (DIR < 0) {
        APD_RS_CASE(LOAD_CPX, STORE_CPX)
        APD_RS_CASE(LOAD_REAL, STORE_CPX)
        APD_RS_CASE(LOAD_REAL2, STORE_CPX)
    } else {
        APD_RS_CASE(LOAD_CPX, STORE_CPX)
        APD_RS_CASE(LOAD_BINS, STORE_CPX)
        APD_RS_CASE(LOAD_CPX, STORE_REAL)
        APD_RS_CASE(LOAD_BINS, STORE_REAL)
        APD_RS_CASE(LOAD_CPX, STORE_REAL2)
    }
#undef APD_RS_CASE
    return cudaErrorInvalidValue;
}
template <int DIR>
cudaError_t launch_radix(int R, const Io& io, int T, int Ns, int batch, int load, int store, cudaStream_t st)
{
    switch (R) {
        case 1: return launch_pass<1, DIR>(io, T, Ns, batch, load, store, st);
        case 2: return launch_pass<2, DIR>(io, T, Ns, batch, load, store, st);
        case 3: return launch_pass<3, DIR>(io, T, Ns, batch, load, store, st);
        case 4: return launch_pass<4, DIR>(io, T, Ns, batch, load, store, st);
        case 5: return launch_pass<5, DIR>(io, T, Ns, batch, load, store, st);
        case 7: return launch_pass<7, DIR>(io, T, Ns, batch, load, store, st);
        case 8: return launch_pass<8, DIR>(io, T, Ns, batch, load, store, st);
        case 16: return launch_pass<16, DIR>(io, T, Ns, batch, load, store, st);
        case 25: return launch_pass<25, DIR>(io, T, Ns, batch, load, store, st);
    }
    return cudaErrorInvalidValue;
}

// Smooth-length transform.  `first` describes where the first pass loads from (src, stride, load kind, bin
// map) and `last` where the last pass stores to; the passes in between ping-pong between p0 and p1 (stride
// `ws_stride`).  When `last` is NULL the result stays in the workspace and *result receives its buffer.
template <int DIR>
cudaError_t fft_smooth(int len, int batch, const Io& first, int load, const Io* last, int store, cpx* p0, cpx* p1,
                       long long ws_stride, cpx** result, cudaStream_t st)
{
    std::vector<int> radices;
    if (!factorize(len, radices)) return cudaErrorInvalidValue;
    int Ns = 1;
    const void* src = first.src;
    long long src_stride = first.src_stride;
    cpx* bufs[2] = {p0, p1};
    int which = (src == (const void*)p0) ? 1 : 0;          // never write the buffer being read
    for (size_t i = 0; i < radices.size(); ++i) {
        const int R = radices[i];
        const bool is_first = i == 0, is_last = i + 1 == radices.size();
        Io io = first;
        io.len = len;
        io.src = src;
        io.src_stride = src_stride;
        int st_kind = STORE_CPX;
        if (is_last && last) {
            io.dst = last->dst;
            io.dst_stride = last->dst_stride;
            io.scale = last->scale;
            st_kind = store;
        } else {
            io.dst = bufs[which];
            io.dst_stride = ws_stride;
        }
        cudaError_t e = launch_radix<DIR>(R, io, len / R, Ns, batch, is_first ? load : LOAD_CPX, st_kind, st);
        if (e != cudaSuccess) return e;
        if (!(is_last && last)) {
            src = bufs[which];
            src_stride = ws_stride;
            if (result) *result = bufs[which];
            which ^= 1;
        }
        Ns *= R;
    }
    return cudaSuccess;
}

// Any-length transform.  Smooth: see fft_smooth.  Otherwise Bluestein through p0/p1 (each batch*ws_stride) and
// the chirp buffers bq0/bq1 (P each); the spectrum then goes to `out_cpx` (stride ws_stride) or, if `last`,
// straight to the float output.
template <int DIR>
cudaError_t fft_any(int len, int batch, const Io& first, int load, const Io* last, int store, cpx* p0, cpx* p1,
                    cpx* out_cpx, cpx* bq0, cpx* bq1, long long ws_stride, cpx** result, cudaStream_t st)
{
    if (smooth(len)) return fft_smooth<DIR>(len, batch, first, load, last, store, p0, p1, ws_stride, result, st);
    const int P = (int)bluestein_len(len);
    const dim3 gridP((unsigned)((P + 255) / 256), (unsigned)batch), gridP1((unsigned)((P + 255) / 256), 1);
    cudaError_t e;
    // chirp filter spectrum
    k_bs_kernel<DIR><<<gridP1, 256, 0, st>>>(bq0, len, P);
    Io bio{};
    bio.src = bq0;
    bio.src_stride = P;
    cpx* bf = nullptr;
    if ((e = fft_smooth<-1>(P, 1, bio, LOAD_CPX, nullptr, STORE_CPX, bq0, bq1, P, &bf, st)) != cudaSuccess) return e;
    // a = x * w, zero padded
    Io pre = first;
    pre.len = len;
    if (load == LOAD_REAL) k_bs_pre<DIR, LOAD_REAL><<<gridP, 256, 0, st>>>(pre, p0, ws_stride, P);
    else if (load == LOAD_REAL2) k_bs_pre<DIR, LOAD_REAL2><<<gridP, 256, 0, st>>>(pre, p0, ws_stride, P);
    else if (load == LOAD_BINS) k_bs_pre<DIR, LOAD_BINS><<<gridP, 256, 0, st>>>(pre, p0, ws_stride, P);
    else k_bs_pre<DIR, LOAD_CPX><<<gridP, 256, 0, st>>>(pre, p0, ws_stride, P);
    Io aio{};
    aio.src = p0;
    aio.src_stride = ws_stride;
    cpx* af = nullptr;
    if ((e = fft_smooth<-1>(P, batch, aio, LOAD_CPX, nullptr, STORE_CPX, p0, p1, ws_stride, &af, st)) != cudaSuccess) return e;
    k_bs_mul<<<gridP, 256, 0, st>>>(af, ws_stride, bf, P);
    Io cio{};
    cio.src = af;
    cio.src_stride = ws_stride;
    cpx* cf = nullptr;
    if ((e = fft_smooth<1>(P, batch, cio, LOAD_CPX, nullptr, STORE_CPX, p0, p1, ws_stride, &cf, st)) != cudaSuccess) return e;
    Io post{};
    post.len = len;
    const dim3 gridL((unsigned)((len + 255) / 256), (unsigned)batch);
    if (last) {
        post.dst = last->dst;
        post.dst_stride = last->dst_stride;
        post.scale = last->scale;
        if (store == STORE_REAL2) k_bs_post<DIR, STORE_REAL2><<<gridL, 256, 0, st>>>(cf, ws_stride, post, P);
        else k_bs_post<DIR, STORE_REAL><<<gridL, 256, 0, st>>>(cf, ws_stride, post, P);
    } else {
        post.dst = out_cpx;
        post.dst_stride = ws_stride;
        k_bs_post<DIR, STORE_CPX><<<gridL, 256, 0, st>>>(cf, ws_stride, post, P);
        if (result) *result = out_cpx;
    }
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_copy_or_zero(const float* __restrict__ in, long long in_stride, float* out,
                                                      long long out_stride, long long n, int copy)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[blockIdx.y * out_stride + i] = copy ? in[blockIdx.y * in_stride + i] : 0.0f;
}

constexpr long long kMaxLen = 1LL << 28;

struct Plan {
    long long stride;      // complex elements per transform in each of the three workspace buffers
    long long chirp;       // complex elements in each of the two chirp buffers (0: both lengths smooth)
};
Plan make_plan(long long n, long long m)
{
    if (n % 2 == 0 && m % 2 == 0) {       // real transforms of even lengths run at half length
        n /= 2;
        m /= 2;
    }
    Plan p;
    p.stride = span(n) > span(m) ? span(n) : span(m);
    const long long cn = smooth(n) ? 0 : bluestein_len(n), cm = smooth(m) ? 0 : bluestein_len(m);
    p.chirp = cn > cm ? cn : cm;
    return p;
}

}  // namespace

extern "C" int apd_resample_workspace_bytes(int64_t n_in, int64_t n_out, int32_t batch, int64_t* bytes)
{
    if (!bytes || n_in < 0 || n_out < 0 || batch < 1 || n_in > kMaxLen || n_out > kMaxLen) return APD_ERR_INVALID;
    if (n_in == 0 || n_out == 0 || n_in == n_out) {
        *bytes = 0;
        return APD_OK;
    }
    const Plan p = make_plan(n_in, n_out);
    *bytes = (int64_t)sizeof(cpx) * (3 * p.stride * (long long)batch + 2 * p.chirp);
    return APD_OK;
}

extern "C" int apd_resample(const float* in_dev, int64_t n_in, int64_t in_stride, float* out_dev, int64_t n_out,
                            int64_t out_stride, int32_t batch, void* workspace_dev, int64_t workspace_bytes,
                            void* cuda_stream)
{
    if (n_in < 0 || n_out < 0 || batch < 1 || batch > 65535 || n_in > kMaxLen || n_out > kMaxLen) return APD_ERR_INVALID;
    if (n_out == 0) return APD_OK;
    if (!out_dev || (n_in > 0 && !in_dev) || in_stride < n_in || out_stride < n_out) return APD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (n_in == 0 || n_in == n_out) {                                     // lib.rs:237-242: zeros / identity
        const dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)batch);
        k_copy_or_zero<<<grid, 256, 0, st>>>(in_dev, in_stride, out_dev, out_stride, n_out, n_in != 0);
        return cudaGetLastError() == cudaSuccess ? APD_OK : APD_ERR_CUDA;
    }
    int64_t need = 0;
    apd_resample_workspace_bytes(n_in, n_out, batch, &need);
    if (!workspace_dev || workspace_bytes < need) return APD_ERR_OVERFLOW;
    const Plan p = make_plan(n_in, n_out);
    cpx* b0 = (cpx*)workspace_dev;
    cpx* b1 = b0 + p.stride * batch;
    cpx* b2 = b1 + p.stride * batch;
    cpx* q0 = b2 + p.stride * batch;
    cpx* q1 = q0 + p.chirp;

    const long long nc = n_in < n_out ? n_in : n_out;
    Io out{};
    out.dst = out_dev;
    out.dst_stride = out_stride;
    out.scale = 1.0 / (double)n_in;
    if (n_in % 2 == 0 && n_out % 2 == 0) {
        // packed real input -> Z (length n/2) -> Z' (length m/2) -> packed real output
        Io fin{};
        fin.src = in_dev;
        fin.src_stride = in_stride;
        cpx* Z = nullptr;
        if (fft_any<-1>((int)(n_in / 2), batch, fin, LOAD_REAL2, nullptr, STORE_CPX, b0, b1, b2, q0, q1, p.stride, &Z,
                        st) != cudaSuccess)
            return APD_ERR_CUDA;
        cpx* u = Z == b0 ? b1 : b0;
        cpx* v = Z == b2 ? b1 : b2;
        const int g = (int)(n_out / 2);
        k_bridge<<<dim3((unsigned)((g + 255) / 256), (unsigned)batch), 256, 0, st>>>(Z, p.stride, u, p.stride, (int)n_in,
                                                                                    (int)n_out, (int)((nc + 1) / 2));
        Io iin{};
        iin.src = u;
        iin.src_stride = p.stride;
        if (fft_any<1>(g, batch, iin, LOAD_CPX, &out, STORE_REAL2, u, v, nullptr, q0, q1, p.stride, nullptr, st) !=
            cudaSuccess)
            return APD_ERR_CUDA;
        return APD_OK;
    }

    // forward: float32 in -> spectrum X (length n_in) somewhere in the workspace
    Io fin{};
    fin.src = in_dev;
    fin.src_stride = in_stride;
    cpx* X = nullptr;
    if (fft_any<-1>((int)n_in, batch, fin, LOAD_REAL, nullptr, STORE_CPX, b0, b1, b2, q0, q1, p.stride, &X, st) !=
        cudaSuccess)
        return APD_ERR_CUDA;
    // inverse over the bin map, through the two buffers that do not hold X
    cpx* u = X == b0 ? b1 : b0;
    cpx* v = X == b2 ? b1 : b2;
    Io iin{};
    iin.src = X;
    iin.src_stride = p.stride;
    iin.len = (int)n_out;
    iin.n_src = (int)n_in;
    iin.pos = (int)((nc + 1) / 2);
    iin.neg = (int)((nc - 1) / 2);
    if (fft_any<1>((int)n_out, batch, iin, LOAD_BINS, &out, STORE_REAL, u, v, nullptr, q0, q1, p.stride, nullptr, st) !=
        cudaSuccess)
        return APD_ERR_CUDA;
    return APD_OK;
}
