// Shared-memory mixed-radix Stockham sub-FFT used by the four-step passes.
//
// Data layout in shared memory: buf[e * LD + q]  (e = element along the transform,
// q = which of the TB = 2^tb_log2 independent transforms of this tile).  Threads are mapped
// with q fastest, so a warp touches consecutive 8-byte words for a fixed butterfly
// and shared-memory accesses are bank-conflict free; the per-pass twiddle is the
// same for all q of a butterfly (a broadcast read).
//
// SIGN = -1: forward (e^{-2 pi i jk/n}), SIGN = +1: inverse (unnormalised).
#pragma once
#include "common.cuh"

namespace apd {

constexpr int kMaxPasses = 8;

struct SubPlan {
    int n;                    // transform length
    int npass;
    int radix[kMaxPasses];    // product == n, each in {2,3,4,5,8}
    const float2* tw;         // tw[t] = e^{-2 pi i t / n}, t in [0, n)  (forward sign; conj for inverse)
};

template <int SIGN>
__device__ __forceinline__ float2 mul_i(float2 v)      // v * (SIGN * i)
{
    return SIGN > 0 ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
}

template <int R, int SIGN> struct Dft;

template <int SIGN> struct Dft<2, SIGN> {
    static __device__ __forceinline__ void run(float2* v)
    {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int SIGN> struct Dft<3, SIGN> {
    static __device__ __forceinline__ void run(float2* v)
    {
        const float s = 0.86602540378443864676f;          // sin(pi/3)
        const float2 t1 = cadd(v[1], v[2]);
        const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        const float2 d = csub(v[1], v[2]);
        const float2 t3 = mul_i<SIGN>(make_float2(s * d.x, s * d.y));
        v[0] = cadd(v[0], t1);
        v[1] = cadd(t2, t3);
        v[2] = csub(t2, t3);
    }
};

template <int SIGN> struct Dft<4, SIGN> {
    static __device__ __forceinline__ void run(float2* v)
    {
        const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        const float2 c = cadd(v[1], v[3]), d = mul_i<SIGN>(csub(v[1], v[3]));
        v[0] = cadd(a, c);
        v[1] = cadd(b, d);
        v[2] = csub(a, c);
        v[3] = csub(b, d);
    }
};

template <int SIGN> struct Dft<5, SIGN> {
    static __device__ __forceinline__ void run(float2* v)
    {
        const float c1 = 0.30901699437494742410f;          // cos(2pi/5)
        const float c2 = -0.80901699437494742410f;         // cos(4pi/5)
        const float s1 = 0.95105651629515357212f;          // sin(2pi/5)
        const float s2 = 0.58778525229247312917f;          // sin(4pi/5)
        const float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
        const float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
        const float2 x0 = v[0];
        v[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
        const float2 p1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
        const float2 p2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
        const float2 q1 = mul_i<SIGN>(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
        const float2 q2 = mul_i<SIGN>(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
        v[1] = cadd(p1, q1);
        v[4] = csub(p1, q1);
        v[2] = cadd(p2, q2);
        v[3] = csub(p2, q2);
    }
};

template <int SIGN> struct Dft<8, SIGN> {
    static __device__ __forceinline__ void run(float2* v)
    {
        const float h = 0.70710678118654752440f;
        // two radix-4 on even / odd inputs, then combine with w8^k
        float2 e[4] = {v[0], v[2], v[4], v[6]};
        float2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4, SIGN>::run(e);
        Dft<4, SIGN>::run(o);
        // w8^1 = (1 + SIGN i)/sqrt2, w8^2 = SIGN i, w8^3 = (-1 + SIGN i)/sqrt2
        const float2 o1 = SIGN > 0 ? make_float2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                                   : make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        const float2 o2 = mul_i<SIGN>(o[2]);
        const float2 o3 = SIGN > 0 ? make_float2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                                   : make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        v[0] = cadd(e[0], o[0]);
        v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);
        v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);
        v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);
        v[7] = csub(e[3], o3);
    }
};

// One Stockham pass of radix R over TB interleaved transforms; src -> dst.
template <int R, int SIGN>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                              const float2* __restrict__ tw, int n, int Ns, int tb_log2, int LD)
{
    const int T = n / R;
    const int tstep = n / (Ns * R);
    const int total = T << tb_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int q = idx & ((1 << tb_log2) - 1);
        const int j = idx >> tb_log2;
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[(j + r * T) * LD + q];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                float2 w = __ldg(&tw[k * r * tstep]);
                if (SIGN > 0) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        Dft<R, SIGN>::run(v);
        const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[(j0 + r * Ns) * LD + q] = v[r];
    }
}

// Full sub-FFT: ping-pongs between bufA and bufB; returns the buffer holding the result.
// All threads of the block must call it; includes the trailing __syncthreads().
template <int SIGN>
__device__ __forceinline__ float2* sub_fft(const SubPlan& sp, float2* bufA, float2* bufB, int tb_log2, int LD)
{
    float2* src = bufA;
    float2* dst = bufB;
    int Ns = 1;
    for (int p = 0; p < sp.npass; ++p) {
        const int R = sp.radix[p];
        switch (R) {
            case 8: stockham_pass<8, SIGN>(src, dst, sp.tw, sp.n, Ns, tb_log2, LD); break;
            case 5: stockham_pass<5, SIGN>(src, dst, sp.tw, sp.n, Ns, tb_log2, LD); break;
            case 4: stockham_pass<4, SIGN>(src, dst, sp.tw, sp.n, Ns, tb_log2, LD); break;
            case 3: stockham_pass<3, SIGN>(src, dst, sp.tw, sp.n, Ns, tb_log2, LD); break;
            default: stockham_pass<2, SIGN>(src, dst, sp.tw, sp.n, Ns, tb_log2, LD); break;
        }
        Ns *= R;
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    return src;
}

}  // namespace apd
