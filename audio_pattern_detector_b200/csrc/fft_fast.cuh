// Register-resident three-pass sub-FFTs for the hot four-step shapes, in packed complex
// arithmetic (cpx2.cuh: FADD2 / FMUL2 / FFMA2).
//
// Each pass is a Stockham radix-R step in which a thread owns one whole butterfly in
// registers.  The FIRST pass reads its R inputs straight from global memory and the LAST
// pass hands its R outputs straight to the epilogue (global store / reduction), so shared
// memory is touched only by the two inter-pass exchanges.  The thread -> (transform q,
// butterfly j) mapping is chosen per kernel so that those direct global accesses are
// contiguous: row kernels put j on the lanes (a warp covers 256 contiguous bytes of a row),
// column kernels put q on the lanes (a warp covers 32..64-byte row segments of adjacent
// columns).
//
// Exchange buffer layouts (XOR swizzles found by exhaustive search over the three access
// patterns of the 8*8*R passes; every LDS.64 / STS.64 of a half-warp hits 16 distinct 8-byte slots):
//   row kernels    : buf[q * N + (e ^ ((e >> 3) & 15))]
//   column kernels : buf[(e ^ ((e >> 3) & 1)) * TB + q]    (TB = 8 adjacent columns on the lanes)
#pragma once
#include "cpx2.cuh"
#include "fft_sub.cuh"

namespace apd {

// ---------------------------------------------------------------- packed butterflies
template <int R, int SIGN> struct Dft2;

template <int SIGN> struct Dft2<4, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const c2 a = v[0] + v[2], b = v[0] - v[2];
        const c2 c = v[1] + v[3], d = rot<SIGN>(v[1] - v[3]);
        v[0] = a + c;
        v[1] = b + d;
        v[2] = a - c;
        v[3] = b - d;
    }
};

template <int SIGN> struct Dft2<8, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const float h = 0.70710678118654752440f;
        c2 e[4] = {v[0], v[2], v[4], v[6]};
        c2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft2<4, SIGN>::run(e);
        Dft2<4, SIGN>::run(o);
        // w8^1 o1 = h (o1 + SIGN i o1),  w8^2 o2 = SIGN i o2,  w8^3 o3 = h (SIGN i o3 - o3)
        const c2 t1 = o[1] + rot<SIGN>(o[1]);
        const c2 t2 = rot<SIGN>(o[2]);
        const c2 t3 = rot<SIGN>(o[3]) - o[3];
        v[0] = e[0] + o[0];
        v[4] = e[0] - o[0];
        v[1] = axpy(e[1], h, t1);
        v[5] = axpy(e[1], -h, t1);
        v[2] = e[2] + t2;
        v[6] = e[2] - t2;
        v[3] = axpy(e[3], h, t3);
        v[7] = axpy(e[3], -h, t3);
    }
};

template <int SIGN> struct Dft2<5, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const float c1 = 0.30901699437494742410f;          // cos(2pi/5)
        const float c2_ = -0.80901699437494742410f;        // cos(4pi/5)
        const float s1 = 0.95105651629515357212f;          // sin(2pi/5)
        const float s2 = 0.58778525229247312917f;          // sin(4pi/5)
        const c2 a1 = v[1] + v[4], b1 = v[1] - v[4];
        const c2 a2 = v[2] + v[3], b2 = v[2] - v[3];
        const c2 x0 = v[0];
        v[0] = x0 + a1 + a2;
        const c2 p1 = axpy(axpy(x0, c1, a1), c2_, a2);
        const c2 p2 = axpy(axpy(x0, c2_, a1), c1, a2);
        const c2 q1 = rot<SIGN>(axpy(mul2(b1, mk(s1, s1)), s2, b2));
        const c2 q2 = rot<SIGN>(axpy(mul2(b1, mk(s2, s2)), -s1, b2));
        v[1] = p1 + q1;
        v[4] = p1 - q1;
        v[2] = p2 + q2;
        v[3] = p2 - q2;
    }
};

template <int SIGN> struct Dft2<10, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        c2 e[5] = {v[0], v[2], v[4], v[6], v[8]};
        c2 o[5] = {v[1], v[3], v[5], v[7], v[9]};
        Dft2<5, SIGN>::run(e);
        Dft2<5, SIGN>::run(o);
        // w10^k = cos(pi k/5) + SIGN i sin(pi k/5)
        const float c1 = 0.80901699437494742410f, s1 = 0.58778525229247312917f;
        const float c2_ = 0.30901699437494742410f, s2 = 0.95105651629515357212f;
        const c2 t0 = o[0];
        const c2 t1 = cmul(o[1], c1, SIGN * s1), t2 = cmul(o[2], c2_, SIGN * s2);
        const c2 t3 = cmul(o[3], -c2_, SIGN * s2), t4 = cmul(o[4], -c1, SIGN * s1);
        v[0] = e[0] + t0; v[5] = e[0] - t0;
        v[1] = e[1] + t1; v[6] = e[1] - t1;
        v[2] = e[2] + t2; v[7] = e[2] - t2;
        v[3] = e[3] + t3; v[8] = e[3] - t3;
        v[4] = e[4] + t4; v[9] = e[4] - t4;
    }
};

template <int N_, int R0_, int R1_, int R2_> struct Shape3 {
    static constexpr int N = N_, R0 = R0_, R1 = R1_, R2 = R2_;
    static_assert(R0_ * R1_ * R2_ == N_, "radices must multiply to N");
};
using Shape512 = Shape3<512, 8, 8, 8>;
using Shape640 = Shape3<640, 8, 8, 10>;

// powers w^0..w^(R-1) of a unit complex number (tree: depth log2 R)
template <int R>
__device__ __forceinline__ void unit_powers(float2 w, float2* p)
{
    p[0] = make_float2(1.0f, 0.0f);
    if (R > 1) p[1] = w;
#pragma unroll
    for (int r = 2; r < R; ++r) p[r] = cmul(p[r >> 1], p[r - (r >> 1)]);
}

// One radix-R Stockham pass for butterfly j of a transform, in pieces so that (a) a single exchange
// buffer can be used (all loads of a pass complete before any store of that pass) and (b) the
// inter-pass twiddles w_{NS R}^{k r}, k = j mod NS, which depend only on the thread, can be computed
// once per CTA and kept in registers while the CTA loops over many transforms:
//   pass_twiddles : pw[r] = w^{k r}          (once)
//   bfly_ld       : v[r] = LD(j + r*T)
//   bfly_tw       : v[r] *= pw[r]
//   Dft2<R>::run  : R-point DFT in registers
//   bfly_store    : ST(j0 + r*NS, v[r])
template <int R, int SIGN, int NS>
__device__ __forceinline__ void pass_twiddles(int j, float2* pw)
{
    const int k = j % NS;
    unit_powers<R>(cispif((float)SIGN * 2.0f * (float)k * (1.0f / (float)(NS * R))), pw);
}

template <int R, int N, class LD>
__device__ __forceinline__ void bfly_ld(int j, LD ld, c2* v)
{
    constexpr int T = N / R;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ld(j + r * T);
}

template <int R>
__device__ __forceinline__ void bfly_tw(c2* v, const float2* pw)
{
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], pw[r]);
}

template <int R, int NS, class ST>
__device__ __forceinline__ void bfly_store(int j, ST st, const c2* v)
{
    const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
    for (int r = 0; r < R; ++r) st(j0 + r * NS, v[r]);
}

// t[r] = base * step^r, r < R (running product: R-1 complex multiplies)
template <int R>
__device__ __forceinline__ void geometric(float2 base, float2 step, float2* t)
{
    t[0] = base;
#pragma unroll
    for (int r = 1; r < R; ++r) t[r] = cmul(t[r - 1], step);
}

// Barrier over the NT threads (a multiple of 32, consecutive warps) that share barrier `id` (1..15):
// the two warps of a row exchange through shared memory without stalling the other rows of the CTA.
template <int NT>
__device__ __forceinline__ void group_sync(int id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory");
}

}  // namespace apd
