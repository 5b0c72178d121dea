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

template <int SIGN> struct Dft2<3, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const float s3 = 0.86602540378443864676f;           // sin(2pi/3)
        const c2 t = v[1] + v[2];
        const c2 u = axpy(v[0], -0.5f, t);
        const c2 d = rot<SIGN>(mul2(v[1] - v[2], mk(s3, s3)));
        v[0] = v[0] + t;
        v[1] = u + d;
        v[2] = u - d;
    }
};

// 6 = 2 x 3: DFT3 of the even and of the odd inputs, then w6^k o[k] = (cos(pi k / 3) + SIGN i sin(pi k / 3)) o[k]
template <int SIGN> struct Dft2<6, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const float s3 = 0.86602540378443864676f;           // sin(pi/3)
        c2 e[3] = {v[0], v[2], v[4]};
        c2 o[3] = {v[1], v[3], v[5]};
        Dft2<3, SIGN>::run(e);
        Dft2<3, SIGN>::run(o);
        const c2 t1 = cmul(o[1], 0.5f, SIGN * s3), t2 = cmul(o[2], -0.5f, SIGN * s3);
        v[0] = e[0] + o[0]; v[3] = e[0] - o[0];
        v[1] = e[1] + t1;   v[4] = e[1] - t1;
        v[2] = e[2] + t2;   v[5] = e[2] - t2;
    }
};

// 7: a_k = v[k] + v[7-k], b_k = v[k] - v[7-k]; X[m], X[7-m] = (v0 + sum_k cos(2 pi k m / 7) a_k) +- SIGN i sum_k sin(2 pi k m / 7) b_k
template <int SIGN> struct Dft2<7, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const float c1 = 0.62348980185873353053f, c2_ = -0.22252093395631440429f, c3 = -0.90096886790241912624f;
        const float s1 = 0.78183148246802980871f, s2 = 0.97492791218182360702f, s3 = 0.43388373911755812048f;
        const c2 a1 = v[1] + v[6], b1 = v[1] - v[6];
        const c2 a2 = v[2] + v[5], b2 = v[2] - v[5];
        const c2 a3 = v[3] + v[4], b3 = v[3] - v[4];
        const c2 x0 = v[0];
        v[0] = x0 + a1 + a2 + a3;
        const c2 p1 = axpy(axpy(axpy(x0, c1, a1), c2_, a2), c3, a3);
        const c2 p2 = axpy(axpy(axpy(x0, c2_, a1), c3, a2), c1, a3);
        const c2 p3 = axpy(axpy(axpy(x0, c3, a1), c1, a2), c2_, a3);
        const c2 q1 = rot<SIGN>(axpy(axpy(mul2(b1, mk(s1, s1)), s2, b2), s3, b3));
        const c2 q2 = rot<SIGN>(axpy(axpy(mul2(b1, mk(s2, s2)), -s3, b2), -s1, b3));
        const c2 q3 = rot<SIGN>(axpy(axpy(mul2(b1, mk(s3, s3)), -s1, b2), s2, b3));
        v[1] = p1 + q1; v[6] = p1 - q1;
        v[2] = p2 + q2; v[5] = p2 - q2;
        v[3] = p3 + q3; v[4] = p3 - q3;
    }
};

// 9 = 3 x 3: DFT3 over r1 of v[3 r1 + r2], twiddles w9^{q1 r2}, DFT3 over r2 -> X[q1 + 3 q2]
template <int SIGN> struct Dft2<9, SIGN> {
    static __device__ __forceinline__ void run(c2* v)
    {
        const float c1 = 0.76604444311897803520f, s1 = 0.64278760968653932632f;     // cos, sin (2pi/9)
        const float c2_ = 0.17364817766693034885f, s2 = 0.98480775301220805937f;    // (4pi/9)
        const float c4 = -0.93969262078590838405f, s4 = 0.34202014332566873304f;    // (8pi/9)
        c2 y[3][3];
#pragma unroll
        for (int r2 = 0; r2 < 3; ++r2) {
            c2 t[3] = {v[r2], v[3 + r2], v[6 + r2]};
            Dft2<3, SIGN>::run(t);
            y[0][r2] = t[0]; y[1][r2] = t[1]; y[2][r2] = t[2];
        }
        y[1][1] = cmul(y[1][1], c1, SIGN * s1);
        y[1][2] = cmul(y[1][2], c2_, SIGN * s2);
        y[2][1] = cmul(y[2][1], c2_, SIGN * s2);
        y[2][2] = cmul(y[2][2], c4, SIGN * s4);
#pragma unroll
        for (int q1 = 0; q1 < 3; ++q1) {
            Dft2<3, SIGN>::run(y[q1]);
            v[q1] = y[q1][0]; v[q1 + 3] = y[q1][1]; v[q1 + 6] = y[q1][2];
        }
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
using Shape384 = Shape3<384, 8, 8, 6>;
using Shape448 = Shape3<448, 8, 8, 7>;
using Shape512 = Shape3<512, 8, 8, 8>;
using Shape576 = Shape3<576, 8, 8, 9>;
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

// ---------------------------------------------------------------- shared-memory access by 32-bit address
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int OFF> __device__ __forceinline__ c2 lds(unsigned a)
{
    c2 r;
    asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(r.v) : "r"(a), "n"(OFF) : "memory");
    return r;
}
__device__ __forceinline__ void sts(unsigned a, c2 v)
{
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v.v) : "memory");
}
__device__ __forceinline__ c2 ldg_nc(const c2* p)          // read-only path, keeps the line in L1
{
    c2 r;
    r.v = __ldg(reinterpret_cast<const unsigned long long*>(p));
    return r;
}
__device__ __forceinline__ c2 ldg_stream(const c2* p)      // streamed once: do not pollute L1
{
    c2 r;
    asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(r.v) : "l"(p));
    return r;
}

// ---------------------------------------------------------------- TMA bulk copies (global -> shared) on an mbarrier
// 1-D cp.async.bulk: one thread arms the mbarrier with the byte count and issues the copy; consumers spin on the
// barrier's phase parity.  Addresses and sizes are multiples of 16 bytes.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()      // make the initialised barriers visible to the async proxy
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_row(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "APD_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra APD_MBAR_DONE;\n"
        "bra APD_MBAR_WAIT;\n"
        "APD_MBAR_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

// Column-kernel exchange layout: TB adjacent columns on the lanes (q = tid % TB), butterfly j = tid / TB,
// word(e, q) = TB * (e ^ ((e >> 3) & SW)) + q with SW = 3 (TB = 4) or 1 (TB = 8); N1 = 8 * 8 * R2.
//   loads  e = j + T r (T = N1/8 in pass 2, 64 in pass 3): 8 TB (j ^ s_r) + 8 q + 8 TB T r with
//          s_r = ((j >> 3) + (T / 8) r) & SW: one of SW + 1 per-thread bases, picked by r at compile time
//          (T = 64: s_r = s_0; T = 80: s_0 or s_0 + 2; T = 72: s_0 + r)         (constant + immediate)
//   stores e = 8 j + r (pass 1)                    : (8 TB (8 j + (j & SW)) + 8 q) ^ (8 TB r)
//   stores e = 64 (j >> 3) + (j & 7) + 8 r (pass 2): (8 TB (64 (j >> 3) + (j & 7)) + 8 q) ^ (8 TB (8 r + (r & SW)))
// The (constant ^ immediate) form needs the buffer base aligned to 8 TB 64 bytes; col_buffer_base()
// rounds a 1 KB aligned static buffer (declared with kColSlack extra elements) up to that.
template <int TB> struct ColLayout {
    static constexpr int SW = TB == 8 ? 1 : 3;
    static constexpr unsigned ALIGN = 8u * TB * 64u;
    static constexpr int SLACK = (int)(ALIGN / 8u);        // extra c2 elements to declare
};

template <int TB>
struct ColAddr {
    unsigned ld[ColLayout<TB>::SW + 1], st1, st2;          // ld[k]: swizzle term ((j >> 3) + k) & SW
    __device__ __forceinline__ ColAddr(const void* raw, int j, int q)
    {
        constexpr int SW = ColLayout<TB>::SW;
        constexpr unsigned AL = ColLayout<TB>::ALIGN;
        const unsigned sb = ((smem_addr(raw) + AL - 1u) & ~(AL - 1u)) + 8u * (unsigned)q;
#pragma unroll
        for (int k = 0; k <= SW; ++k) ld[k] = sb + 8u * TB * (unsigned)(j ^ (((j >> 3) + k) & SW));
        st1 = sb + 8u * TB * (unsigned)(8 * j + (j & SW));
        st2 = sb + 8u * TB * (unsigned)(64 * (j >> 3) + (j & 7));
    }
};

// v[r] = exchange[j + T r], r < R
// OFF: byte offset of the exchange buffer used (a multiple of ColLayout<TB>::ALIGN) when a kernel ping-pongs
// between two buffers to save the barrier that protects buffer reuse.
template <int TB, int T, int R, int OFF = 0> struct ColLoad {
    template <int r> static __device__ __forceinline__ c2 one(const ColAddr<TB>& A)
    {
        static_assert(T % 8 == 0, "butterfly count must be a multiple of 8");
        return lds<OFF + 8 * TB * T * r>(A.ld[((T / 8) * r) & ColLayout<TB>::SW]);
    }
    static __device__ __forceinline__ void run(const ColAddr<TB>& A, c2* v)
    {
        v[0] = one<0>(A); v[1] = one<1>(A); v[2] = one<2>(A); v[3] = one<3>(A);
        v[4] = one<4>(A); v[5] = one<5>(A);
        if (R >= 7) v[6] = one<6>(A);
        if (R >= 8) v[7] = one<7>(A);
        if (R >= 9) v[8] = one<8>(A);
        if (R >= 10) v[9] = one<9>(A);
    }
};
// 128-bit variants: a thread owns two adjacent columns (q even), whose words are adjacent in the exchange buffer
template <int OFF> __device__ __forceinline__ void lds128(unsigned a, c2& x, c2& y)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(x.v), "=l"(y.v) : "r"(a), "n"(OFF) : "memory");
}
__device__ __forceinline__ void sts128(unsigned a, c2 x, c2 y)
{
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(x.v), "l"(y.v) : "memory");
}
template <int TB, int T, int R, int OFF = 0> struct ColLoad2 {
    template <int r> static __device__ __forceinline__ void one(const ColAddr<TB>& A, c2* va, c2* vb)
    {
        static_assert(T % 8 == 0, "butterfly count must be a multiple of 8");
        lds128<OFF + 8 * TB * T * r>(A.ld[((T / 8) * r) & ColLayout<TB>::SW], va[r], vb[r]);
    }
    static __device__ __forceinline__ void run(const ColAddr<TB>& A, c2* va, c2* vb)
    {
        one<0>(A, va, vb); one<1>(A, va, vb); one<2>(A, va, vb); one<3>(A, va, vb);
        one<4>(A, va, vb); one<5>(A, va, vb);
        if (R >= 7) one<6>(A, va, vb);
        if (R >= 8) one<7>(A, va, vb);
        if (R >= 9) one<8>(A, va, vb);
        if (R >= 10) one<9>(A, va, vb);
    }
};
template <int TB, unsigned OFF = 0>
__device__ __forceinline__ void col_store1_x2(const ColAddr<TB>& A, const c2* va, const c2* vb)
{
#pragma unroll
    for (int r = 0; r < 8; ++r) sts128((A.st1 + OFF) ^ (8u * TB * r), va[r], vb[r]);
}
template <int TB, unsigned OFF = 0>
__device__ __forceinline__ void col_store2_x2(const ColAddr<TB>& A, const c2* va, const c2* vb)
{
#pragma unroll
    for (int r = 0; r < 8; ++r) sts128((A.st2 + OFF) ^ (8u * TB * (8 * r + (r & ColLayout<TB>::SW))), va[r], vb[r]);
}

template <int TB, unsigned OFF = 0> __device__ __forceinline__ void col_store1(const ColAddr<TB>& A, const c2* v)
{
    static_assert(OFF % ColLayout<TB>::ALIGN == 0, "buffer offset must keep the store address alignment");
#pragma unroll
    for (int r = 0; r < 8; ++r) sts((A.st1 + OFF) ^ (8u * TB * r), v[r]);
}
template <int TB, unsigned OFF = 0> __device__ __forceinline__ void col_store2(const ColAddr<TB>& A, const c2* v)
{
    static_assert(OFF % ColLayout<TB>::ALIGN == 0, "buffer offset must keep the store address alignment");
#pragma unroll
    for (int r = 0; r < 8; ++r) sts((A.st2 + OFF) ^ (8u * TB * (8 * r + (r & ColLayout<TB>::SW))), v[r]);
}

}  // namespace apd
