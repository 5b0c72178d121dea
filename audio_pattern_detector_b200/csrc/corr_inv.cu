// K2 -- fused spectral multiply + inverse four-step FFT + |.| + per-unit max (or normalised write-out)
// for the hot shapes M = N1 x 512, N1 in {512, 640}.
//
// Replaces abs(fft_correlation.fft_correlate_1d(section, clip, 'full')), max and the divide of
// reference audio_pattern_detector.py:491-494 for a whole launch of (chunk x pattern) units.
//
//   k_unit_desc : one 32-byte descriptor per unit of the launch (spectrum row pointers, N_out, where the
//                 unit's maximum lives, the divisor in write mode) so the hot kernels do no unit
//                 bookkeeping of their own.
//   k_corr_rows : rows c of  X[c][:] .* H[c][:]  ->  512-point inverse FFT  ->  four-step twiddle  -> W[c][:]
//   k_corr_cols : columns b of W  ->  N1-point inverse FFT  ->  e^{+i pi m/N}/M  ->  |Re|, |Im|  ->  max / write
//
// Both kernels keep one radix-8 (radix-10) butterfly per thread in registers, in packed complex
// arithmetic (cpx2.cuh: FADD2/FMUL2/FFMA2), read the first pass straight from global memory and hand
// the last pass straight to the epilogue.  All twiddles and all shared-memory addresses are loop
// invariants of the per-CTA loop over units: the XOR-swizzled exchange layouts (fft_fast.cuh) reduce to
// "thread constant + immediate" for loads and "thread constant ^ immediate" for stores.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <utility>

#include <cuda.h>

#include "fft_fast.cuh"
#include "internal.h"

namespace apd {

struct __align__(16) UnitDesc {
    const float2* xs;     // section spectrum of the unit's (chunk, group), row 0
    const float2* hs;     // spectrum of the unit's reversed clip, row 0
    int n_out;            // length of the 'full' correlation; < 0: slot not in use
    int max_idx;          // index into unit_max_bits
    float mc;             // write mode: max(self max, unit max)  (apd.py:493)
    int pad;
};
static_assert(sizeof(UnitDesc) == 32, "UnitDesc is loaded as two 16-byte words");

__global__ void k_unit_desc(UnitSrc U, UnitCtx C, const float2* __restrict__ spec, long long spec_stride,
                            InvOut O, int nunits, int write, UnitDesc* __restrict__ D)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nunits) return;
    UnitDesc d;
    int2 unit;
    if (!get_unit(U, u, &unit)) {
        d.xs = nullptr; d.hs = nullptr; d.n_out = -1; d.max_idx = 0; d.mc = 1.0f; d.pad = 0;
        D[u] = d;
        return;
    }
    long long start;
    int n;
    section_bounds(C.geoms[C.clip_group[unit.y]], unit.x, start, n);
    d.xs = spec + (long long)unit.x * spec_stride + C.clip_spec_off[unit.y];
    d.hs = C.clip_spec[unit.y];
    d.n_out = n > 0 ? n + C.clip_len[unit.y] - 1 : 0;
    d.max_idx = unit.x * O.n_clips + unit.y;
    d.mc = 1.0f;
    if (write) d.mc = fmaxf(O.self_max[unit.y], __uint_as_float(O.unit_max_bits[d.max_idx]));
    d.pad = 0;
    D[u] = d;
}

__device__ __forceinline__ c2 ldg_l2(const c2* p)          // L2 only: W may have been written by another SM of this launch
{
    c2 r;
    asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(r.v) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ UnitDesc load_desc(const UnitDesc* p)
{
    UnitDesc d;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    d.xs = reinterpret_cast<const float2*>(((unsigned long long)a.y << 32) | a.x);
    d.hs = reinterpret_cast<const float2*>(((unsigned long long)a.w << 32) | a.z);
    d.n_out = (int)b.x; d.max_idx = (int)b.y; d.mc = __uint_as_float(b.z); d.pad = 0;
    return d;
}

__device__ int g_walias = 0;       // timing experiment only (APD_B200_WALIAS): units share W slots, results invalid

constexpr int kN2 = 512;          // row length (complex)
constexpr int kRowsPerCta = 4;    // 64 threads per row

// ---------------------------------------------------------------- rows
// 64 threads (two warps) per row, kRowsPerCta rows per CTA; the two warps of a row synchronise among
// themselves only (named barrier q + 1).  Exchange layout of a row: slot(e) = e ^ ((e >> 3) & 15).
//   loads  e = j + 64 r            : 8 * (j ^ (j >> 3)) [^ 64 for odd r] + 512 r      (constant + immediate)
//   stores e = 8 j + r      (pass 1): 8 * ((8 j) ^ (j & 15))            ^ (8 r)       (constant ^ immediate)
//   stores e = 64 (j >> 3) + (j & 7) + 8 r (pass 2):
//                                     8 * (64 (j >> 3) + 8 ((j >> 3) & 1) + (j & 7)) ^ (72 r)
template <int R>
struct RowAddr {
    unsigned ld0, ld1, st1, st2;
    __device__ __forceinline__ RowAddr(unsigned row_base, int j)
    {
        ld0 = row_base + 8u * (unsigned)(j ^ (j >> 3));
        ld1 = ld0 ^ 64u;
        st1 = row_base + 8u * (unsigned)((8 * j) ^ (j & 15));
        st2 = row_base + 8u * (unsigned)(64 * (j >> 3) + 8 * ((j >> 3) & 1) + (j & 7));
    }
};

template <int R, unsigned OFF> __device__ __forceinline__ c2 row_ld(const RowAddr<8>& A)
{
    return (R & 1) ? lds<OFF + 512 * R>(A.ld1) : lds<OFF + 512 * R>(A.ld0);
}

#define APD_ROW_LOAD8(A, OFF, v)                                                                              \
    v[0] = row_ld<0, OFF>(A); v[1] = row_ld<1, OFF>(A); v[2] = row_ld<2, OFF>(A); v[3] = row_ld<3, OFF>(A);  \
    v[4] = row_ld<4, OFF>(A); v[5] = row_ld<5, OFF>(A); v[6] = row_ld<6, OFF>(A); v[7] = row_ld<7, OFF>(A);

// One work item of the row pass: rows [tile * ROWS, (tile + 1) * ROWS) of units [u_begin, u_end); unit u writes
// its rows to Wg + (u - u_begin) * M.  buf: ROWS * 2 * 512 complex of shared memory, 1 KB aligned.
// TMA = true (dense launches, every unit in use): the section-spectrum rows are staged by 1-D TMA bulk copies
// (cp.async.bulk on an mbarrier per row and stage) two units ahead, instead of a one-unit register prefetch.
// buf then holds ROWS * 4 * 512 complex: per row two exchange buffers and two staging buffers; bars: ROWS * 2.
template <int ROWS, bool KEEP_H, bool TMA = false>
__device__ __forceinline__ void rows_item(c2* buf, const UnitDesc* __restrict__ D, int u_begin, int u_end, int tile,
                                          int M, float2* __restrict__ Wg, unsigned long long* bars = nullptr)
{
    const int q = threadIdx.x >> 6, j = threadIdx.x & 63;
    const int c = tile * ROWS + q;
    constexpr int RS = TMA ? 4 : 2;                        // 512-element buffers per row
    // two exchange buffers per row (pass 1 -> 2 and pass 2 -> 3): one barrier per exchange, none for reuse
    const RowAddr<8> A(smem_addr(buf + q * (RS * kN2)), j);
    constexpr unsigned kB = kN2 * 8u;                      // byte offset of the second buffer
    const unsigned stage0 = smem_addr(buf + q * (RS * kN2) + 2 * kN2);     // staging buffers of this row (TMA)
    const unsigned bar0 = TMA ? smem_addr(bars + q * 2) : 0u;
    float2 tw2[8], tw3[8], fs[8];
    pass_twiddles<8, +1, 8>(j, tw2);
    pass_twiddles<8, +1, 64>(j, tw3);
    {   // outputs b = j + 64 r ; four-step twiddle w_M^{+b c} = base * step^r
        const float invM = 1.0f / (float)M;
        geometric<8>(twiddle_frac(j * c, invM, +1.0f), twiddle_frac(64 * c, invM, +1.0f), fs);
    }
    const long long row_off = (long long)c * kN2 + j;
    const float2* cur_hs = nullptr;
    c2 h[8], xn[8];
    // software pipeline: the operands of unit u + 1 are requested right after the first exchange of unit u,
    // so their latency is covered by two passes of arithmetic instead of by other warps
    UnitDesc dn = load_desc(D + u_begin);
    if (TMA) {
        if (j == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar0 + 8, 1);
            mbar_init_fence();
            const long long rrow = (long long)c * kN2;
            tma_load_row(stage0, dn.xs + rrow, kN2 * 8u, bar0);
            if (u_begin + 1 < u_end) tma_load_row(stage0 + kB, load_desc(D + u_begin + 1).xs + rrow, kN2 * 8u, bar0 + 8);
        }
        group_sync<64>(q + 1);                             // barriers initialised before anybody waits on them
    } else if (dn.n_out >= 0) {
        const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
    }
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = dn;
        const bool more = u + 1 < u_end;
        if (more) dn = load_desc(D + u + 1);
        if (TMA) {
            const int k = u - u_begin;
            const unsigned sb = stage0 + (k & 1) * kB + 8u * (unsigned)j;
            mbar_wait(bar0 + 8u * (k & 1), (unsigned)((k >> 1) & 1));
            xn[0] = lds<0>(sb); xn[1] = lds<512>(sb); xn[2] = lds<1024>(sb); xn[3] = lds<1536>(sb);
            xn[4] = lds<2048>(sb); xn[5] = lds<2560>(sb); xn[6] = lds<3072>(sb); xn[7] = lds<3584>(sb);
        }
        if (!TMA && d.n_out < 0) {
            if (more && dn.n_out >= 0) {
                const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
                for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
            }
            continue;
        }
        const c2* __restrict__ hs = reinterpret_cast<const c2*>(d.hs + row_off);
        c2 v[8];
        if (KEEP_H) {
            // consecutive units of a launch share the clip (clip-major order): its row stays in registers
            if (d.hs != cur_hs) {
#pragma unroll
                for (int r = 0; r < 8; ++r) h[r] = ldg_nc(hs + 64 * r);
                cur_hs = d.hs;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = cmul(xn[r], h[r]);
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = cmul(xn[r], ldg_nc(hs + 64 * r));
        }
        Dft2<8, +1>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts(A.st1 ^ (8u * r), v[r]);
        if (!TMA && more && dn.n_out >= 0) {
            const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
            for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
        }
        group_sync<64>(q + 1);
        if (TMA && j == 0 && u + 2 < u_end) {
            // every thread of the row has read this unit's staging buffer (it is behind the barrier): refill it
            const int k = u - u_begin;
            tma_load_row(stage0 + (k & 1) * kB, load_desc(D + u + 2).xs + (long long)c * kN2, kN2 * 8u, bar0 + 8u * (k & 1));
        }
        APD_ROW_LOAD8(A, 0, v)
        bfly_tw<8>(v, tw2);
        Dft2<8, +1>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts((A.st2 + kB) ^ (72u * r), v[r]);
        group_sync<64>(q + 1);
        APD_ROW_LOAD8(A, kB, v)
        bfly_tw<8>(v, tw3);
        Dft2<8, +1>::run(v);
        const int wal = g_walias;
        c2* __restrict__ out = reinterpret_cast<c2*>(Wg + (long long)((wal ? u % wal : u) - u_begin) * M + row_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) out[64 * r] = cmul(v[r], fs[r]);
    }
}

template <bool KEEP_H>
__global__ void __launch_bounds__(kRowsPerCta * 64, KEEP_H ? 2 : 3)
k_corr_rows(const UnitDesc* __restrict__ D, int nunits, int per, int M, float2* __restrict__ W, int swap)
{
    __shared__ __align__(1024) c2 buf[kRowsPerCta * 2 * kN2];
    const int bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    rows_item<kRowsPerCta, KEEP_H>(buf, D, by * per, min(nunits, (by + 1) * per), bx, M, W + (long long)(by * per) * M);
}

template <bool KEEP_H>
__global__ void __launch_bounds__(kRowsPerCta * 64, 3)
k_corr_rows3(const UnitDesc* __restrict__ D, int nunits, int per, int M, float2* __restrict__ W, int swap)
{
    __shared__ __align__(1024) c2 buf[kRowsPerCta * 2 * kN2];
    const int bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    rows_item<kRowsPerCta, KEEP_H>(buf, D, by * per, min(nunits, (by + 1) * per), bx, M, W + (long long)(by * per) * M);
}

// ---------------------------------------------------------------- rows, two passes (opt-in, APD_B200_ROWS2=1)
// 512 = 32 x 16 with ONE shared-memory exchange instead of two (tools/proto_two_pass.py):
// 16 threads per row, thread j owns the 32 elements e = j + 16 s at load and at store time (every global access of
// a half-warp is one contiguous 128-byte run); pass 1 is one radix-32 butterfly per thread (Ns = 1, no twiddles),
// pass 2 two radix-16 butterflies per thread (t = j and t = j + 16) whose inter-pass twiddles w_512^{t r} and
// four-step twiddles w_M^{(t + 32 q) c} are powers of hoisted bases, rebuilt per unit by a depth-4 product tree.
// The 16 threads of a row sit in one warp, so the exchange needs __syncwarp only: no block or named barriers.
// Exchange layout (conflict-free 64-bit accesses): element 32 jj + q of a row lives at 32 jj + (q ^ jj).
__device__ __forceinline__ c2 mul_root32(c2 x, int k)          // x * e^{+2 pi i k / 32}, k a compile-time constant
{
    constexpr float c[24] = {1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f,
                             0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f,
                             -0.831469612f, -0.923879533f, -0.98078528f, -1.0f, -0.98078528f, -0.923879533f,
                             -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f};
    constexpr float sn[24] = {0.0f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f,
                              0.98078528f, 1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f,
                              0.382683432f, 0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f,
                              -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f};
    if (k == 0) return x;
    if (k == 8) return rot_p(x);
    return cmul(x, c[k], sn[k]);
}
// inverse (sign +) DFTs of 32 = 4 x 8 and 16 = 4 x 4 points in registers, natural order in and out
__device__ __forceinline__ void idft32(c2* v)
{
    c2 y[4][8];
#pragma unroll
    for (int r2 = 0; r2 < 8; ++r2) {                                  // DFT4 over r1 of v[8 r1 + r2]
        c2 t[4] = {v[r2], v[8 + r2], v[16 + r2], v[24 + r2]};
        Dft2<4, +1>::run(t);
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) y[q1][r2] = mul_root32(t[q1], q1 * r2);
    }
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {                                  // DFT8 over r2 -> X[q1 + 4 q2]
        Dft2<8, +1>::run(y[q1]);
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) v[q1 + 4 * q2] = y[q1][q2];
    }
}
__device__ __forceinline__ void idft16(c2* v)
{
    c2 y[4][4];
#pragma unroll
    for (int r2 = 0; r2 < 4; ++r2) {
        c2 t[4] = {v[r2], v[4 + r2], v[8 + r2], v[12 + r2]};
        Dft2<4, +1>::run(t);
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) y[q1][r2] = mul_root32(t[q1], 2 * q1 * r2);
    }
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
        Dft2<4, +1>::run(y[q1]);
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) v[q1 + 4 * q2] = y[q1][q2];
    }
}
// p[r] = w^r, r < 16, by a product tree (depth 4) in packed arithmetic
__device__ __forceinline__ void powers16(c2 w, c2* p)
{
    p[0] = mk(1.0f, 0.0f);
    p[1] = w;
#pragma unroll
    for (int r = 2; r < 16; ++r) p[r] = cmul(p[r >> 1], p[r - (r >> 1)]);
}

constexpr int kRows2PerCta = 8;           // 16 threads per row -> 128 threads, 32 KB of shared memory

__global__ void __launch_bounds__(kRows2PerCta * 16, 4)
k_corr_rows2(const UnitDesc* __restrict__ D, int nunits, int per, int M, float2* __restrict__ W, int swap)
{
    __shared__ __align__(1024) c2 buf[kRows2PerCta * kN2];
    const int bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    const int u_begin = by * per, u_end = min(nunits, (by + 1) * per);
    const int q = threadIdx.x >> 4, j = threadIdx.x & 15;
    const int c = bx * kRows2PerCta + q;                                   // row of the N1 x 512 matrix
    const unsigned rb = smem_addr(buf + q * kN2);
    const unsigned st_base = rb + 8u * (unsigned)(32 * j);                 // + 8 * (q2 ^ j)
    const float invM = 1.0f / (float)M;
    // hoisted bases: inter-pass twiddle w_512^t and four-step twiddle w_M^{t c} for t = j, j + 16; step w_M^{32 c}
    c2 wt[2], f0[2];
#pragma unroll
    for (int u2 = 0; u2 < 2; ++u2) {
        const int t = j + 16 * u2;
        wt[u2] = from_f2(cispif(2.0f * (float)t * (1.0f / 512.0f)));
        f0[u2] = from_f2(twiddle_frac(t * c, invM, +1.0f));
    }
    const c2 fstep = from_f2(twiddle_frac(32 * c, invM, +1.0f));
    const long long row_off = (long long)c * kN2 + j;
    float2* __restrict__ Wg = W + (long long)u_begin * M;
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = load_desc(D + u);
        if (d.n_out < 0) continue;
        const c2* __restrict__ xs = reinterpret_cast<const c2*>(d.xs + row_off);
        const c2* __restrict__ hs = reinterpret_cast<const c2*>(d.hs + row_off);
        c2 v[32];
#pragma unroll
        for (int s_ = 0; s_ < 32; ++s_) v[s_] = ldg_stream(xs + 16 * s_);
#pragma unroll
        for (int s_ = 0; s_ < 32; ++s_) v[s_] = cmul(v[s_], ldg_nc(hs + 16 * s_));
        idft32(v);                                                         // butterfly j of T = 16: outputs 32 j + q2
#pragma unroll
        for (int q2 = 0; q2 < 32; ++q2) sts(st_base + 8u * (unsigned)(q2 ^ j), v[q2]);
        __syncwarp();
        c2 fq[16];
        powers16(fstep, fq);
        c2* __restrict__ out = reinterpret_cast<c2*>(Wg + (long long)(u - u_begin) * M + row_off);
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
            const int t = j + 16 * u2;
            c2 z[16], p[16];
            // element (jj = r, q2 = t) sits at 32 r + (t ^ r)
#pragma unroll
            for (int r = 0; r < 16; ++r) z[r] = lds<0>(rb + 8u * (unsigned)(32 * r + (t ^ r)));
            powers16(wt[u2], p);
#pragma unroll
            for (int r = 1; r < 16; ++r) z[r] = cmul(z[r], p[r]);
            idft16(z);                                                     // outputs e = t + 32 q
#pragma unroll
            for (int qq = 0; qq < 16; ++qq) out[16 * u2 + 32 * qq] = cmul(z[qq], cmul(f0[u2], fq[qq]));
        }
        __syncwarp();
    }
}

// Dense phase-1 launches: section-spectrum rows staged by TMA bulk copies (64 KB of dynamic shared memory per CTA).
__global__ void __launch_bounds__(kRowsPerCta * 64, 3)
k_corr_rows_tma(const UnitDesc* __restrict__ D, int nunits, int per, int M, float2* __restrict__ W, int swap)
{
    extern __shared__ __align__(1024) unsigned char rows_tma_smem[];
    c2* buf = reinterpret_cast<c2*>(rows_tma_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(buf + kRowsPerCta * 4 * kN2);
    const int bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    rows_item<kRowsPerCta, false, true>(buf, D, by * per, min(nunits, (by + 1) * per), bx, M,
                                        W + (long long)(by * per) * M, bars);
}

// ---------------------------------------------------------------- columns
// kTB adjacent columns on the lanes; exchange layout and address forms: ColAddr / ColLoad (fft_fast.cuh).
constexpr int kTB = 4;

// e^{i pi r / (2 R)}, r < R: the r-dependent part of the post-twiddle (immediates after unrolling)
__device__ __forceinline__ float2 post_const8(int r)
{
    constexpr float c[8] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f};
    constexpr float s[8] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f, 0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f};
    return make_float2(c[r], s[r]);
}
__device__ __forceinline__ float2 post_const10(int r)
{
    constexpr float c[10] = {1.0f, 0.98768834059513777f, 0.95105651629515353f, 0.8910065241883679f, 0.80901699437494745f, 0.70710678118654757f, 0.58778525229247314f, 0.4539904997395468f, 0.30901699437494745f, 0.15643446504023092f};
    constexpr float s[10] = {0.0f, 0.15643446504023087f, 0.3090169943749474f, 0.45399049973954675f, 0.58778525229247314f, 0.70710678118654746f, 0.80901699437494745f, 0.89100652418836779f, 0.95105651629515353f, 0.98768834059513777f};
    return make_float2(c[r], s[r]);
}

// One work item of the column pass: columns [tile * kTB, (tile + 1) * kTB) of units [u_begin, u_end); unit u reads
// Wg + (u - u_begin) * M and (WRITE) writes corr_g + (u - u_begin) * corr_stride.
// raw: 2 * N1 * kTB + ColLayout<kTB>::SLACK complex of shared memory, 1 KB aligned; red: two floats per warp.
template <class S, bool WRITE>
__device__ __forceinline__ void cols_item(c2* raw, float* red, const UnitDesc* __restrict__ D, int u_begin, int u_end,
                                          int tile, int M, const float2* __restrict__ Wg,
                                          unsigned int* __restrict__ unit_max_bits, float* __restrict__ corr_g,
                                          long long corr_stride)
{
    constexpr int N1 = S::N;
    constexpr int T1 = N1 / 8;            // butterflies per column in the radix-8 passes
    constexpr int R2 = S::R2;             // last radix: 8 or 10
    constexpr int NLAST = N1 / R2;        // 64 butterflies in the last pass
    constexpr int NW = kTB * T1 / 32;     // warps per CTA
    const int q = threadIdx.x % kTB, j = threadIdx.x / kTB;
    const int bcol = tile * kTB + q;
    const ColAddr<kTB> A(raw, j, q);
    float2 tw2[8], tw3[R2];
    pass_twiddles<8, +1, 8>(j, tw2);
    pass_twiddles<R2, +1, 64>(j, tw3);
    {
        // post-twiddle e^{+i pi m / N} / M with m = (j + 64 r) 512 + b = base * e^{i pi r / (2 R2)}: the base
        // is a per-thread scalar that commutes with the last butterfly, so it is folded into that pass's
        // input twiddles (r = 0 included); the r-dependent part is a compile-time constant (below).
        const float invN = 1.0f / (2.0f * (float)M);
        const float invM = 1.0f / (float)M;
        const float2 base = cispif((float)((j % NLAST) * kN2 + bcol) * invN);
        const float2 sbase = make_float2(base.x * invM, base.y * invM);
#pragma unroll
        for (int r = 0; r < R2; ++r) tw3[r] = cmul(tw3[r], sbase);
    }
    const long long col_off = (long long)j * kN2 + bcol;
    const int m0 = (j % NLAST) * kN2 + bcol;                 // output index of r = 0; r adds 64 * 512
    constexpr unsigned kBufB = N1 * kTB * 8u;                // second exchange buffer (pass 2 -> 3)
    int pending = -1, parity = 0;                            // unit whose per-warp maxima wait in red[parity ^ 1]
    c2 vn[8];
    UnitDesc dn = load_desc(D + u_begin);
    if (dn.n_out >= 0) {
        const c2* __restrict__ in = reinterpret_cast<const c2*>(Wg + col_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) vn[r] = ldg_l2(in + (long long)T1 * kN2 * r);
    }
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = dn;
        const bool more = u + 1 < u_end;
        if (more) dn = load_desc(D + u + 1);
        const c2* __restrict__ in_next = reinterpret_cast<const c2*>(Wg + (long long)(u + 1 - u_begin) * M + col_off);
        if (d.n_out < 0) {
            if (more && dn.n_out >= 0) {
#pragma unroll
                for (int r = 0; r < 8; ++r) vn[r] = ldg_l2(in_next + (long long)T1 * kN2 * r);
            }
            continue;
        }
        c2 v[R2 > 8 ? R2 : 8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = vn[r];
        Dft2<8, +1>::run(v);
        col_store1<kTB>(A, v);
        if (more && dn.n_out >= 0) {      // request the next unit's column while this one is transformed
#pragma unroll
            for (int r = 0; r < 8; ++r) vn[r] = ldg_l2(in_next + (long long)T1 * kN2 * r);
        }
        __syncthreads();
        if (!WRITE && pending >= 0 && threadIdx.x < 32) {     // block maximum of the previous unit (see below)
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
        ColLoad<kTB, T1, 8>::run(A, v);
        bfly_tw<8>(v, tw2);
        Dft2<8, +1>::run(v);
        col_store2<kTB, kBufB>(A, v);
        __syncthreads();
        float best = 0.0f;
        if (N1 == 512 || j < NLAST) {
            ColLoad<kTB, 64, R2, kBufB>::run(A, v);
#pragma unroll
            for (int r = 0; r < R2; ++r) v[r] = cmul(v[r], tw3[r]);
            Dft2<R2, +1>::run(v);
            float* __restrict__ out = WRITE ? corr_g + (long long)(u - u_begin) * corr_stride : nullptr;
            const int lim0 = d.n_out - m0, lim1 = d.n_out - M - m0;      // valid iff 32768 r < lim
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                // e^{i pi r / (2 R2)} as immediates
                const float2 pc = R2 == 8 ? post_const8(r < 8 ? r : 0) : post_const10(r);
                const float cr = pc.x, sr = pc.y;
                float zr, zi;
                split(r == 0 ? v[0] : cmul(v[r], cr, sr), zr, zi);
                const float y0 = fabsf(zr), y1 = fabsf(zi);
                if (WRITE) {
                    if (64 * kN2 * r < lim0) out[m0 + 64 * kN2 * r] = y0 / d.mc;          // apd.py:494 (float32 divide)
                    if (64 * kN2 * r < lim1) out[m0 + 64 * kN2 * r + M] = y1 / d.mc;
                } else {
                    if (64 * kN2 * r < lim0) best = fmaxf(best, y0);
                    if (64 * kN2 * r < lim1) best = fmaxf(best, y1);
                }
            }
        }
        if (!WRITE) {
            // per-warp maxima go to red[parity]; they are combined after the NEXT barrier the CTA passes anyway
            // (the first exchange of the next unit, or the one after the loop), so a unit costs two barriers
            best = warp_max(best);
            if ((threadIdx.x & 31) == 0) red[parity * NW + (threadIdx.x >> 5)] = best;
            pending = d.max_idx;
            parity ^= 1;
        }
    }
    if (!WRITE) {
        __syncthreads();
        if (pending >= 0 && threadIdx.x < 32) {
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
    }
}

template <class S, bool WRITE>
__global__ void __launch_bounds__(kTB * (S::N / 8), S::N == 512 ? 3 : 2)
k_corr_cols(const UnitDesc* __restrict__ D, int nunits, int per, int M, const float2* __restrict__ W,
            unsigned int* __restrict__ unit_max_bits, float* __restrict__ corr, long long corr_stride, int swap)
{
    __shared__ __align__(1024) c2 raw[2 * S::N * kTB + ColLayout<kTB>::SLACK];
    __shared__ float red[2 * (kTB * (S::N / 8) / 32)];
    const int bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    const int u0 = by * per;
    cols_item<S, WRITE>(raw, red, D, u0, min(nunits, u0 + per), bx, M, W + (long long)u0 * M, unit_max_bits,
                        WRITE ? corr + (long long)u0 * corr_stride : nullptr, corr_stride);
}

// ---------------------------------------------------------------- columns, two per thread (phase 1)
// Same pass structure as cols_item, but a thread owns the butterflies j of TWO adjacent columns: the exchange
// moves 16 bytes per instruction (the pair is adjacent in the swizzled layout, still conflict-free), the W
// column pair is one 16-byte global load, the twiddles are shared by the two butterflies, and a CTA executes
// twice the arithmetic between barriers.  The column-dependent part of the post-twiddle differs between the
// two columns by e^{i pi / N}; it is folded into per-output constants kept in constant memory.
__constant__ float2 c_post[2][2][10];        // [shape: 512, 640][column of the pair][r] = e^{i pi (r / (2 R2) + col / N)}

__device__ __forceinline__ void ldg_l2_x2(const c2* p, c2& x, c2& y)
{
    asm volatile("ld.global.cg.v2.b64 {%0, %1}, [%2];" : "=l"(x.v), "=l"(y.v) : "l"(p) : "memory");
}

template <class S>
__device__ __forceinline__ void cols2_item(c2* raw, float* red, const UnitDesc* __restrict__ D, int u_begin, int u_end,
                                           int tile, int M, const float2* __restrict__ Wg,
                                           unsigned int* __restrict__ unit_max_bits)
{
    constexpr int N1 = S::N;
    constexpr int T1 = N1 / 8;
    constexpr int R2 = S::R2;
    constexpr int NLAST = N1 / R2;
    constexpr int NW = 2 * T1 / 32;                          // warps per CTA (4 or 5)
    constexpr int SH = N1 == 512 ? 0 : 1;
    constexpr unsigned kBufB = N1 * kTB * 8u;
    const int p = threadIdx.x & 1, j = threadIdx.x >> 1;
    const int bcol = tile * kTB + 2 * p;                     // first column of the pair
    const ColAddr<kTB> A(raw, j, 2 * p);
    float2 tw2[8], tw3[R2];
    pass_twiddles<8, +1, 8>(j, tw2);
    pass_twiddles<R2, +1, 64>(j, tw3);
    {
        const float invN = 1.0f / (2.0f * (float)M);
        const float invM = 1.0f / (float)M;
        const float2 base = cispif((float)((j % NLAST) * kN2 + bcol) * invN);
        const float2 sbase = make_float2(base.x * invM, base.y * invM);
#pragma unroll
        for (int r = 0; r < R2; ++r) tw3[r] = cmul(tw3[r], sbase);
    }
    const long long col_off = (long long)j * kN2 + bcol;
    const int m0 = (j % NLAST) * kN2 + bcol;
    int pending = -1, parity = 0;
    c2 na[8], nb[8];
    UnitDesc dn = load_desc(D + u_begin);
    if (dn.n_out >= 0) {
        const c2* __restrict__ in = reinterpret_cast<const c2*>(Wg + col_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) ldg_l2_x2(in + (long long)T1 * kN2 * r, na[r], nb[r]);
    }
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = dn;
        const bool more = u + 1 < u_end;
        if (more) dn = load_desc(D + u + 1);
        const c2* __restrict__ in_next = reinterpret_cast<const c2*>(Wg + (long long)(u + 1 - u_begin) * M + col_off);
        if (d.n_out < 0) {
            if (more && dn.n_out >= 0) {
#pragma unroll
                for (int r = 0; r < 8; ++r) ldg_l2_x2(in_next + (long long)T1 * kN2 * r, na[r], nb[r]);
            }
            continue;
        }
        c2 va[R2 > 8 ? R2 : 8], vb[R2 > 8 ? R2 : 8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { va[r] = na[r]; vb[r] = nb[r]; }
        Dft2<8, +1>::run(va);
        Dft2<8, +1>::run(vb);
        col_store1_x2<kTB>(A, va, vb);
        if (more && dn.n_out >= 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) ldg_l2_x2(in_next + (long long)T1 * kN2 * r, na[r], nb[r]);
        }
        __syncthreads();
        if (pending >= 0 && threadIdx.x < 32) {
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
        ColLoad2<kTB, T1, 8>::run(A, va, vb);
        bfly_tw<8>(va, tw2);
        bfly_tw<8>(vb, tw2);
        Dft2<8, +1>::run(va);
        Dft2<8, +1>::run(vb);
        col_store2_x2<kTB, kBufB>(A, va, vb);
        __syncthreads();
        float best = 0.0f;
        if (N1 == 512 || j < NLAST) {
            ColLoad2<kTB, 64, R2, kBufB>::run(A, va, vb);
#pragma unroll
            for (int r = 0; r < R2; ++r) { va[r] = cmul(va[r], tw3[r]); vb[r] = cmul(vb[r], tw3[r]); }
            Dft2<R2, +1>::run(va);
            Dft2<R2, +1>::run(vb);
            const int lim0 = d.n_out - m0, lim1 = d.n_out - M - m0;      // column a valid iff 32768 r < lim; b: + 1
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                float ar, ai, br, bi;
                split(cmul(va[r], c_post[SH][0][r]), ar, ai);
                split(cmul(vb[r], c_post[SH][1][r]), br, bi);
                if (64 * kN2 * r < lim0) best = fmaxf(best, fabsf(ar));
                if (64 * kN2 * r < lim1) best = fmaxf(best, fabsf(ai));
                if (64 * kN2 * r + 1 < lim0) best = fmaxf(best, fabsf(br));
                if (64 * kN2 * r + 1 < lim1) best = fmaxf(best, fabsf(bi));
            }
        }
        best = warp_max(best);
        if ((threadIdx.x & 31) == 0) red[parity * NW + (threadIdx.x >> 5)] = best;
        pending = d.max_idx;
        parity ^= 1;
    }
    __syncthreads();
    if (pending >= 0 && threadIdx.x < 32) {
        float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
        t = warp_max(t);
        if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
    }
}

template <class S, int MINB>
__global__ void __launch_bounds__(2 * (S::N / 8), MINB)
k_corr_cols2(const UnitDesc* __restrict__ D, int nunits, int per, int M, const float2* __restrict__ W,
             unsigned int* __restrict__ unit_max_bits, int swap)
{
    __shared__ __align__(1024) c2 raw[2 * S::N * kTB + ColLayout<kTB>::SLACK];
    __shared__ float red[2 * (2 * (S::N / 8) / 32)];
    const int bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    const int u0 = by * per;
    cols2_item<S>(raw, red, D, u0, min(nunits, u0 + per), bx, M, W + (long long)u0 * M, unit_max_bits);
}

// ---------------------------------------------------------------- tiled intermediate (TMA in / TMA out)
// The four-step intermediate W of a unit is kept in 128-byte pieces: piece (ct, t) holds rows c = 4 ct .. 4 ct + 3 of
// columns b = 4 t .. 4 t + 3,
//     W[((c >> 2) * 128 + (b >> 2)) * 16 + (((c & 3) ^ ((b >> 2) & 3)) << 2) + (b & 3)],
// so that (a) the row pass, whose CTA owns four rows, stages its 16 KB of output in shared memory in exactly this
// order (the XOR makes the staging stores conflict-free) and hands it to ONE bulk shared->global copy per unit,
// and (b) the column pass, whose CTA owns four columns, fetches its N1 x 4 tile with ONE TMA tensor copy of
// N1/4 pieces.  Neither transfer passes through the LSU data pipe, which bounds both kernels: a column-pass
// LDG.128 touched 16 lines (32 bytes of each), a row-pass STG cost twice the wavefronts of an STS.
constexpr int kWTiles = kN2 / 4;                 // 128 column tiles
constexpr int kWBlock = kWTiles * 16;            // complex elements of one block of four rows (16 KB)

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_tile(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bytes, unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_tile3(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bytes, unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// the N1 x 4 tile `tile` of launch-local unit u: tiled layout (2-D map, N1/4 pieces of 128 bytes) or plain row-major
// layout (3-D map over [8][N1/8][512], N1 pieces of 32 bytes)
template <bool TILED, int N1>
__device__ __forceinline__ void tma_in(unsigned dst, const CUtensorMap* map, int tile, int u, unsigned bytes, unsigned bar)
{
    const int wal = g_walias;
    if (wal) u %= wal;
    if (TILED) tma_load_tile(dst, map, tile * 16, u * (N1 / 4), bytes, bar);
    else tma_load_tile3(dst, map, tile * 4, 0, u * 8, bytes, bar);
}

// best = max(best, y) if a < b (one ISETP + one predicated FMNMX; y >= 0)
__device__ __forceinline__ float max_if_lt(float best, float y, int a, int b)
{
    asm("{\n.reg .pred p;\nsetp.lt.s32 p, %2, %3;\n@p max.f32 %0, %0, %1;\n}" : "+f"(best) : "f"(y), "r"(a), "r"(b));
    return best;
}

// Row pass, tiled output.  Same passes as rows_item; the last pass writes to a double-buffered staging area.
template <bool KEEP_H>
__global__ void __launch_bounds__(kRowsPerCta * 64, 3)
k_corr_rows_t(const UnitDesc* __restrict__ D, int nunits, int per, int M, float2* __restrict__ W, int swap)
{
    extern __shared__ unsigned char rows_t_smem[];
    c2* buf = reinterpret_cast<c2*>((reinterpret_cast<uintptr_t>(rows_t_smem) + 1023) & ~uintptr_t(1023));
    const int tile = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    const int u_begin = by * per, u_end = min(nunits, (by + 1) * per);
    float2* __restrict__ Wg = W + (long long)u_begin * M + (long long)tile * kWBlock;
    const int q = threadIdx.x >> 6, j = threadIdx.x & 63;
    const int c = tile * kRowsPerCta + q;
    const RowAddr<8> A(smem_addr(buf + q * (2 * kN2)), j);
    constexpr unsigned kB = kN2 * 8u;
    // staging: two buffers of kWBlock elements behind the exchange buffers; output r of this thread is column
    // b = j + 64 r, i.e. piece t = (j >> 2) + 16 r, slot q ^ (t & 3), word j & 3
    const unsigned stage0 = smem_addr(buf + kRowsPerCta * 2 * kN2);
    const unsigned st_out = stage0 + 128u * (unsigned)(j >> 2) + 32u * (unsigned)(q ^ ((j >> 2) & 3)) + 8u * (unsigned)(j & 3);
    constexpr unsigned kStageB = kWBlock * 8u;
    float2 tw2[8], tw3[8], fs[8];
    pass_twiddles<8, +1, 8>(j, tw2);
    pass_twiddles<8, +1, 64>(j, tw3);
    {
        const float invM = 1.0f / (float)M;
        geometric<8>(twiddle_frac(j * c, invM, +1.0f), twiddle_frac(64 * c, invM, +1.0f), fs);
    }
    const long long row_off = (long long)c * kN2 + j;
    const float2* cur_hs = nullptr;
    c2 h[8], xn[8];
    UnitDesc dn = load_desc(D + u_begin);
    if (dn.n_out >= 0) {
        const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
    }
    int k = 0;                                             // units staged so far
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = dn;
        const bool more = u + 1 < u_end;
        if (more) dn = load_desc(D + u + 1);
        if (d.n_out < 0) {
            if (more && dn.n_out >= 0) {
                const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
                for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
            }
            continue;
        }
        const c2* __restrict__ hs = reinterpret_cast<const c2*>(d.hs + row_off);
        c2 v[8];
        if (KEEP_H) {
            if (d.hs != cur_hs) {
#pragma unroll
                for (int r = 0; r < 8; ++r) h[r] = ldg_nc(hs + 64 * r);
                cur_hs = d.hs;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = cmul(xn[r], h[r]);
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = cmul(xn[r], ldg_nc(hs + 64 * r));
        }
        Dft2<8, +1>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts(A.st1 ^ (8u * r), v[r]);
        if (more && dn.n_out >= 0) {
            const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
            for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
        }
        group_sync<64>(q + 1);
        APD_ROW_LOAD8(A, 0, v)
        bfly_tw<8>(v, tw2);
        Dft2<8, +1>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts((A.st2 + kB) ^ (72u * r), v[r]);
        group_sync<64>(q + 1);
        APD_ROW_LOAD8(A, kB, v)
        bfly_tw<8>(v, tw3);
        Dft2<8, +1>::run(v);
        const unsigned so = st_out + (k & 1) * kStageB;
#pragma unroll
        for (int r = 0; r < 8; ++r) sts(so + 2048u * r, cmul(v[r], fs[r]));
        fence_proxy_async_smem();                          // staged words visible to the bulk-copy engine
        // the copy that last read the OTHER staging buffer has finished reading before anybody passes this barrier
        if (threadIdx.x == 0) bulk_wait_read0();
        __syncthreads();
        if (threadIdx.x == 0) bulk_store(Wg + (long long)(u - u_begin) * M, stage0 + (k & 1) * kStageB, kStageB);
        ++k;
    }
    if (threadIdx.x == 0) bulk_wait_read0();
}

// Column pass, tiled input: the N1 x 4 tile of unit u arrives by one TMA tensor copy (box 16 x N1/4 of 8-byte
// elements) on an mbarrier; two columns per thread as in cols2_item.  NS input buffers: with one, the next unit's
// tile is requested after the first exchange (its latency is covered by two passes); with two, a whole unit ahead.
// WRITE: normalised correlation written out (phase 2) with the arithmetic of the maximum (phase 1), so that the
// phase-2 values divided by max(self max, phase-1 max) reproduce that maximum bit for bit.
template <class S, bool WRITE, int NS, bool TILED, int EPI = 1>
__global__ void __launch_bounds__(2 * (S::N / 8), WRITE ? 2 : (S::N == 512 ? 4 : 3))
k_corr_cols_t(const __grid_constant__ CUtensorMap tmap, const UnitDesc* __restrict__ D, int nunits, int per, int M,
              unsigned int* __restrict__ unit_max_bits, float* __restrict__ corr, long long corr_stride, int swap)
{
    constexpr int N1 = S::N;
    constexpr int T1 = N1 / 8;
    constexpr int R2 = S::R2;
    constexpr int NLAST = N1 / R2;
    constexpr int NW = 2 * T1 / 32;
    constexpr int SH = N1 == 512 ? 0 : 1;
    constexpr unsigned kBufB = N1 * kTB * 8u;
    constexpr unsigned kTileBytes = N1 * kTB * 8u;
    extern __shared__ unsigned char cols_t_smem[];
    __shared__ float red[2 * NW];
    __shared__ __align__(8) unsigned long long bars[NS];
    // [exchange buffers: 2 * N1 * kTB, aligned to ColLayout::ALIGN][input tiles: NS * N1 * kTB]
    constexpr uintptr_t AL = ColLayout<kTB>::ALIGN;
    c2* raw = reinterpret_cast<c2*>((reinterpret_cast<uintptr_t>(cols_t_smem) + AL - 1) & ~(AL - 1));
    const unsigned in0 = smem_addr(raw + 2 * N1 * kTB);
    const int tile = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    const int u_begin = by * per, u_end = min(nunits, (by + 1) * per);
    const int p = threadIdx.x & 1, j = threadIdx.x >> 1;
    const int bcol = tile * kTB + 2 * p;
    const ColAddr<kTB> A(raw, j, 2 * p);
    const unsigned ld_in = in0 + 32u * (unsigned)(TILED ? (j ^ (tile & 3)) : j) + 16u * (unsigned)p;
    const unsigned bar0 = smem_addr(bars);
    float2 tw2[8], tw3[R2];
    pass_twiddles<8, +1, 8>(j, tw2);
    pass_twiddles<R2, +1, 64>(j, tw3);
    {
        const float invN = 1.0f / (2.0f * (float)M);
        const float invM = 1.0f / (float)M;
        const float2 base = cispif((float)((j % NLAST) * kN2 + bcol) * invN);
        const float2 sbase = make_float2(base.x * invM, base.y * invM);
#pragma unroll
        for (int r = 0; r < R2; ++r) tw3[r] = cmul(tw3[r], sbase);
    }
    const int m0 = (j % NLAST) * kN2 + bcol;
    int pending = -1, parity = 0;
    int un = u_begin;                                      // next unit whose tile has not been requested (thread 0)
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8u * s, 1);
        mbar_init_fence();
        while (un < u_end && load_desc(D + un).n_out < 0) ++un;
        if (un < u_end) { tma_in<TILED, N1>(in0, &tmap, tile, un, kTileBytes, bar0); ++un; }
    }
    __syncthreads();
    int k = 0;                                             // valid units processed so far
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = load_desc(D + u);
        if (d.n_out < 0) continue;
        const int sb = k % NS;
        if (NS == 2 && threadIdx.x == 0) {
            // the other buffer was last read in the first pass of the previous unit: free since that unit's barriers
            while (un < u_end && load_desc(D + un).n_out < 0) ++un;
            if (un < u_end) { tma_in<TILED, N1>(in0 + (sb ^ 1) * kTileBytes, &tmap, tile, un, kTileBytes, bar0 + 8u * (sb ^ 1)); ++un; }
        }
        mbar_wait(bar0 + 8u * sb, (unsigned)((k / NS) & 1));
        c2 va[R2 > 8 ? R2 : 8], vb[R2 > 8 ? R2 : 8];
        {
            const unsigned a = ld_in + sb * kTileBytes;
            lds128<0 * 32 * T1>(a, va[0], vb[0]); lds128<1 * 32 * T1>(a, va[1], vb[1]);
            lds128<2 * 32 * T1>(a, va[2], vb[2]); lds128<3 * 32 * T1>(a, va[3], vb[3]);
            lds128<4 * 32 * T1>(a, va[4], vb[4]); lds128<5 * 32 * T1>(a, va[5], vb[5]);
            lds128<6 * 32 * T1>(a, va[6], vb[6]); lds128<7 * 32 * T1>(a, va[7], vb[7]);
        }
        Dft2<8, +1>::run(va);
        Dft2<8, +1>::run(vb);
        col_store1_x2<kTB>(A, va, vb);
        __syncthreads();
        if (NS == 1 && threadIdx.x == 0) {
            while (un < u_end && load_desc(D + un).n_out < 0) ++un;
            if (un < u_end) { tma_in<TILED, N1>(in0, &tmap, tile, un, kTileBytes, bar0); ++un; }
        }
        if (!WRITE && pending >= 0 && threadIdx.x < 32) {
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
        ColLoad2<kTB, T1, 8>::run(A, va, vb);
        bfly_tw<8>(va, tw2);
        bfly_tw<8>(vb, tw2);
        Dft2<8, +1>::run(va);
        Dft2<8, +1>::run(vb);
        col_store2_x2<kTB, kBufB>(A, va, vb);
        __syncthreads();
        float best = 0.0f;
        if (N1 == 512 || j < NLAST) {
            ColLoad2<kTB, 64, R2, kBufB>::run(A, va, vb);
#pragma unroll
            for (int r = 0; r < R2; ++r) { va[r] = cmul(va[r], tw3[r]); vb[r] = cmul(vb[r], tw3[r]); }
            Dft2<R2, +1>::run(va);
            Dft2<R2, +1>::run(vb);
            const int lim0 = d.n_out - m0, lim1 = d.n_out - M - m0;      // column a valid iff 32768 r < lim; b: + 1
            const bool all_re = d.n_out >= M;                            // unit-uniform (false only for a short last chunk)
            float* __restrict__ out = WRITE ? corr + (long long)u * corr_stride + m0 : nullptr;
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                float ar, ai, br, bi;
                split(cmul(va[r], c_post[SH][0][r]), ar, ai);
                split(cmul(vb[r], c_post[SH][1][r]), br, bi);
                ar = fabsf(ar); ai = fabsf(ai); br = fabsf(br); bi = fabsf(bi);
                if (WRITE) {
                    // apd.py:494 (float32 divide); m0 is even and corr_stride a multiple of 32: 8-byte stores
                    if (64 * kN2 * r + 1 < lim0) *reinterpret_cast<float2*>(out + 64 * kN2 * r) = make_float2(ar / d.mc, br / d.mc);
                    else if (64 * kN2 * r < lim0) out[64 * kN2 * r] = ar / d.mc;
                    if (64 * kN2 * r + 1 < lim1) *reinterpret_cast<float2*>(out + 64 * kN2 * r + M) = make_float2(ai / d.mc, bi / d.mc);
                    else if (64 * kN2 * r < lim1) out[64 * kN2 * r + M] = ai / d.mc;
                } else if (EPI == 0) {
                    if (64 * kN2 * r < lim0) best = fmaxf(best, ar);
                    if (64 * kN2 * r < lim1) best = fmaxf(best, ai);
                    if (64 * kN2 * r + 1 < lim0) best = fmaxf(best, br);
                    if (64 * kN2 * r + 1 < lim1) best = fmaxf(best, bi);
                } else if (all_re) {
                    // n_out >= M: every real part is a valid output; the imaginary parts (outputs m + M) need the test
                    best = fmaxf(best, fmaxf(ar, br));
                    best = max_if_lt(best, ai, 64 * kN2 * r, lim1);
                    best = max_if_lt(best, bi, 64 * kN2 * r + 1, lim1);
                } else {
                    best = max_if_lt(best, ar, 64 * kN2 * r, lim0);
                    best = max_if_lt(best, ai, 64 * kN2 * r, lim1);
                    best = max_if_lt(best, br, 64 * kN2 * r + 1, lim0);
                    best = max_if_lt(best, bi, 64 * kN2 * r + 1, lim1);
                }
            }
        }
        if (!WRITE) {
            best = warp_max(best);
            if ((threadIdx.x & 31) == 0) red[parity * NW + (threadIdx.x >> 5)] = best;
            pending = d.max_idx;
            parity ^= 1;
        }
        ++k;
    }
    if (!WRITE) {
        __syncthreads();
        if (pending >= 0 && threadIdx.x < 32) {
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
    }
}

// Column pass, one column per thread (4 x N1/8 threads, twice the warps of the two-column kernel for the same
// shared memory): the pass is bound by latency (barriers, shared-memory round trips), not by issue slots.  The last
// pass's twiddles are rebuilt per unit from two registers (w and the folded post-twiddle base) to stay under 68
// registers (three 320-thread CTAs per SM).
template <class S, int MINB>
__global__ void __launch_bounds__(kTB * (S::N / 8), MINB)
k_corr_cols_u(const __grid_constant__ CUtensorMap tmap, const UnitDesc* __restrict__ D, int nunits, int per, int M,
              unsigned int* __restrict__ unit_max_bits, int swap)
{
    constexpr int N1 = S::N;
    constexpr int T1 = N1 / 8;
    constexpr int R2 = S::R2;
    constexpr int NLAST = N1 / R2;
    constexpr int NW = kTB * T1 / 32;
    constexpr unsigned kBufB = N1 * kTB * 8u;
    constexpr unsigned kTileBytes = N1 * kTB * 8u;
    extern __shared__ unsigned char cols_u_smem[];
    __shared__ float red[2 * NW];
    __shared__ __align__(8) unsigned long long bars[1];
    constexpr uintptr_t AL = ColLayout<kTB>::ALIGN;
    c2* raw = reinterpret_cast<c2*>((reinterpret_cast<uintptr_t>(cols_u_smem) + AL - 1) & ~(AL - 1));
    const unsigned in0 = smem_addr(raw + 2 * N1 * kTB);
    const int tile = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
    const int u_begin = by * per, u_end = min(nunits, (by + 1) * per);
    const int q = threadIdx.x % kTB, j = threadIdx.x / kTB;
    const int bcol = tile * kTB + q;
    const ColAddr<kTB> A(raw, j, q);
    const unsigned ld_in = in0 + 32u * (unsigned)j + 8u * (unsigned)q;
    const unsigned bar0 = smem_addr(bars);
    float2 tw2[8];
    pass_twiddles<8, +1, 8>(j, tw2);
    // last pass: tw3[r] = w^r * sbase, w = e^{2 pi i (j mod 64) / (64 R2)}, sbase = e^{i pi m0 / N} / M
    const float2 w3 = cispif(2.0f * (float)(j % 64) * (1.0f / (float)(64 * R2)));
    float2 sbase;
    {
        const float invN = 1.0f / (2.0f * (float)M);
        const float invM = 1.0f / (float)M;
        const float2 base = cispif((float)((j % NLAST) * kN2 + bcol) * invN);
        sbase = make_float2(base.x * invM, base.y * invM);
    }
    const int m0 = (j % NLAST) * kN2 + bcol;
    int pending = -1, parity = 0;
    int un = u_begin;
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init_fence();
        while (un < u_end && load_desc(D + un).n_out < 0) ++un;
        if (un < u_end) { tma_in<false, N1>(in0, &tmap, tile, un, kTileBytes, bar0); ++un; }
    }
    __syncthreads();
    int k = 0;
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = load_desc(D + u);
        if (d.n_out < 0) continue;
        mbar_wait(bar0, (unsigned)(k & 1));
        c2 v[R2 > 8 ? R2 : 8];
        v[0] = lds<0 * 32 * T1>(ld_in); v[1] = lds<1 * 32 * T1>(ld_in); v[2] = lds<2 * 32 * T1>(ld_in);
        v[3] = lds<3 * 32 * T1>(ld_in); v[4] = lds<4 * 32 * T1>(ld_in); v[5] = lds<5 * 32 * T1>(ld_in);
        v[6] = lds<6 * 32 * T1>(ld_in); v[7] = lds<7 * 32 * T1>(ld_in);
        Dft2<8, +1>::run(v);
        col_store1<kTB>(A, v);
        __syncthreads();
        if (threadIdx.x == 0) {
            while (un < u_end && load_desc(D + un).n_out < 0) ++un;
            if (un < u_end) { tma_in<false, N1>(in0, &tmap, tile, un, kTileBytes, bar0); ++un; }
        }
        if (pending >= 0 && threadIdx.x < 32) {
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
        ColLoad<kTB, T1, 8>::run(A, v);
        bfly_tw<8>(v, tw2);
        Dft2<8, +1>::run(v);
        col_store2<kTB, kBufB>(A, v);
        __syncthreads();
        float best = 0.0f;
        if (N1 == 512 || j < NLAST) {
            ColLoad<kTB, 64, R2, kBufB>::run(A, v);
            {
                c2 t = from_f2(sbase);
                const c2 w = from_f2(w3);
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    v[r] = cmul(v[r], t);
                    if (r + 1 < R2) t = cmul(t, w);
                }
            }
            Dft2<R2, +1>::run(v);
            const int lim1 = d.n_out - M - m0;
            const int lim0 = d.n_out - m0;
            const bool all_re = d.n_out >= M;
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const float2 pc = R2 == 8 ? post_const8(r < 8 ? r : 0) : post_const10(r);
                float zr, zi;
                split(r == 0 ? v[0] : cmul(v[r], pc.x, pc.y), zr, zi);
                zr = fabsf(zr); zi = fabsf(zi);
                if (all_re) best = fmaxf(best, zr);
                else best = max_if_lt(best, zr, 64 * kN2 * r, lim0);
                best = max_if_lt(best, zi, 64 * kN2 * r, lim1);
            }
        }
        best = warp_max(best);
        if ((threadIdx.x & 31) == 0) red[parity * NW + (threadIdx.x >> 5)] = best;
        pending = d.max_idx;
        parity ^= 1;
        ++k;
    }
    __syncthreads();
    if (pending >= 0 && threadIdx.x < 32) {
        float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
        t = warp_max(t);
        if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
    }
}

// ---------------------------------------------------------------- fused persistent kernel
// Phase 1 (max only) as ONE persistent launch in which the four-step intermediate W never leaves L2.
// Units are taken in groups of U; a group's row pass (128 items: 4 or 5 rows each) and column pass (128 items: 4
// columns each) are work items of one queue, ordered
//     rows(0) .. rows(LAG-1), [rows(s), cols(s-LAG)] for s = LAG .. G-1, cols(G-LAG) .. cols(G-1)
// and handed out by an atomic counter.  A column item waits until all 128 row items of its group have
// published their W (release/acquire on a per-group counter); W lives in `slots` group-sized slots that are
// recycled once the group that used a slot has been read (second counter).  Every item waits only for items
// that were handed out before it, so the schedule cannot deadlock whatever the number of resident CTAs; with
// U * (LAG + 2) units in flight (~50 MB) W is written to and read back from the 126 MB L2, not HBM.
__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <class S, bool KEEP_H>
__global__ void __launch_bounds__(S::N / 2, 2)
k_corr_fused(const UnitDesc* __restrict__ D, int nunits, int U, int lag, int slots, int M, float2* __restrict__ W,
             unsigned int* __restrict__ unit_max_bits, int* __restrict__ counters)
{
    constexpr int N1 = S::N;
    constexpr int ROWS = N1 / 128;                       // rows per row item; ROWS * 64 == kTB * N1 / 8 threads
    constexpr int ITEMS = 128;                           // row items == column items per group
    constexpr int SMEM = (2 * ROWS * kN2 > 2 * N1 * kTB + ColLayout<kTB>::SLACK) ? 2 * ROWS * kN2 : 2 * N1 * kTB + ColLayout<kTB>::SLACK;
    static_assert(kN2 / kTB == ITEMS && N1 / ROWS == ITEMS, "work item counts");
    __shared__ __align__(1024) c2 smem[SMEM];
    __shared__ float red[2 * (N1 / 2 / 32)];
    __shared__ int s_item;
    const int G = (nunits + U - 1) / U;
    if (lag > G) lag = G;
    const int total = 2 * G * ITEMS;
    int* rows_done = counters + 1;
    int* cols_done = counters + 1 + G;
    for (;;) {
        __syncthreads();                                 // previous item done with smem and s_item
        if (threadIdx.x == 0) s_item = atomicAdd(counters, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        const int p = item / ITEMS, tile = item % ITEMS;
        bool is_rows;
        int g;
        if (p < lag) { is_rows = true; g = p; }
        else if (p < 2 * G - lag) { const int t = p - lag; is_rows = !(t & 1); g = is_rows ? lag + (t >> 1) : (t >> 1); }
        else { is_rows = false; g = p - G; }
        const int u0 = g * U, u1 = min(nunits, u0 + U);
        float2* Wg = W + (long long)((g % slots) * U) * M;
        if (is_rows) {
            if (g >= slots) {                            // the slot's previous tenant must have been read
                if (threadIdx.x == 0)
                    while (ld_acquire(cols_done + g - slots) < ITEMS) __nanosleep(100);
                __syncthreads();
            }
            rows_item<ROWS, KEEP_H>(smem, D, u0, u1, tile, M, Wg);
            __threadfence();                             // publish this thread's W stores
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(rows_done + g, 1);
        } else {
            if (threadIdx.x == 0)
                while (ld_acquire(rows_done + g) < ITEMS) __nanosleep(100);
            __syncthreads();
            cols_item<S, false>(smem, red, D, u0, u1, tile, M, Wg, unit_max_bits, nullptr, 0);
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(cols_done + g, 1);
        }
    }
}

// ---------------------------------------------------------------- launcher
static int env_int2(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool corr_inv_supported(const Fft4Plan& P) { return P.N2 == kN2 && (P.N1 == 512 || P.N1 == 640); }

// descriptors, then the fused kernel's counters: queue head + two per group (at most one group per unit)
static size_t desc_counters_offset(int nunits) { return sizeof(UnitDesc) * (size_t)(nunits > 0 ? nunits : 1); }
size_t corr_inv_desc_bytes(int nunits) { return desc_counters_offset(nunits) + sizeof(int) * (size_t)(2 * nunits + 8); }

static int fused_tc() { static int v = env_int2("APD_B200_FUSED_TC", 4); return v; }     // clips per unit tile
static int fused_tk() { static int v = env_int2("APD_B200_FUSED_TK", 2); return v; }     // chunks per unit tile
// off by default: measured slower than the two-kernel path ("fused persistent variant")
static int fused_on() { static int v = env_int2("APD_B200_FUSED", 0) && fused_tc() > 0 && fused_tk() > 0; return v; }

void corr_inv_tiling(int* tile_clips, int* tile_chunks)
{
    *tile_clips = fused_on() ? fused_tc() : 0;
    *tile_chunks = fused_on() ? fused_tk() : 0;
}

long long corr_inv_dense_units(int ns, int nb)
{
    if (!fused_on()) return (long long)ns * nb;
    const int tc = fused_tc(), tk = fused_tk();
    return (long long)((ns + tc - 1) / tc) * ((nb + tk - 1) / tk) * tc * tk;
}

// ---- tiled path: tensor maps over the intermediate, one per (buffer, shape) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static const CUtensorMap* w_tensor_map(float2* scratch, int N1, bool tiled)
{
    static std::map<std::pair<void*, int>, CUtensorMap> cache;
    static EncodeTiledFn encode = nullptr;
    const auto key = std::make_pair((void*)scratch, tiled ? N1 : -N1);
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return nullptr;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    // rows of kWBlock 8-byte elements (one block of four matrix rows); N1 / 4 of them per unit
    const cuuint64_t dims[2] = {(cuuint64_t)kWBlock, (cuuint64_t)(N1 / 4) * 65536ull};
    const cuuint64_t strides[1] = {(cuuint64_t)kWBlock * 8ull};
    const cuuint32_t box[2] = {16u, (cuuint32_t)(N1 / 4)};
    const cuuint32_t estr[2] = {1u, 1u};
    CUtensorMap m;
    if (!tiled) {
        // plain [c][b] layout seen as [8 * units][N1 / 8][512]: row c = j + (N1 / 8) r is element (., j, r)
        const cuuint64_t d3[3] = {512ull, (cuuint64_t)(N1 / 8), 8ull * 65536ull};
        const cuuint64_t s3[2] = {512ull * 8ull, (cuuint64_t)(N1 / 8) * 512ull * 8ull};
        const cuuint32_t b3[3] = {4u, (cuuint32_t)(N1 / 8), 8u};
        const cuuint32_t e3[3] = {1u, 1u, 1u};
        if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, scratch, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return nullptr;
        return &cache.emplace(key, m).first->second;
    }
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, scratch, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return nullptr;
    return &cache.emplace(key, m).first->second;
}

template <class S, bool WRITE, int NS, bool TILED, int EPI = 1>
static void launch_cols_t(const CUtensorMap& map, const UnitDesc* D, int nunits, int per, int M, const InvOut& out, dim3 grid,
                          int swap, cudaStream_t st)
{
    constexpr size_t smem = (size_t)(2 + NS) * S::N * kTB * sizeof(c2) + ColLayout<kTB>::ALIGN;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_corr_cols_t<S, WRITE, NS, TILED, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    k_corr_cols_t<S, WRITE, NS, TILED, EPI><<<grid, 2 * (S::N / 8), smem, st>>>(map, D, nunits, per, M, out.unit_max_bits, out.corr,
                                                                    out.corr_stride, swap);
}

static void upload_post_constants()
{
    static bool uploaded = false;
    if (uploaded) return;
    float2 h[2][2][10];
    const int r2[2] = {8, 10};
    const double n[2] = {2.0 * 512 * 512, 2.0 * 640 * 512};
    for (int sh = 0; sh < 2; ++sh)
        for (int col = 0; col < 2; ++col)
            for (int r = 0; r < 10; ++r) {
                const double a = M_PI * ((double)r / (2.0 * r2[sh]) + (double)col / n[sh]);
                h[sh][col][r] = make_float2((float)cos(a), (float)sin(a));
            }
    cudaMemcpyToSymbol(c_post, h, sizeof(h));
    uploaded = true;
}

static void launch_tiled(const Fft4Plan& P, const UnitDesc* D, int nunits, float2* scratch, const InvOut& out, bool write,
                         cudaStream_t st)
{
    static const int per_max = std::max(1, env_int2("APD_B200_PER", 16));
    static const int keep_h = env_int2("APD_B200_KEEP_H", 1);
    static const int swap = env_int2("APD_B200_SWAP", 1);
    static const int ns2 = env_int2("APD_B200_COLS_NS", 1) == 2;
    static const int epi0 = env_int2("APD_B200_EPI", 1) == 0;
    static const int tiled_w = env_int2("APD_B200_TILED", 1) == 2;     // 2: tiled intermediate; 1: plain layout, 32-byte pieces
    const CUtensorMap* map = w_tensor_map(scratch, P.N1, tiled_w);
    if (!map) { fprintf(stderr, "apd_b200: cuTensorMapEncodeTiled failed\n"); abort(); }
    upload_post_constants();
    static bool wal_set = false;
    if (!wal_set) {
        const int wal = env_int2("APD_B200_WALIAS", 0);
        if (wal > 0) cudaMemcpyToSymbol(g_walias, &wal, sizeof(int));
        wal_set = true;
    }
    int per = per_max;
    while (per > 1 && (long long)((nunits + per - 1) / per) * 64 < 148 * 8) per >>= 1;
    const int ny = (nunits + per - 1) / per;
    const int row_tiles = P.N1 / kRowsPerCta, col_tiles = kN2 / kTB;
    const dim3 gr = swap ? dim3(ny, row_tiles) : dim3(row_tiles, ny);
    const dim3 gc = swap ? dim3(ny, col_tiles) : dim3(col_tiles, ny);
    constexpr size_t kRowSmem = (size_t)(kRowsPerCta * 2 * kN2 + 2 * kWBlock) * sizeof(c2) + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_corr_rows_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmem);
        cudaFuncSetAttribute(k_corr_rows_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmem);
        attr = true;
    }
    if (tiled_w) {
        if (keep_h) k_corr_rows_t<true><<<gr, kRowsPerCta * 64, kRowSmem, st>>>(D, nunits, per, P.M, scratch, swap);
        else k_corr_rows_t<false><<<gr, kRowsPerCta * 64, kRowSmem, st>>>(D, nunits, per, P.M, scratch, swap);
    } else {
        if (keep_h) k_corr_rows3<true><<<gr, kRowsPerCta * 64, 0, st>>>(D, nunits, per, P.M, scratch, swap);
        else k_corr_rows3<false><<<gr, kRowsPerCta * 64, 0, st>>>(D, nunits, per, P.M, scratch, swap);
    }
    static const int cols_u = env_int2("APD_B200_COLS_U", 0);
    if (cols_u && !write && !tiled_w) {
        constexpr size_t sm512 = (size_t)3 * 512 * kTB * sizeof(c2) + ColLayout<kTB>::ALIGN;
        constexpr size_t sm640 = (size_t)3 * 640 * kTB * sizeof(c2) + ColLayout<kTB>::ALIGN;
        static bool attr_u = false;
        if (!attr_u) {
            cudaFuncSetAttribute(k_corr_cols_u<Shape512, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm512);
            cudaFuncSetAttribute(k_corr_cols_u<Shape640, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm640);
            cudaFuncSetAttribute(k_corr_cols_u<Shape512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm512);
            cudaFuncSetAttribute(k_corr_cols_u<Shape640, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm640);
            attr_u = true;
        }
        if (P.N1 == 512) {
            if (cols_u == 3) k_corr_cols_u<Shape512, 3><<<gc, kTB * 64, sm512, st>>>(*map, D, nunits, per, P.M, out.unit_max_bits, swap);
            else k_corr_cols_u<Shape512, 2><<<gc, kTB * 64, sm512, st>>>(*map, D, nunits, per, P.M, out.unit_max_bits, swap);
        } else {
            if (cols_u == 3) k_corr_cols_u<Shape640, 3><<<gc, kTB * 80, sm640, st>>>(*map, D, nunits, per, P.M, out.unit_max_bits, swap);
            else k_corr_cols_u<Shape640, 2><<<gc, kTB * 80, sm640, st>>>(*map, D, nunits, per, P.M, out.unit_max_bits, swap);
        }
        return;
    }
#define APD_COLS_T(SHAPE, TL)                                                                                        \
    do {                                                                                                             \
        if (write) launch_cols_t<SHAPE, true, 1, TL>(*map, D, nunits, per, P.M, out, gc, swap, st);                  \
        else if (ns2) launch_cols_t<SHAPE, false, 2, TL>(*map, D, nunits, per, P.M, out, gc, swap, st);              \
        else if (epi0) launch_cols_t<SHAPE, false, 1, TL, 0>(*map, D, nunits, per, P.M, out, gc, swap, st);          \
        else launch_cols_t<SHAPE, false, 1, TL>(*map, D, nunits, per, P.M, out, gc, swap, st);                       \
    } while (0)
    if (P.N1 == 512) { if (tiled_w) APD_COLS_T(Shape512, true); else APD_COLS_T(Shape512, false); }
    else { if (tiled_w) APD_COLS_T(Shape640, true); else APD_COLS_T(Shape640, false); }
#undef APD_COLS_T
}

void launch_corr_inv(const Fft4Plan& P, const UnitCtx& C, const float2* spec, long long spec_slab, const UnitSrc& U,
                     int nunits, float2* scratch, void* desc, const InvOut& out, bool write, cudaStream_t st)
{
    static const int per_max = std::max(1, env_int2("APD_B200_PER", 16));
    // 0: both operands re-read per unit (80 registers, 3 CTAs/SM); 1: the clip row stays in registers (128 registers,
    // 2 CTAs/SM).  Equal alone; the smaller CTAs lose less when phase-2 CTAs share the SMs.
    static const int keep_h = env_int2("APD_B200_KEEP_H", 0);
    static const int swap = env_int2("APD_B200_SWAP", 1);
    static const int lag = std::max(1, env_int2("APD_B200_FUSED_LAG", 1));
    static const int slots = std::max(lag + 1, env_int2("APD_B200_FUSED_SLOTS", 4));
    UnitDesc* D = static_cast<UnitDesc*>(desc);
    k_unit_desc<<<(nunits + 127) / 128, 128, 0, st>>>(U, C, spec, spec_slab, out, nunits, write ? 1 : 0, D);
    if (!write && U.tc > 0) {
        // fused persistent launch: one group = one unit tile
        const int gu = U.tc * U.tk;
        int* counters = reinterpret_cast<int*>(static_cast<char*>(desc) + desc_counters_offset(nunits));
        cudaMemsetAsync(counters, 0, sizeof(int) * (size_t)(2 * ((nunits + gu - 1) / gu) + 1), st);
        static int grid512 = 0, grid640 = 0;
        if (!grid512) {
            int dev = 0, sms = 148, b = 2;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_corr_fused<Shape512, true>, 256, 0);
            grid512 = sms * std::max(b, 1);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_corr_fused<Shape640, true>, 320, 0);
            grid640 = sms * std::max(b, 1);
        }
        if (P.N1 == 512) {
            if (keep_h) k_corr_fused<Shape512, true><<<grid512, 256, 0, st>>>(D, nunits, gu, lag, slots, P.M, scratch, out.unit_max_bits, counters);
            else k_corr_fused<Shape512, false><<<grid512, 256, 0, st>>>(D, nunits, gu, lag, slots, P.M, scratch, out.unit_max_bits, counters);
        } else {
            if (keep_h) k_corr_fused<Shape640, true><<<grid640, 320, 0, st>>>(D, nunits, gu, lag, slots, P.M, scratch, out.unit_max_bits, counters);
            else k_corr_fused<Shape640, false><<<grid640, 320, 0, st>>>(D, nunits, gu, lag, slots, P.M, scratch, out.unit_max_bits, counters);
        }
        return;
    }
    static const int tiled = env_int2("APD_B200_TILED", 1);
    if (tiled) {
        launch_tiled(P, D, nunits, scratch, out, write, st);
        return;
    }
    // keep at least ~8 CTAs per SM in the grid; otherwise amortise the twiddles over up to `per` units
    int per = per_max;
    while (per > 1 && (long long)((nunits + per - 1) / per) * 64 < 148 * 8) per >>= 1;
    const int ny = (nunits + per - 1) / per;
    const int row_tiles = P.N1 / kRowsPerCta, col_tiles = kN2 / kTB;
    const dim3 gr = swap ? dim3(ny, row_tiles) : dim3(row_tiles, ny);
    const dim3 gc = swap ? dim3(ny, col_tiles) : dim3(col_tiles, ny);
    // opt-in: measured 3 % slower than the register prefetch (profiles/sweeps_r1.txt) -- the row pass already runs at
    // ~70 % of the HBM copy bandwidth and the staged rows cost an extra shared-memory read per element
    static const int rows_tma = env_int2("APD_B200_ROWS_TMA", 0);
    // opt-in two-pass row kernel (one exchange, 16 threads per row); phase 1 only
    static const int rows2 = env_int2("APD_B200_ROWS2", 0);
    if (rows_tma && !write && U.list == nullptr) {
        // dense launch: every unit position is in use, so the staging pipeline needs no holes
        constexpr size_t kSmem = (size_t)kRowsPerCta * 4 * kN2 * sizeof(c2) + kRowsPerCta * 2 * sizeof(unsigned long long);
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(k_corr_rows_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
            attr = true;
        }
        k_corr_rows_tma<<<gr, kRowsPerCta * 64, kSmem, st>>>(D, nunits, per, P.M, scratch, swap);
    } else if (rows2 && !write) {
        const int tiles2 = P.N1 / kRows2PerCta;
        const dim3 gr2 = swap ? dim3(ny, tiles2) : dim3(tiles2, ny);
        k_corr_rows2<<<gr2, kRows2PerCta * 16, 0, st>>>(D, nunits, per, P.M, scratch, swap);
    } else if (keep_h) k_corr_rows<true><<<gr, kRowsPerCta * 64, 0, st>>>(D, nunits, per, P.M, scratch, swap);
    else k_corr_rows<false><<<gr, kRowsPerCta * 64, 0, st>>>(D, nunits, per, P.M, scratch, swap);
    static const int cols2 = env_int2("APD_B200_COLS2", 1);
    if (cols2 && !write) {
        static bool uploaded = false;
        if (!uploaded) {
            float2 h[2][2][10];
            const int r2[2] = {8, 10};
            const double n[2] = {2.0 * 512 * 512, 2.0 * 640 * 512};
            for (int sh = 0; sh < 2; ++sh)
                for (int col = 0; col < 2; ++col)
                    for (int r = 0; r < 10; ++r) {
                        const double a = M_PI * ((double)r / (2.0 * r2[sh]) + (double)col / n[sh]);
                        h[sh][col][r] = make_float2((float)cos(a), (float)sin(a));
                    }
            cudaMemcpyToSymbol(c_post, h, sizeof(h));
            uploaded = true;
        }
        static const int dense = env_int2("APD_B200_COLS2_DENSE", 0);      // one more CTA per SM (fewer registers)
        if (P.N1 == 512) {
            if (dense) k_corr_cols2<Shape512, 5><<<gc, 128, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, swap);
            else k_corr_cols2<Shape512, 4><<<gc, 128, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, swap);
        } else {
            if (dense) k_corr_cols2<Shape640, 4><<<gc, 160, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, swap);
            else k_corr_cols2<Shape640, 3><<<gc, 160, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, swap);
        }
        return;
    }
    if (P.N1 == 512) {
        if (write) k_corr_cols<Shape512, true><<<gc, kTB * 64, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, out.corr, out.corr_stride, swap);
        else k_corr_cols<Shape512, false><<<gc, kTB * 64, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, out.corr, out.corr_stride, swap);
    } else {
        if (write) k_corr_cols<Shape640, true><<<gc, kTB * 80, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, out.corr, out.corr_stride, swap);
        else k_corr_cols<Shape640, false><<<gc, kTB * 80, 0, st>>>(D, nunits, per, P.M, scratch, out.unit_max_bits, out.corr, out.corr_stride, swap);
    }
}

}  // namespace apd
