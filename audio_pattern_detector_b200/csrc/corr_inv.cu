// K2 -- fused spectral multiply + inverse four-step FFT + |.| + per-unit max (or normalised write-out)
// for the hot shapes M = N1 x 512, N1 in {384, 448, 512, 576, 640} (8 x 8 x R2, R2 = 6 .. 10).
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
// Both kernels keep one radix-8 (last column pass: radix-6 .. radix-10) butterfly per thread in registers, in packed complex
// arithmetic (cpx2.cuh: FADD2/FMUL2/FFMA2), and hand the last pass straight to the epilogue.  All twiddles
// and all shared-memory addresses are loop invariants of the per-CTA loop over units: the XOR-swizzled
// exchange layouts (fft_fast.cuh) reduce to "thread constant + immediate" for loads and "thread constant ^
// immediate" for stores.
//
// What bounds them is the LSU data pipe of the SM (one 128-byte wavefront per cycle shared by shared-memory and
// global accesses), not HBM (profiles/sweeps_r2.txt: with W aliased onto an L2-resident ring the stage time does
// not change).  Hence:
//   * the row pass keeps the clip's spectrum row in registers across the consecutive units of a clip (clip-major
//     unit order), so a unit costs one 64-bit global load per point instead of two;
//   * the column pass gets its N1 x 4 tile of W by ONE TMA tensor copy per unit (cp.async.bulk.tensor.3d on an
//     mbarrier, N1 pieces of 32 bytes) instead of LDG.128s that touched 16 lines (32 bytes of each) per request
//     and cost 16 wavefronts: 88 M -> 60 M data-pipe wavefronts per 512 units (profiles/ncu_r2_*).
// Phase 2 (write == true) runs the same two kernels, so the values it writes are bit-identical to those whose
// maximum phase 1 took: the normalised correlation of a unit whose maximum is the divisor peaks at exactly 1.0f.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// A CTA whose units are all unused slots (phase 2 is launched for a full round of slots without waiting for the
// selected count) leaves before it computes its twiddles.
__device__ __forceinline__ bool any_unit_in_use(const UnitDesc* __restrict__ D, int u_begin, int u_end)
{
    bool any = false;
    for (int u = u_begin; u < u_end; ++u) any |= __ldg(&D[u].n_out) >= 0;
    return any;
}

constexpr int kN2 = 512;          // row length (complex)
constexpr int kRowsPerCta = 4;    // 64 threads per row

// ---------------------------------------------------------------- rows
// 64 threads (two warps) per row, kRowsPerCta rows per CTA; the two warps of a row synchronise among
// themselves only (named barrier q + 1), so the rows of a CTA drift apart.  Two exchange buffers per row
// (pass 1 -> 2 and pass 2 -> 3): one barrier per exchange, none for reuse.
// Exchange layout of a row: slot(e) = e ^ ((e >> 3) & 15).
//   loads  e = j + 64 r            : 8 * (j ^ (j >> 3)) [^ 64 for odd r] + 512 r      (constant + immediate)
//   stores e = 8 j + r      (pass 1): 8 * ((8 j) ^ (j & 15))            ^ (8 r)       (constant ^ immediate)
//   stores e = 64 (j >> 3) + (j & 7) + 8 r (pass 2):
//                                     8 * (64 (j >> 3) + 8 ((j >> 3) & 1) + (j & 7)) ^ (72 r)
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

template <int R, unsigned OFF> __device__ __forceinline__ c2 row_ld(const RowAddr& A)
{
    return (R & 1) ? lds<OFF + 512 * R>(A.ld1) : lds<OFF + 512 * R>(A.ld0);
}

#define APD_ROW_LOAD8(A, OFF, v)                                                                              \
    v[0] = row_ld<0, OFF>(A); v[1] = row_ld<1, OFF>(A); v[2] = row_ld<2, OFF>(A); v[3] = row_ld<3, OFF>(A);  \
    v[4] = row_ld<4, OFF>(A); v[5] = row_ld<5, OFF>(A); v[6] = row_ld<6, OFF>(A); v[7] = row_ld<7, OFF>(A);

// grid: (unit groups of `per`, N1 / kRowsPerCta); CTA = rows [4 tile, 4 tile + 4) of units [bx per, (bx + 1) per).
// Software pipeline: the section-spectrum row of unit u + 1 is requested right after the first exchange of
// unit u (ld.global.L1::no_allocate), so its latency is covered by two passes of arithmetic; the clip's row is
// re-loaded only when the clip changes (consecutive units of a launch share the clip).
__global__ void __launch_bounds__(kRowsPerCta * 64, 3)
k_corr_rows(const UnitDesc* __restrict__ D, int nunits, int per, int M, float2* __restrict__ W)
{
    __shared__ __align__(1024) c2 buf[kRowsPerCta * 2 * kN2];
    const int tile = blockIdx.y, u_begin = blockIdx.x * per, u_end = min(nunits, u_begin + per);
    if (!any_unit_in_use(D, u_begin, u_end)) return;
    const int q = threadIdx.x >> 6, j = threadIdx.x & 63;
    const int c = tile * kRowsPerCta + q;
    const RowAddr A(smem_addr(buf + q * (2 * kN2)), j);
    constexpr unsigned kB = kN2 * 8u;                      // byte offset of the second buffer
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
    UnitDesc dn = load_desc(D + u_begin);
    if (dn.n_out >= 0) {
        const c2* __restrict__ xs = reinterpret_cast<const c2*>(dn.xs + row_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) xn[r] = ldg_stream(xs + 64 * r);
    }
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
        if (d.hs != cur_hs) {
            const c2* __restrict__ hs = reinterpret_cast<const c2*>(d.hs + row_off);
#pragma unroll
            for (int r = 0; r < 8; ++r) h[r] = ldg_nc(hs + 64 * r);
            cur_hs = d.hs;
        }
        c2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = cmul(xn[r], h[r]);
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
        c2* __restrict__ out = reinterpret_cast<c2*>(W + (long long)u * M + row_off);
#pragma unroll
        for (int r = 0; r < 8; ++r) out[64 * r] = cmul(v[r], fs[r]);
    }
}

// ---------------------------------------------------------------- columns
// kTB = 4 adjacent columns per CTA, two per thread (threads = 2 * N1 / 8): the pair is adjacent in the swizzled
// exchange layout (ColAddr / ColLoad2, fft_fast.cuh), so exchanges move 16 bytes per instruction, the twiddles are
// shared by two butterflies and a CTA does twice the arithmetic between barriers.  The post-twiddle
// e^{+i pi m / N} / M with m = (j + 64 r) 512 + b factors into a per-thread scalar, folded into the last pass's
// input twiddles, and per-output constants e^{i pi (r / (2 R2) + col / N)} kept in constant memory.
constexpr int kTB = 4;
__constant__ float2 c_post[5][2][10];        // [shape: 384, 448, 512, 576, 640][column of the pair][r]

template <class S> struct ShapeIndex { static constexpr int value = S::R2 - 6; };

__device__ __forceinline__ void tma_load_tile3(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bytes,
                                               unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// best = max(best, y) if a < b (one ISETP + one predicated FMNMX; y >= 0)
__device__ __forceinline__ float max_if_lt(float best, float y, int a, int b)
{
    asm("{\n.reg .pred p;\nsetp.lt.s32 p, %2, %3;\n@p max.f32 %0, %0, %1;\n}" : "+f"(best) : "f"(y), "r"(a), "r"(b));
    return best;
}

// grid: (unit groups of `per`, 512 / kTB); CTA = columns [4 tile, 4 tile + 4) of units [bx per, (bx + 1) per).
// tmap: W seen as [8 * units][N1 / 8][512] 8-byte elements (row c = j + (N1 / 8) r is element (., j, r)); the box
// {4, N1 / 8, 8} lands in shared memory as [c][4 columns].  NS = 1: the tile of the next unit is requested right after
// the first exchange of the current one (the input buffer is free then), i.e. two passes ahead; NS = 2: a unit ahead.
// two threads per butterfly; passes 1-2 have N1 / 8 butterflies per column, pass 3 always 64 (N1 = 384, 448: fewer
// than 64 in passes 1-2; 576, 640: more)
template <class S> struct ColThreads {
    static constexpr int value = (2 * (S::N / 8 > 64 ? S::N / 8 : 64) + 31) / 32 * 32;                    // whole warps
};

// input tiles in flight per CTA: one (the next unit's tile is requested after the first exchange), which lets four
// CTAs of N1 = 512 (3 x 16 KB) share an SM; two (a whole unit ahead) for N1 = 576, where a fourth CTA does not fit
// but a fourth buffer does (measured equal to one)
template <class S> struct ColStages { static constexpr int value = S::N == 576 ? 2 : 1; };
template <class S, bool WRITE> struct ColCtas { static constexpr int value = WRITE ? 2 : (S::N <= 512 ? 4 : 3); };

template <class S, bool WRITE, int NS>
__global__ void __launch_bounds__(ColThreads<S>::value, ColCtas<S, WRITE>::value)
k_corr_cols(const __grid_constant__ CUtensorMap tmap, const UnitDesc* __restrict__ D, int nunits, int per, int M,
            unsigned int* __restrict__ unit_max_bits, float* __restrict__ corr, long long corr_stride)
{
    constexpr int N1 = S::N;
    constexpr int T1 = N1 / 8;
    constexpr int R2 = S::R2;
    constexpr int NLAST = N1 / R2;
    static_assert(NLAST == 64, "last pass: 64 butterflies per column");
    constexpr int NW = ColThreads<S>::value / 32;
    constexpr int SH = ShapeIndex<S>::value;
    constexpr unsigned kBufB = N1 * kTB * 8u;
    constexpr unsigned kTileBytes = N1 * kTB * 8u;
    extern __shared__ unsigned char cols_smem[];
    __shared__ float red[2 * NW];
    __shared__ __align__(8) unsigned long long bar[NS];
    // [exchange buffers: 2 * N1 * kTB, aligned to ColLayout::ALIGN][input tiles: NS * N1 * kTB]
    constexpr uintptr_t AL = ColLayout<kTB>::ALIGN;
    c2* raw = reinterpret_cast<c2*>((reinterpret_cast<uintptr_t>(cols_smem) + AL - 1) & ~(AL - 1));
    const unsigned in0 = smem_addr(raw + 2 * N1 * kTB);
    const int tile = blockIdx.y, u_begin = blockIdx.x * per, u_end = min(nunits, u_begin + per);
    if (!any_unit_in_use(D, u_begin, u_end)) return;
    const int p = threadIdx.x & 1, j = threadIdx.x >> 1;
    const bool act = 2 * T1 == ColThreads<S>::value || j < T1;  // threads past the butterflies of passes 1-2 keep the barriers
    constexpr bool kAllLast = 2 * NLAST == ColThreads<S>::value;  // every thread owns a pass-3 butterfly
    const int bcol = tile * kTB + 2 * p;                     // first column of the pair
    const ColAddr<kTB> A(raw, j, 2 * p);
    const unsigned ld_in0 = in0 + 32u * (unsigned)j + 16u * (unsigned)p;
    const unsigned bar0 = smem_addr(bar);
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
    int pending = -1, parity = 0;                            // unit whose per-warp maxima wait in red[parity ^ 1]
    int un = u_begin;                                        // next unit whose tile has not been requested (thread 0)
    auto request_next = [&](int stage) {                     // thread 0: tile of the next unit in use -> input buffer `stage`
        while (un < u_end && load_desc(D + un).n_out < 0) ++un;
        if (un < u_end) {
            tma_load_tile3(in0 + stage * kTileBytes, &tmap, tile * kTB, 0, un * 8, kTileBytes, bar0 + 8u * stage);
            ++un;
        }
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8u * s, 1);
        mbar_init_fence();
        request_next(0);
    }
    __syncthreads();
    int k = 0;                                               // units processed so far (buffer and mbarrier phase)
    for (int u = u_begin; u < u_end; ++u) {
        const UnitDesc d = load_desc(D + u);
        if (d.n_out < 0) continue;
        const int sb = k % NS;
        // the other buffer was last read in the first pass of the previous unit: free since that unit's barriers
        if (NS == 2 && threadIdx.x == 0) request_next(sb ^ 1);
        mbar_wait(bar0 + 8u * sb, (unsigned)((k / NS) & 1));
        c2 va[R2 > 8 ? R2 : 8], vb[R2 > 8 ? R2 : 8];
        const unsigned ld_in = ld_in0 + sb * kTileBytes;
        if (act) {
            lds128<0 * 32 * T1>(ld_in, va[0], vb[0]); lds128<1 * 32 * T1>(ld_in, va[1], vb[1]);
            lds128<2 * 32 * T1>(ld_in, va[2], vb[2]); lds128<3 * 32 * T1>(ld_in, va[3], vb[3]);
            lds128<4 * 32 * T1>(ld_in, va[4], vb[4]); lds128<5 * 32 * T1>(ld_in, va[5], vb[5]);
            lds128<6 * 32 * T1>(ld_in, va[6], vb[6]); lds128<7 * 32 * T1>(ld_in, va[7], vb[7]);
            Dft2<8, +1>::run(va);
            Dft2<8, +1>::run(vb);
            col_store1_x2<kTB>(A, va, vb);
        }
        __syncthreads();
        if (NS == 1 && threadIdx.x == 0) request_next(0);    // every thread has consumed the input tile
        if (!WRITE && pending >= 0 && threadIdx.x < 32) {    // block maximum of the previous unit (see below)
            float t = threadIdx.x < NW ? red[(parity ^ 1) * NW + threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(unit_max_bits + pending, __float_as_uint(t));
        }
        if (act) {
            ColLoad2<kTB, T1, 8>::run(A, va, vb);
            bfly_tw<8>(va, tw2);
            bfly_tw<8>(vb, tw2);
            Dft2<8, +1>::run(va);
            Dft2<8, +1>::run(vb);
            col_store2_x2<kTB, kBufB>(A, va, vb);
        }
        __syncthreads();
        float best = 0.0f;
        if (kAllLast || j < NLAST) {
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
            // per-warp maxima go to red[parity]; they are combined after the NEXT barrier the CTA passes anyway
            // (the first exchange of the next unit, or the one after the loop), so a unit costs two barriers
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

// ---------------------------------------------------------------- launcher
static int env_int2(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool corr_inv_supported(const Fft4Plan& P)
{
    return P.N2 == kN2 && (P.N1 == 384 || P.N1 == 448 || P.N1 == 512 || P.N1 == 576 || P.N1 == 640);
}

size_t corr_inv_desc_bytes(int nunits) { return sizeof(UnitDesc) * (size_t)(nunits > 0 ? nunits : 1); }

// Tensor maps over the intermediate, one per (buffer, N1): cuTensorMapEncodeTiled through the runtime's driver
// entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static const CUtensorMap* w_tensor_map(float2* scratch, int N1)
{
    static std::map<std::pair<void*, int>, CUtensorMap> cache;
    static EncodeTiledFn encode = nullptr;
    const auto key = std::make_pair((void*)scratch, N1);
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return nullptr;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t dims[3] = {512ull, (cuuint64_t)(N1 / 8), 8ull * 65536ull};
    const cuuint64_t strides[2] = {512ull * 8ull, (cuuint64_t)(N1 / 8) * 512ull * 8ull};
    const cuuint32_t box[3] = {(cuuint32_t)kTB, (cuuint32_t)(N1 / 8), 8u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUtensorMap m;
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, scratch, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return nullptr;
    return &cache.emplace(key, m).first->second;
}

static void upload_post_constants()
{
    static bool uploaded = false;
    if (uploaded) return;
    float2 h[5][2][10];
    const int r2[5] = {6, 7, 8, 9, 10};
    const double n[5] = {2.0 * 384 * 512, 2.0 * 448 * 512, 2.0 * 512 * 512, 2.0 * 576 * 512, 2.0 * 640 * 512};
    for (int sh = 0; sh < 5; ++sh)
        for (int col = 0; col < 2; ++col)
            for (int r = 0; r < 10; ++r) {
                const double a = M_PI * ((double)r / (2.0 * r2[sh]) + (double)col / n[sh]);
                h[sh][col][r] = make_float2((float)cos(a), (float)sin(a));
            }
    cudaMemcpyToSymbol(c_post, h, sizeof(h));
    uploaded = true;
}

template <class S, bool WRITE, int NS = WRITE ? 1 : ColStages<S>::value>
static void launch_cols(const CUtensorMap& map, const UnitDesc* D, int nunits, int per, int M, const InvOut& out, dim3 grid,
                        cudaStream_t st)
{
    constexpr size_t smem = (size_t)(2 + NS) * S::N * kTB * sizeof(c2) + ColLayout<kTB>::ALIGN;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_corr_cols<S, WRITE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    k_corr_cols<S, WRITE, NS><<<grid, ColThreads<S>::value, smem, st>>>(map, D, nunits, per, M, out.unit_max_bits, out.corr, out.corr_stride);
}

void launch_corr_inv(const Fft4Plan& P, const UnitCtx& C, const float2* spec, long long spec_slab, const UnitSrc& U,
                     int nunits, float2* scratch, void* desc, const InvOut& out, bool write, cudaStream_t st)
{
    static const int per_max = std::max(1, env_int2("APD_B200_PER", 16));
    UnitDesc* D = static_cast<UnitDesc*>(desc);
    k_unit_desc<<<(nunits + 127) / 128, 128, 0, st>>>(U, C, spec, spec_slab, out, nunits, write ? 1 : 0, D);
    const CUtensorMap* map = w_tensor_map(scratch, P.N1);
    if (!map) { fprintf(stderr, "apd_b200: cuTensorMapEncodeTiled failed\n"); abort(); }
    upload_post_constants();
    // keep at least ~8 CTAs per SM in the grid; otherwise amortise the twiddles over up to `per` units
    int per = per_max;
    while (per > 1 && (long long)((nunits + per - 1) / per) * 64 < 148 * 8) per >>= 1;
    const int ny = (nunits + per - 1) / per;
    // unit groups fastest: the CTAs of a tile run together, so the tile's clip rows stay in L1/L2 across the launch
    const dim3 gr(ny, P.N1 / kRowsPerCta), gc(ny, kN2 / kTB);
    k_corr_rows<<<gr, kRowsPerCta * 64, 0, st>>>(D, nunits, per, P.M, scratch);
    if (P.N1 == 384) {
        if (write) launch_cols<Shape384, true>(*map, D, nunits, per, P.M, out, gc, st);
        else launch_cols<Shape384, false>(*map, D, nunits, per, P.M, out, gc, st);
    } else if (P.N1 == 448) {
        if (write) launch_cols<Shape448, true>(*map, D, nunits, per, P.M, out, gc, st);
        else launch_cols<Shape448, false>(*map, D, nunits, per, P.M, out, gc, st);
    } else if (P.N1 == 512) {
        if (write) launch_cols<Shape512, true>(*map, D, nunits, per, P.M, out, gc, st);
        else launch_cols<Shape512, false>(*map, D, nunits, per, P.M, out, gc, st);
    } else if (P.N1 == 576) {
        if (write) launch_cols<Shape576, true>(*map, D, nunits, per, P.M, out, gc, st);
        else launch_cols<Shape576, false>(*map, D, nunits, per, P.M, out, gc, st);
    } else {
        if (write) launch_cols<Shape640, true>(*map, D, nunits, per, P.M, out, gc, st);
        else launch_cols<Shape640, false>(*map, D, nunits, per, P.M, out, gc, st);
    }
}

}  // namespace apd
