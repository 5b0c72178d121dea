// Four-step real FFT cross-correlation kernels (Step 1 of the detection path).
//
// Replaces fft_correlation.fft_correlate_1d(section, clip, 'full') + abs + max
// (reference audio_pattern_detector.py:487-494) for a whole batch of
// (chunk x pattern) units at once.
//
// Real transform trick (no mirrored-bin pass): for a real x of length N = 2M,
//     u[m] = (x[m] - i x[m+M]) * e^{-i pi m / N},   U = FFT_M(u)
// gives the odd-frequency spectrum X_o[2g] = U[g].  Point-wise products of two
// such spectra are the spectrum of the *negacyclic* convolution, which equals
// the linear one because N >= S + L - 1.  The inverse undoes the same steps:
//     v = IFFT_M(U_a U_b),  z = v[m] e^{+i pi m / N},  y[m] = Re z, y[m+M] = -Im z.
//
// Complex FFT_M with M = N1*N2 is done as two tiled passes ("four-step"):
//   forward : column pass (N1-point FFTs over a tile of TB adjacent columns, fused with the
//             section load: loudness gain, clip, NaN scrub, packing, pre-twiddle) -> twiddle
//             -> row pass (N2-point FFTs) ; spectrum stored as [c][d], frequency g = c + N1 d.
//   inverse : row pass (fused spectral multiply) -> twiddle -> column pass (fused post-twiddle,
//             |.|, per-unit max or normalised write-out).
// Every global access of every pass is a contiguous run of >= 32 bytes.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "fft_fast.cuh"
#include "internal.h"

namespace apd {

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(512)
k_fwd_cols(Fft4Plan P, SectionGeom G, const double* __restrict__ gains, int gain_stride, float2* __restrict__ T)
{
    extern __shared__ float2 smem[];
    const int TB = 1 << P.tb_log2;
    float2* A = smem;
    float2* B = smem + P.N1 * TB;
    const int sec = blockIdx.y;
    const int b0 = blockIdx.x * TB;
    long long start;
    int n;
    section_bounds(G, sec, start, n);
    const float* __restrict__ x = G.audio + (start - G.base);
    const double gain = gains ? gains[(long long)sec * gain_stride] : 1.0;
    const int M = P.M;
    const float invN = 1.0f / (2.0f * (float)M);
    const int total = P.N1 << P.tb_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int q = idx & (TB - 1);
        const int a = idx >> P.tb_log2;
        const int m = a * P.N2 + b0 + q;
        const float x0 = m < n ? normalize_sample(x[m], gain) : 0.0f;
        const float x1 = m + M < n ? normalize_sample(x[m + M], gain) : 0.0f;
        A[idx] = cmul(make_float2(x0, -x1), cispif(-(float)m * invN));
    }
    __syncthreads();
    const float2* R = sub_fft<-1>(P.col, A, B, P.tb_log2, TB);
    const float invM = 1.0f / (float)M;
    float2* __restrict__ out = T + (long long)sec * M;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int q = idx & (TB - 1);
        const int c = idx >> P.tb_log2;
        const int b = b0 + q;
        out[c * P.N2 + b] = cmul(R[idx], twiddle_frac(b * c, invM, -1.0f));
    }
}

__global__ void __launch_bounds__(512)
k_fwd_rows(Fft4Plan P, const float2* __restrict__ T, float2* __restrict__ spec, long long spec_stride)
{
    extern __shared__ float2 smem[];
    const int TR = 1 << P.tr_log2;
    const int LD = TR + 1;
    float2* A = smem;
    float2* B = smem + P.N2 * LD;
    const int sec = blockIdx.y;
    const int c0 = blockIdx.x * TR;
    const float2* __restrict__ in = T + (long long)sec * P.M + (long long)c0 * P.N2;
    const int total = P.N2 << P.tr_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int e = idx & (P.N2 - 1);
        const int q = idx / P.N2;
        A[e * LD + q] = in[q * P.N2 + e];
    }
    __syncthreads();
    const float2* R = sub_fft<-1>(P.row, A, B, P.tr_log2, LD);
    float2* __restrict__ out = spec + (long long)sec * spec_stride + (long long)c0 * P.N2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int d = idx & (P.N2 - 1);
        const int q = idx / P.N2;
        out[q * P.N2 + d] = R[d * LD + q];
    }
}

__global__ void __launch_bounds__(512)
k_inv_rows(Fft4Plan P, const float2* __restrict__ spec, long long spec_stride, UnitSrc U, UnitCtx C,
           float2* __restrict__ W)
{
    extern __shared__ float2 smem[];
    const int TR = 1 << P.tr_log2;
    const int LD = TR + 1;
    float2* A = smem;
    float2* B = smem + P.N2 * LD;
    const int u = blockIdx.y;
    int2 unit;
    if (!get_unit(U, u, &unit)) return;
    const int c0 = blockIdx.x * TR;
    const float2* __restrict__ xs = spec + (long long)unit.x * spec_stride + C.clip_spec_off[unit.y] +
                                    (long long)c0 * P.N2;
    const float2* __restrict__ hs = C.clip_spec[unit.y] + (long long)c0 * P.N2;
    const int total = P.N2 << P.tr_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int e = idx & (P.N2 - 1);
        const int q = idx / P.N2;
        A[e * LD + q] = cmul(xs[q * P.N2 + e], __ldg(&hs[q * P.N2 + e]));
    }
    __syncthreads();
    const float2* R = sub_fft<+1>(P.row, A, B, P.tr_log2, LD);
    const float invM = 1.0f / (float)P.M;
    float2* __restrict__ out = W + (long long)u * P.M + (long long)c0 * P.N2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int b = idx & (P.N2 - 1);
        const int q = idx / P.N2;
        out[q * P.N2 + b] = cmul(R[b * LD + q], twiddle_frac(b * (c0 + q), invM, +1.0f));
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(512)
k_inv_cols(Fft4Plan P, UnitCtx C, UnitSrc U, const float2* __restrict__ W, InvOut O)
{
    extern __shared__ float2 smem[];
    __shared__ float red[16];
    const int TB = 1 << P.tb_log2;
    float2* A = smem;
    float2* B = smem + P.N1 * TB;
    const int u = blockIdx.y;
    int2 unit;
    if (!get_unit(U, u, &unit)) return;
    const int b0 = blockIdx.x * TB;
    const int M = P.M;
    const float2* __restrict__ in = W + (long long)u * M;
    const int total = P.N1 << P.tb_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int q = idx & (TB - 1);
        const int c = idx >> P.tb_log2;
        A[idx] = in[c * P.N2 + b0 + q];
    }
    __syncthreads();
    const float2* R = sub_fft<+1>(P.col, A, B, P.tb_log2, TB);

    long long start;
    int n;
    section_bounds(C.geoms[C.clip_group[unit.y]], unit.x, start, n);
    const int n_out = n > 0 ? n + C.clip_len[unit.y] - 1 : 0;
    const float invN = 1.0f / (2.0f * (float)M);
    const float invM = 1.0f / (float)M;
    float mc = 1.0f;
    if (WRITE) {
        const float um = __uint_as_float(O.unit_max_bits[(long long)unit.x * O.n_clips + unit.y]);
        mc = fmaxf(O.self_max[unit.y], um);                       // apd.py:493
    }
    float* __restrict__ corr = WRITE ? O.corr + (long long)u * O.corr_stride : nullptr;
    float best = 0.0f;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int q = idx & (TB - 1);
        const int a = idx >> P.tb_log2;
        const int m = a * P.N2 + b0 + q;
        const float2 z = cmul(R[idx], cispif((float)m * invN));
        const float y0 = fabsf(z.x * invM);
        const float y1 = fabsf(z.y * invM);
        if (WRITE) {
            if (m < n_out) corr[m] = y0 / mc;                      // apd.py:494 (float32 divide)
            if (m + M < n_out) corr[m + M] = y1 / mc;
        } else {
            if (m < n_out) best = fmaxf(best, y0);
            if (m + M < n_out) best = fmaxf(best, y1);
        }
    }
    if (!WRITE) {
        best = warp_max(best);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
            v = warp_max(v);
            if (threadIdx.x == 0)
                atomicMax(&O.unit_max_bits[(long long)unit.x * O.n_clips + unit.y], __float_as_uint(v));
        }
    }
}

// ------------------------------------------------------------------ fast kernels (hot shapes)
// Forward transforms of the hot shapes (the inverse ones live in corr_inv.cu).
// Row kernels handle N2 = 512 = 8*8*8 with 64 threads per row (one radix-8 butterfly per thread per
// pass); column kernels handle N1 in {512 = 8*8*8, 576 = 8*8*9, 640 = 8*8*10} with one thread per butterfly and
// TB adjacent columns on the lanes.  All complex arithmetic is packed (cpx2.cuh, fft_fast.cuh).
constexpr int kRowN = 512;
constexpr int kRowPitch = kRowN;
__device__ __forceinline__ int rpad(int e) { return e ^ ((e >> 3) & 15); }
template <int TB> __device__ __forceinline__ int cswz(int e) { return TB == 8 ? (e ^ ((e >> 3) & 1)) : (e ^ ((e >> 3) & 3)); }
__device__ __forceinline__ c2 ldg_c2(const c2* p)
{
    c2 r;
    r.v = __ldg(reinterpret_cast<const unsigned long long*>(p));
    return r;
}

// Every fast kernel fixes its tile (rows c0.. / columns b0..) and loops over `per` consecutive
// transforms of the batch (units or sections), so all twiddles are loop invariants in registers.
// In the row kernels the 64 threads (two warps) of a row synchronise among themselves only
// (named barrier q + 1), so the rows of a CTA drift apart and overlap each other's memory waits.
template <int TR>
__global__ void __launch_bounds__(TR * 64)
k_fwd_rows_fast(Fft4Plan P, const float2* __restrict__ T, float2* __restrict__ spec, long long spec_stride,
                FwdGroups FG, int ntr, int per)
{
    __shared__ c2 buf[TR * kRowPitch];
    const int q = threadIdx.x >> 6, j = threadIdx.x & 63;
    const int c = blockIdx.x * TR + q;
    c2* b = buf + q * kRowPitch;
    float2 tw2[8], tw3[8];
    pass_twiddles<8, -1, 8>(j, tw2);
    pass_twiddles<8, -1, 64>(j, tw3);
    const int t_end = min(ntr, (int)(blockIdx.y + 1) * per);
    for (int t = blockIdx.y * per; t < t_end; ++t) {
        const int sec = FG.ng ? t / FG.ng : t;
        const long long goff = FG.ng ? FG.spec_off[t - sec * FG.ng] : 0;
        const c2* __restrict__ in = reinterpret_cast<const c2*>(T + (long long)t * P.M + (long long)c * kRowN);
        c2 v[8];
        bfly_ld<8, kRowN>(j, [&](int e) { return in[e]; }, v);
        Dft2<8, -1>::run(v);
        bfly_store<8, 1>(j, [&](int e, c2 x) { b[rpad(e)] = x; }, v);
        group_sync<64>(q + 1);
        bfly_ld<8, kRowN>(j, [&](int e) { return b[rpad(e)]; }, v);
        group_sync<64>(q + 1);
        bfly_tw<8>(v, tw2);
        Dft2<8, -1>::run(v);
        bfly_store<8, 8>(j, [&](int e, c2 x) { b[rpad(e)] = x; }, v);
        group_sync<64>(q + 1);
        bfly_ld<8, kRowN>(j, [&](int e) { return b[rpad(e)]; }, v);
        group_sync<64>(q + 1);
        bfly_tw<8>(v, tw3);
        Dft2<8, -1>::run(v);
        c2* __restrict__ out = reinterpret_cast<c2*>(spec + (long long)sec * spec_stride + goff + (long long)c * kRowN);
#pragma unroll
        for (int r = 0; r < 8; ++r) out[j + 64 * r] = v[r];
    }
}

// Column kernel: S::N = N1 = 8 * 8 * R2, TB adjacent columns on the lanes, threads = TB * (N1 / 8).
// Fused with the section load: loudness gain, clamp, NaN scrub (lib.rs:220-227, apd.py:489-490), packing
// u[m] = x[m] - i x[m + M] and the pre-twiddle e^{-i pi m / N}.  Exchange addressing: ColAddr / ColLoad.
// Threads: TB per butterfly; passes 1-2 have N1 / 8 butterflies per column, pass 3 always 64 (N1 = 384, 448: fewer than 64
// in passes 1-2, the surplus threads only keep the barriers there).
template <class S, int TB> struct FwdThreads { static constexpr int value = TB * (S::N / 8 > 64 ? S::N / 8 : 64); };

template <class S, int TB>
__global__ void __launch_bounds__(FwdThreads<S, TB>::value, FwdThreads<S, TB>::value <= 320 ? 2 : 1)
k_fwd_cols_fast(Fft4Plan P, SectionGeom G, FwdGroups FG, const double* __restrict__ gains, int gain_stride,
                float2* __restrict__ T, int ntr, int per)
{
    constexpr int N1 = S::N;
    constexpr int T1 = N1 / 8;
    constexpr int R2 = S::R2;
    constexpr int NLAST = N1 / R2;
    extern __shared__ __align__(1024) unsigned char fwd_cols_smem[];           // two exchange buffers (dynamic: > 48 KB for TB = 8)
    c2* raw = reinterpret_cast<c2*>(fwd_cols_smem);
    const int q = threadIdx.x % TB, j = threadIdx.x / TB;
    const bool act = T1 >= 64 || j < T1;                     // owns a butterfly of passes 1-2
    const int bcol = blockIdx.x * TB + q;
    const int M = P.M;
    const ColAddr<TB> A(raw, j, q);
    float2 pre[8], tw2[8], tw3[R2], fs[R2];
    {
        // pre-twiddle e^{-i pi m / N}, m = (j + r N1/8) 512 + b  = base * (e^{-i pi / 16})^r
        const float invN = 1.0f / (2.0f * (float)M);
        geometric<8>(cispif(-(float)(j * kRowN + bcol) * invN), cispif(-0.5f / 8.0f), pre);
        pass_twiddles<8, -1, 8>(j, tw2);
        pass_twiddles<R2, -1, 64>(j, tw3);
        // outputs c = j + r * NLAST ; four-step twiddle w_M^{-b c} = base * step^r
        const float invM = 1.0f / (float)M;
        geometric<R2>(twiddle_frac(bcol * (j % NLAST), invM, -1.0f), twiddle_frac(bcol * NLAST, invM, -1.0f), fs);
    }
    const int m_lo = j * kRowN + bcol;                       // first-pass input r: m = m_lo + r * T1 * 512
    const int t_end = min(ntr, (int)(blockIdx.y + 1) * per);
    constexpr unsigned kBufB = N1 * TB * 8u;                 // second exchange buffer (pass 2 -> 3)
    // software pipeline: the raw samples (and gain) of transform t + 1 are requested right after the first
    // exchange of transform t
    float x0[8], x1[8];
    double gain = 1.0;
    auto fetch = [&](int t) {
        const int sec = FG.ng ? t / FG.ng : t;
        int gcol = 0;
        if (FG.ng) {
            const int gi = t - sec * FG.ng;
            G.halo = FG.halo[gi];
            gcol = FG.gain_index[gi];
        }
        long long start;
        int n;
        section_bounds(G, sec, start, n);
        const float* __restrict__ x = G.audio + (start - G.base) + m_lo;
        gain = gains ? gains[(long long)sec * gain_stride + gcol] : 1.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int m = m_lo + r * T1 * kRowN;
            x0[r] = m < n ? x[r * T1 * kRowN] : 0.0f;
            x1[r] = m + M < n ? x[r * T1 * kRowN + M] : 0.0f;
        }
    };
    if (act && (int)(blockIdx.y * per) < t_end) fetch(blockIdx.y * per);
    for (int t = blockIdx.y * per; t < t_end; ++t) {
        c2 v[R2 > 8 ? R2 : 8];
        if (act) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                v[r] = cmul(mk(normalize_sample(x0[r], gain), -normalize_sample(x1[r], gain)), pre[r]);
            Dft2<8, -1>::run(v);
            col_store1<TB>(A, v);
            if (t + 1 < t_end) fetch(t + 1);
        }
        __syncthreads();
        if (act) {
            ColLoad<TB, T1, 8>::run(A, v);
            bfly_tw<8>(v, tw2);
            Dft2<8, -1>::run(v);
            col_store2<TB, kBufB>(A, v);
        }
        __syncthreads();
        if (T1 <= 64 || j < NLAST) {
            ColLoad<TB, 64, R2, kBufB>::run(A, v);
            bfly_tw<R2>(v, tw3);
            Dft2<R2, -1>::run(v);
            c2* __restrict__ out = reinterpret_cast<c2*>(T + (long long)t * M + (long long)j * kRowN + bcol);
#pragma unroll
            for (int r = 0; r < R2; ++r) out[(long long)r * NLAST * kRowN] = cmul(v[r], fs[r]);
        }
    }
}

constexpr int kFastTR = 4;     // rows per CTA  (256 threads)

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
// tuning knobs (defaults chosen on B200; overridable for experiments)
static int fast_tb() { static int v = env_int("APD_B200_TB", 8) == 4 ? 4 : 8; return v; }      // columns per forward CTA
static int fast_per_max() { static int v = std::max(1, env_int("APD_B200_PER", 8)); return v; }  // transforms per CTA

template <int TB>
static void launch_fwd_cols_fast(int fs, const Fft4Plan& P, const SectionGeom& G, const FwdGroups& FG,
                                 const double* gains, int gain_stride, float2* scratch, int ntr, int per, int ny,
                                 cudaStream_t st)
{
    dim3 gc(P.N2 / TB, ny);
    const size_t sm384 = (size_t)(2 * 384 * TB + ColLayout<TB>::SLACK) * sizeof(c2);
    const size_t sm448 = (size_t)(2 * 448 * TB + ColLayout<TB>::SLACK) * sizeof(c2);
    const size_t sm512 = (size_t)(2 * 512 * TB + ColLayout<TB>::SLACK) * sizeof(c2);
    const size_t sm576 = (size_t)(2 * 576 * TB + ColLayout<TB>::SLACK) * sizeof(c2);
    const size_t sm640 = (size_t)(2 * 640 * TB + ColLayout<TB>::SLACK) * sizeof(c2);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_fwd_cols_fast<Shape384, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm384);
        cudaFuncSetAttribute(k_fwd_cols_fast<Shape448, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm448);
        cudaFuncSetAttribute(k_fwd_cols_fast<Shape512, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm512);
        cudaFuncSetAttribute(k_fwd_cols_fast<Shape576, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm576);
        cudaFuncSetAttribute(k_fwd_cols_fast<Shape640, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm640);
        attr = true;
    }
    if (fs == 384)
        k_fwd_cols_fast<Shape384, TB><<<gc, TB * 64, sm384, st>>>(P, G, FG, gains, gain_stride, scratch, ntr, per);
    else if (fs == 448)
        k_fwd_cols_fast<Shape448, TB><<<gc, TB * 64, sm448, st>>>(P, G, FG, gains, gain_stride, scratch, ntr, per);
    else if (fs == 512)
        k_fwd_cols_fast<Shape512, TB><<<gc, TB * 64, sm512, st>>>(P, G, FG, gains, gain_stride, scratch, ntr, per);
    else if (fs == 576)
        k_fwd_cols_fast<Shape576, TB><<<gc, TB * 72, sm576, st>>>(P, G, FG, gains, gain_stride, scratch, ntr, per);
    else
        k_fwd_cols_fast<Shape640, TB><<<gc, TB * 80, sm640, st>>>(P, G, FG, gains, gain_stride, scratch, ntr, per);
}

static int fast_shape(const Fft4Plan& P)
{
    if (P.N2 != kRowN) return 0;
    if (P.N1 == 384) return 384;
    if (P.N1 == 448) return 448;
    if (P.N1 == 512) return 512;
    if (P.N1 == 576) return 576;
    if (P.N1 == 640) return 640;
    return 0;
}

// ------------------------------------------------------------------ plans
static bool factor(int n, SubPlan* sp)
{
    sp->n = n;
    sp->npass = 0;
    int rest = n;
    int odd[kMaxPasses];
    int nodd = 0;
    while (rest % 5 == 0) { if (nodd >= kMaxPasses) return false; odd[nodd++] = 5; rest /= 5; }
    while (rest % 3 == 0) { if (nodd >= kMaxPasses) return false; odd[nodd++] = 3; rest /= 3; }
    while (rest % 8 == 0) { if (sp->npass >= kMaxPasses) return false; sp->radix[sp->npass++] = 8; rest /= 8; }
    if (rest % 4 == 0) { if (sp->npass >= kMaxPasses) return false; sp->radix[sp->npass++] = 4; rest /= 4; }
    if (rest % 2 == 0) { if (sp->npass >= kMaxPasses) return false; sp->radix[sp->npass++] = 2; rest /= 2; }
    if (rest != 1) return false;
    for (int i = 0; i < nodd; ++i) {
        if (sp->npass >= kMaxPasses) return false;
        sp->radix[sp->npass++] = odd[i];
    }
    return true;
}

static bool upload_twiddles(SubPlan* sp)
{
    std::vector<float2> h(sp->n);
    for (int t = 0; t < sp->n; ++t) {
        const double ang = -2.0 * M_PI * (double)t / (double)sp->n;
        h[t] = make_float2((float)cos(ang), (float)sin(ang));
    }
    float2* d = nullptr;
    if (cudaMalloc(&d, sizeof(float2) * sp->n) != cudaSuccess) return false;
    if (cudaMemcpy(d, h.data(), sizeof(float2) * sp->n, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    sp->tw = d;
    return true;
}

long long plan_min_M_for(long long n_out) { return (n_out + 1) / 2; }

bool build_plan(int M_min, Fft4Plan* plan, std::string* err)
{
    // candidate column lengths: 2^a * 3^b * 5^c, a >= 3, b <= 2, c <= 1
    std::vector<int> n1s;
    for (int a = 3; a <= 11; ++a)
        for (int b = 0; b <= 2; ++b)
            for (int c = 0; c <= 1; ++c) {
                long long v = (1LL << a);
                for (int i = 0; i < b; ++i) v *= 3;
                for (int i = 0; i < c; ++i) v *= 5;
                if (v <= 2048) n1s.push_back((int)v);
            }
    n1s.push_back(448);                                      // 8 * 8 * 7: hot-shape kernels only (paired with 512 rows)
    double best_cost = 1e300;
    int bN1 = 0, bN2 = 0;
    for (int n2 = 32; n2 <= 1024; n2 *= 2)
        for (int n1 : n1s) {
            const long long M = (long long)n1 * n2;
            if (M < M_min) continue;
            if (n1 == 448 && n2 != kRowN) continue;
            const double skew = std::fabs(std::log2((double)n1 / (double)n2));
            double cost = (double)M * (1.0 + 0.03 * skew);
            if (n2 == kRowN && (n1 == 384 || n1 == 448 || n1 == 512 || n1 == 576 || n1 == 640))
                cost *= 0.6;                                                            // register-resident fast kernels exist
            if (cost < best_cost) { best_cost = cost; bN1 = n1; bN2 = n2; }
        }
    if (!bN1) {
        if (err) *err = "section + clip too long for the four-step FFT (max 2^22 real samples)";
        return false;
    }
    Fft4Plan P;
    P.N1 = bN1; P.N2 = bN2; P.M = bN1 * bN2;
    // (the generic kernels have no radix-7 pass; a 448 x 512 plan is only ever run by the hot-shape kernels)
    if (!(factor(P.N1, &P.col) || fast_shape(P)) || !factor(P.N2, &P.row)) {
        if (err) *err = "internal: cannot factor sub-FFT length";
        return false;
    }
    // column tile: as wide as fits ~96 KB for both ping-pong buffers, 4..32 columns
    int tb = 5;
    while (tb > 2 && (size_t)2 * P.N1 * (1u << tb) * sizeof(float2) > 96u * 1024u) --tb;
    P.tb_log2 = tb;
    int tr = 3;
    while (tr > 2 && (size_t)2 * P.N2 * ((1u << tr) + 1) * sizeof(float2) > 96u * 1024u) --tr;
    P.tr_log2 = tr;
    P.smem_col = (size_t)2 * P.N1 * (1u << P.tb_log2) * sizeof(float2);
    P.smem_row = (size_t)2 * P.N2 * ((1u << P.tr_log2) + 1) * sizeof(float2);
    P.threads = 256;
    if (!upload_twiddles(&P.col) || !upload_twiddles(&P.row)) {
        if (err) *err = "cudaMalloc failed for twiddle tables";
        return false;
    }
    *plan = P;
    return true;
}

void free_plan(Fft4Plan* plan)
{
    if (plan->col.tw) cudaFree((void*)plan->col.tw);
    if (plan->row.tw) cudaFree((void*)plan->row.tw);
    plan->col.tw = plan->row.tw = nullptr;
}

// ------------------------------------------------------------------ launchers
static void ensure_attrs()
{
    static bool done = false;
    if (done) return;
    const int big = 200 * 1024;
    cudaFuncSetAttribute(k_fwd_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_fwd_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_inv_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_inv_cols<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_inv_cols<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    done = true;
}

void launch_forward(const Fft4Plan& P, const SectionGeom& G, const FwdGroups& FG, const double* gains, int gain_stride,
                    int nsec, float2* scratch, float2* spec, long long spec_stride, cudaStream_t st)
{
    if (nsec <= 0) return;
    ensure_attrs();
    const int fs = fast_shape(P);
    if (fs) {
        const int ntr = nsec * std::max(FG.ng, 1);
        const int tb = fast_tb();
        // keep >= ~4 CTAs per SM in the grid, otherwise amortise the twiddle set-up over many transforms
        int per = std::max(fast_per_max(), 32);
        while (per > 1 && (long long)((ntr + per - 1) / per) * (P.N2 / tb) < 148 * 4) per >>= 1;
        const int ny = (ntr + per - 1) / per;
        if (tb == 8) launch_fwd_cols_fast<8>(fs, P, G, FG, gains, gain_stride, scratch, ntr, per, ny, st);
        else launch_fwd_cols_fast<4>(fs, P, G, FG, gains, gain_stride, scratch, ntr, per, ny, st);
        int perr = std::max(fast_per_max(), 32);
        while (perr > 1 && (long long)((ntr + perr - 1) / perr) * (P.N1 / kFastTR) < 148 * 8) perr >>= 1;
        dim3 gr(P.N1 / kFastTR, (ntr + perr - 1) / perr);
        k_fwd_rows_fast<kFastTR><<<gr, kFastTR * 64, 0, st>>>(P, scratch, spec, spec_stride, FG, ntr, perr);
        return;
    }
    // generic shapes: one group per launch (FG.ng == 0)
    dim3 gc(P.N2 >> P.tb_log2, nsec), gr(P.N1 >> P.tr_log2, nsec);
    k_fwd_cols<<<gc, P.threads, P.smem_col, st>>>(P, G, gains, gain_stride, scratch);
    k_fwd_rows<<<gr, P.threads, P.smem_row, st>>>(P, scratch, spec, spec_stride);
}

void launch_inverse(const Fft4Plan& P, const UnitCtx& C, const float2* spec, long long spec_slab,
                    const UnitSrc& U, int nunits, float2* scratch, void* desc, const InvOut& out, bool write,
                    cudaStream_t st)
{
    if (nunits <= 0) return;
    ensure_attrs();
    if (corr_inv_supported(P) && desc) {
        launch_corr_inv(P, C, spec, spec_slab, U, nunits, scratch, desc, out, write, st);
        return;
    }
    dim3 gr(P.N1 >> P.tr_log2, nunits), gc(P.N2 >> P.tb_log2, nunits);
    k_inv_rows<<<gr, P.threads, P.smem_row, st>>>(P, spec, spec_slab, U, C, scratch);
    if (write) k_inv_cols<true><<<gc, P.threads, P.smem_col, st>>>(P, C, U, scratch, out);
    else k_inv_cols<false><<<gc, P.threads, P.smem_col, st>>>(P, C, U, scratch, out);
}

}  // namespace apd
