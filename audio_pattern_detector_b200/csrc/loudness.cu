// K0: BS.1770-4 gated integrated loudness of every section of a batch -> normalisation gain.
//
// Replaces native-helper integrated_loudness (reference lib.rs:75-214) +
// the gain of loudness_normalize (lib.rs:220-223) as called at
// audio_pattern_detector.py:414-420 (sections) and :166-171 (clips).
//
// The reference runs the two K-weighting biquads as ONE serial float64 recurrence
// over the whole section.  Here the section is cut into cells of `cell` samples
// (cell divides the 100 ms gating hop):
//   pass A  every cell is filtered from zero state        -> end state z_i   (parallel)
//   scan    s_i = Mc s_{i-1} + z_i, Mc = A^cell (4x4)     -> true carry-in   (warp scan)
//   pass B  every cell is filtered again from its carry   -> cell energy E_i (parallel)
//   gate    400 ms blocks = 4*k consecutive cells, absolute (-70 LUFS) and relative
//           (-10 LU) gates, LUFS, gain = 10^((-16 - LUFS)/20)        (one warp per section)
// which is the same linear recurrence re-associated, all in float64.
//
// Sharing across sliding-window groups: all groups of a chunk end at the same sample and differ only
// in how far they look back, so passes A/scan/B run ONCE per chunk over the longest look-back
// ("union" section).  A shorter group starts from zero state at a later sample t0; the K-weighting
// filters forget their state within `patch_cells` cells (n r^n < 1e-20 for the slowest pole), so only
// those first cells are re-filtered from zero state (k_kw_patch) and every later cell energy is, to
// float64 rounding, the union's.
//
// The cell scheme needs the gating hop 0.1*sr to be a whole number of samples (sr % 10 == 0) so that block edges
// fall on cell edges; other rates take the serial general-rate kernel at the end of this file (k_kw_serial).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "internal.h"
#include "loudness.h"

namespace apd {

__device__ __forceinline__ void kw_step(const double* cf, double x, double& s1, double& s2, double& h1,
                                        double& h2, double& v)
{
    const double u = cf[0] * x + s1;
    s1 = cf[1] * x - cf[4] * u + s2;
    s2 = cf[2] * x - cf[5] * u;
    v = cf[6] * u + h1;
    h1 = cf[7] * u - cf[10] * v + h2;
    h2 = cf[8] * u - cf[11] * v;
}

// Staging of a CTA's samples into tile[cell][cells + 1 pad]: 4-byte asynchronous copies, so that all of a thread's
// loads are in flight at once (a plain load -> store loop exposed one DRAM latency per element and made both passes
// latency-bound at a tenth of the FP64 rate).
__device__ __forceinline__ void stage_cells(float* tile, const float* __restrict__ x, int cnt, int cell)
{
    const int ld = cell + 1;
    int row = 0, col = threadIdx.x;
    while (col >= cell) { col -= cell; ++row; }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(tile);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + 4u * (unsigned)(row * ld + col)),
                     "l"(x + i) : "memory");
        col += blockDim.x;
        while (col >= cell) { col -= cell; ++row; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}

// PASS 0: zero-state run, writes end state into state[sec][cell][4].  The end state of a zero-state run is linear
//         in the samples: z = sum_t x_t g[len-1-t], g[j] = the state j samples after a unit impulse (K.imp holds g
//         reversed).  Four independent accumulators instead of the serial recurrence.
// PASS 1: run from state[sec][cell] (carry-in), writes energy[sec][cell]
//         (and energy_m1[sec] = energy of the last cell without its final sample).
template <int PASS>
__global__ void __launch_bounds__(128)
k_kw_cells(KwConfig K, SectionGeom G, int cells_stride, double* __restrict__ state,
           double* __restrict__ energy, double* __restrict__ energy_m1)
{
    extern __shared__ float tile[];
    const int sec = blockIdx.y;
    long long start;
    int n;
    section_bounds(G, sec, start, n);
    const int cell = K.cell;
    const int first = blockIdx.x * blockDim.x;              // first cell of this CTA
    const long long s0 = (long long)first * cell;           // first sample (section relative)
    if (s0 >= n) return;
    const int cnt = (int)min((long long)blockDim.x * cell, (long long)n - s0);
    const int ld = cell + 1;
    stage_cells(tile, G.audio + (start - G.base) + s0, cnt, cell);
    const int ci = first + threadIdx.x;
    const int lo = threadIdx.x * cell;
    if (lo >= cnt) return;
    const int len = min(cell, cnt - lo);
    const long long slot = ((long long)sec * cells_stride + ci);
    const float* row = tile + threadIdx.x * ld;
    if (PASS == 0) {
        const double2* __restrict__ g = reinterpret_cast<const double2*>(K.imp) + (size_t)(cell - len) * 2;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
        for (int t = 0; t < len; ++t) {
            const double xd = (double)row[t];
            const double2 g01 = __ldg(g + 2 * t), g23 = __ldg(g + 2 * t + 1);
            a0 = fma(xd, g01.x, a0);
            a1 = fma(xd, g01.y, a1);
            a2 = fma(xd, g23.x, a2);
            a3 = fma(xd, g23.y, a3);
        }
        double2* sp = reinterpret_cast<double2*>(state + slot * 4);
        sp[0] = make_double2(a0, a1);
        sp[1] = make_double2(a2, a3);
    } else {
        const double2* sp = reinterpret_cast<const double2*>(state + slot * 4);
        const double2 c01 = sp[0], c23 = sp[1];
        double s1 = c01.x, s2 = c01.y, h1 = c23.x, h2 = c23.y, v;
        double e = 0, e_prev = 0;
        for (int t = 0; t < len; ++t) {
            kw_step(K.cf, (double)row[t], s1, s2, h1, h2, v);
            e_prev = e;
            e += v * v;
        }
        energy[slot] = e;
        if ((long long)ci * cell + len == n) energy_m1[sec] = e_prev;
    }
}

__device__ __forceinline__ void mat4_apply_add(const double* __restrict__ Mx, const double t[4], double s[4])
{
#pragma unroll
    for (int r = 0; r < 4; ++r)
        s[r] += Mx[r * 4 + 0] * t[0] + Mx[r * 4 + 1] * t[1] + Mx[r * 4 + 2] * t[2] + Mx[r * 4 + 3] * t[3];
}

__device__ __forceinline__ void mat4_vec(const double* __restrict__ Mx, const double t[4], double r[4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
        r[i] = Mx[i * 4 + 0] * t[0] + Mx[i * 4 + 1] * t[1] + Mx[i * 4 + 2] * t[2] + Mx[i * 4 + 3] * t[3];
}

// One CTA of kScanWarps warps per union section: turns zero-state end states into carry-in states, in
// place.  The cells are cut into one contiguous slice per warp; each warp scans its slice from zero state
// (32 cells per step: warp scan with the transition powers Mc^1..Mc^32), the slice carries are chained by
// one thread (c_{w+1} = Mc^{cells of slice w} c_w + e_w), and a second sweep adds Mc^{cells since the slice
// start} c_w to every cell.
constexpr int kScanWarps = 8;

__global__ void __launch_bounds__(kScanWarps * 32)
k_kw_scan(KwConfig K, SectionGeom G, int nsec, int cells_stride, double* __restrict__ state)
{
    __shared__ double s_end[kScanWarps][4];      // slice end state from zero carry-in, then the slice carry-in
    const int sec = blockIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long start;
    int n;
    section_bounds(G, sec, start, n);
    const int ncells = (n + K.cell - 1) / K.cell;
    const int ngroups = (ncells + 31) / 32;
    const int gper = (ngroups + kScanWarps - 1) / kScanWarps;
    const int g_lo = w * gper, g_hi = min(ngroups, g_lo + gper);
    double* st = state + (long long)sec * cells_stride * 4;
    double prev[4] = {0, 0, 0, 0};                           // state at the start of this group of 32 cells
    for (int g = g_lo; g < g_hi; ++g) {
        const int i = g * 32 + lane;
        double s[4] = {0, 0, 0, 0};
        if (i < ncells) { s[0] = st[i * 4]; s[1] = st[i * 4 + 1]; s[2] = st[i * 4 + 2]; s[3] = st[i * 4 + 3]; }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double t[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) t[r] = __shfl_up_sync(0xffffffffu, s[r], 1 << k);
            if (lane >= (1 << k)) mat4_apply_add(K.mpow + ((1 << k) - 1) * 16, t, s);
        }
        mat4_apply_add(K.mpow + lane * 16, prev, s);          // + Mc^(lane+1) * prev
        double c[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            c[r] = __shfl_up_sync(0xffffffffu, s[r], 1);
            if (lane == 0) c[r] = prev[r];
        }
        if (i < ncells) { st[i * 4] = c[0]; st[i * 4 + 1] = c[1]; st[i * 4 + 2] = c[2]; st[i * 4 + 3] = c[3]; }
#pragma unroll
        for (int r = 0; r < 4; ++r) prev[r] = __shfl_sync(0xffffffffu, s[r], 31);
    }
    if (lane == 0) { s_end[w][0] = prev[0]; s_end[w][1] = prev[1]; s_end[w][2] = prev[2]; s_end[w][3] = prev[3]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        // T = Mc^(32 * gper) by repeated multiplication with Mc^32 (a full slice; only the last one can be shorter)
        double T[16], c[4] = {0, 0, 0, 0};
        const double* M32 = K.mpow + 31 * 16;
        for (int e = 0; e < 16; ++e) T[e] = (e % 5 == 0) ? 1.0 : 0.0;
        for (int p = 0; p < gper; ++p) {
            double U[16];
            for (int r = 0; r < 4; ++r)
                for (int cc = 0; cc < 4; ++cc) {
                    double acc = 0;
                    for (int q = 0; q < 4; ++q) acc += M32[r * 4 + q] * T[q * 4 + cc];
                    U[r * 4 + cc] = acc;
                }
            for (int e = 0; e < 16; ++e) T[e] = U[e];
        }
        for (int ww = 0; ww < kScanWarps; ++ww) {
            const double e[4] = {s_end[ww][0], s_end[ww][1], s_end[ww][2], s_end[ww][3]};
            s_end[ww][0] = c[0]; s_end[ww][1] = c[1]; s_end[ww][2] = c[2]; s_end[ww][3] = c[3];
            double nx[4];
            mat4_vec(T, c, nx);
            for (int r = 0; r < 4; ++r) c[r] = nx[r] + e[r];
        }
    }
    __syncthreads();
    if (w == 0) return;                                       // the first slice starts from the true zero state
    double d[4] = {s_end[w][0], s_end[w][1], s_end[w][2], s_end[w][3]};       // Mc^(32 gi) c_w
    for (int g = g_lo; g < g_hi; ++g) {
        const int i = g * 32 + lane;
        double add[4] = {d[0], d[1], d[2], d[3]};
        if (lane > 0) mat4_vec(K.mpow + (lane - 1) * 16, d, add);             // Mc^lane d
        if (i < ncells) { st[i * 4] += add[0]; st[i * 4 + 1] += add[1]; st[i * 4 + 2] += add[2]; st[i * 4 + 3] += add[3]; }
        double nd[4];
        mat4_vec(K.mpow + 31 * 16, d, nd);
#pragma unroll
        for (int r = 0; r < 4; ++r) d[r] = nd[r];
    }
}

// One warp per (chunk, group) whose section starts later than the union's: re-filter the first
// patch_cells cells of the group's section from zero state, with the same cell-parallel scheme as
// the main passes (zero-state run per lane, warp scan of the carries, second run for the energies).
// patch[(ci*G + g) * patch_cells + k].
__global__ void k_kw_patch(KwConfig K, SectionGeom Gu, const SectionGeom* __restrict__ geoms, int nsec, int G,
                           double* __restrict__ patch)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nsec * G) return;
    const int ci = t / G, g = t % G;
    long long su, sg;
    int nu, ng;
    section_bounds(Gu, ci, su, nu);
    section_bounds(geoms[g], ci, sg, ng);
    if (sg <= su || ng <= 0) return;                          // same start as the union: nothing to patch
    const float* __restrict__ x = geoms[g].audio + (sg - geoms[g].base);
    const int cell = K.cell;
    double prev[4] = {0, 0, 0, 0};                            // state at the start of this group of 32 cells
    for (int g0 = 0; g0 < K.patch_cells; g0 += 32) {
        const int i = g0 + lane;
        const int lo = (int)min((long long)i * cell, (long long)ng);
        const int hi = i < K.patch_cells ? min(lo + cell, ng) : lo;
        double s[4] = {0, 0, 0, 0}, v;
        for (int p = lo; p < hi; ++p) kw_step(K.cf, (double)x[p], s[0], s[1], s[2], s[3], v);
        // inclusive scan: s_i = Mc s_{i-1} + z_i  (state at the end of cell i)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double u[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) u[r] = __shfl_up_sync(0xffffffffu, s[r], 1 << k);
            if (lane >= (1 << k)) mat4_apply_add(K.mpow + ((1 << k) - 1) * 16, u, s);
        }
        mat4_apply_add(K.mpow + lane * 16, prev, s);          // + Mc^(lane+1) * prev
        double c[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            c[r] = __shfl_up_sync(0xffffffffu, s[r], 1);
            if (lane == 0) c[r] = prev[r];
        }
        double e = 0;
        for (int p = lo; p < hi; ++p) {
            kw_step(K.cf, (double)x[p], c[0], c[1], c[2], c[3], v);
            e += v * v;
        }
        if (i < K.patch_cells) patch[(long long)t * K.patch_cells + i] = e;
#pragma unroll
        for (int r = 0; r < 4; ++r) prev[r] = __shfl_sync(0xffffffffu, s[r], 31);
    }
}

// One warp per (chunk, group): gating + gain.  lib.rs:142-214.
__global__ void k_kw_gate(KwConfig K, SectionGeom Gu, const SectionGeom* __restrict__ geoms, int nsec, int G,
                          int cells_stride, const double* __restrict__ energy, const double* __restrict__ energy_m1,
                          const double* __restrict__ patch, double* __restrict__ lufs_out,
                          double* __restrict__ gain_out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nsec * G) return;
    const int ci = warp / G, g = warp % G;
    long long su, start;
    int nu, n;
    section_bounds(Gu, ci, su, nu);
    section_bounds(geoms[g], ci, start, n);
    const int cell = K.cell;
    const int i0 = (int)((start - su) / cell);                // first union cell of this group's section
    const int npatch = start > su ? K.patch_cells : 0;
    const double* Eu = energy + (long long)ci * cells_stride + i0;
    const double* Ep = patch + (long long)warp * K.patch_cells;
    auto E = [&](int c) { return c < npatch ? Ep[c] : Eu[c]; };
    const int ncells = (n + cell - 1) / cell;
    const double rate = (double)K.rate;
    const double NEG_INF = -INFINITY;
    double lufs = NEG_INF;
    if (n > 0) {
        const double T = __ddiv_rn((double)n, rate);
        if (T < 0.5) {
            // block_size = T (apd.py:417): a single block [0, trunc(T*rate)) clipped to n
            const double win = __dmul_rn(T, rate);
            long long u = (long long)win;
            if (u > n) u = n;
            double tot = 0;
            for (int i = lane; i < ncells; i += 32) tot += E(i);
            tot = warp_sum(tot);
            if (u == (long long)n - 1) {
                // drop the last sample: replace the last cell's energy by the one without it
                tot = 0;
                for (int i = lane; i < ncells - 1; i += 32) tot += E(i);
                tot = warp_sum(tot) + energy_m1[ci];
            }
            // lib.rs:149-157 (num_blocks = round(0) + 1 = 1) then the two gates on one block
            if (u > 0) {
                const double ms = tot / (double)u;
                if (ms > 0.0) {
                    const double l = -0.691 + 10.0 * log10(ms);
                    // abs gate, then relative gate = l - 10 < l: the block survives iff l >= -70
                    if (l >= -70.0) lufs = -0.691 + 10.0 * log10(ms);
                }
            }
        } else {
            const double tg = 0.4;
            const long long nb = llround(__ddiv_rn(__dsub_rn(T, tg), __dmul_rn(tg, 0.25))) + 1;   // lib.rs:149
            const int hop = cell * K.k_per_hop, win = 4 * hop;
            if (nb <= 0) {
                double tot = 0;
                for (int i = lane; i < ncells; i += 32) tot += E(i);
                tot = warp_sum(tot);
                const double ms = tot / (double)n;
                lufs = ms <= 0.0 ? NEG_INF : -0.691 + 10.0 * log10(ms);
            } else {
                double gate = NEG_INF;
                for (int pass = 0; pass < 2; ++pass) {
                    double sum = 0;
                    double cnt = 0;
                    for (long long j = lane; j < nb; j += 32) {
                        const long long l = j * hop;
                        long long u = l + win;
                        if (u > n) u = n;
                        if (l >= u) continue;
                        const int c0 = (int)(j * K.k_per_hop);
                        const int c1 = min(c0 + 4 * K.k_per_hop, ncells);
                        double e = 0;
                        for (int c = c0; c < c1; ++c) e += E(c);
                        const double ms = e / (double)(u - l);
                        if (!(ms > 0.0)) continue;
                        const double ld = -0.691 + 10.0 * log10(ms);
                        const bool keep = pass == 0 ? (ld >= -70.0) : (ld > gate && ld >= -70.0);
                        if (keep) { sum += ms; cnt += 1.0; }
                    }
                    sum = warp_sum(sum);
                    cnt = warp_sum(cnt);
                    if (cnt == 0.0) { lufs = NEG_INF; break; }
                    const double mean = sum / cnt;
                    if (pass == 0) gate = -0.691 + 10.0 * log10(mean) - 10.0;
                    else lufs = -0.691 + 10.0 * log10(mean);
                }
            }
        }
    }
    if (lane == 0) {
        lufs_out[warp] = lufs;
        gain_out[warp] = pow(10.0, (kTargetLufs - lufs) / 20.0);   // lib.rs:221-222
    }
}

// General-rate path (0.1 * sample_rate is not a whole number of samples, e.g. 11 025 Hz): the block bounds of
// lib.rs:113-122 then fall on no cell grid.  One WARP per (chunk, group) section: the section is cut into 32 slices, lane
// w filters its slice with the serial float64 recurrence after a warm-up of K.warm samples from zero state (the
// K-weighting filters have forgotten their start state by then: n^2 r^n < 1e-20, the argument of k_kw_patch; a slice
// that starts within K.warm samples of the section start simply starts there, exactly), accumulates the squared
// prefix LOCAL to the slice and samples it at the truncated block bounds it passes.  A warp scan of the slice totals
// turns the local prefixes into the reference's P[hi] - P[lo] (same sums, re-associated), then the two gates run
// with the blocks spread over the lanes.  work[(ci * G + g) * stride ...]: [0, stride/2) prefix at lower bounds, then
// mean squares; [stride/2, stride) prefix at upper bounds.
__global__ void __launch_bounds__(128)
k_kw_serial(KwConfig K, const SectionGeom* __restrict__ geoms, int nsec, int G, long long stride,
            double* __restrict__ work_all, double* __restrict__ lufs_out, double* __restrict__ gain_out)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nsec * G) return;
    const int ci = t / G, g = t % G;
    long long start;
    int n;
    section_bounds(geoms[g], ci, start, n);
    const double NEG_INF = -INFINITY;
    double lufs = NEG_INF;
    if (n > 0) {                                             // warp-uniform
        const float* __restrict__ x = geoms[g].audio + (start - geoms[g].base);
        double* __restrict__ plo = work_all + (long long)t * stride;
        double* __restrict__ phi = plo + stride / 2;
        const double rate = (double)K.rate;
        const double T = __ddiv_rn((double)n, rate);
        const double block = T < 0.5 ? T : 0.4;                                   // apd.py:417
        const double win = __dmul_rn(block, rate), hop = __dmul_rn(win, 0.25);    // lib.rs:142-146
        long long nb = llround(__ddiv_rn(__dsub_rn(T, block), __dmul_rn(block, 0.25))) + 1;   // lib.rs:149
        if (nb > stride / 2) nb = stride / 2;                                     // (sized for the longest section)
        auto lo_of = [&](long long j) { return (long long)__dmul_rn((double)j, hop); };
        auto hi_of = [&](long long j) {
            const long long h = (long long)__dadd_rn(__dmul_rn((double)j, hop), win);
            return h > n ? (long long)n : h;
        };
        const int len = (n + 31) / 32;                       // slice length
        const int last_lane = (n - 1) / len;                 // last non-empty slice
        auto lane_of = [&](long long pos) { return pos >= n ? last_lane : (int)(pos / len); };
        const int a = min(n, lane * len), b = min(n, a + len);
        // this lane records the bounds at positions [a, b), and position n as well if it owns the section's end
        const long long own_hi = (lane == last_lane) ? (long long)n + 1 : (long long)b;
        double P = 0;
        if (a < b) {
            long long jl = 0, jh = 0;
            while (jl < nb && lo_of(jl) < a) ++jl;
            while (jh < nb && hi_of(jh) < a) ++jh;
            long long next_lo = jl < nb ? lo_of(jl) : -1, next_hi = jh < nb ? hi_of(jh) : -1;
            auto next_event = [&]() {
                long long e = own_hi;
                if (jl < nb && next_lo < e) e = next_lo;
                if (jh < nb && next_hi < e) e = next_hi;
                return e;
            };
            long long evt = next_event();
            double s1 = 0, s2 = 0, h1 = 0, h2 = 0, v;
            const int w0 = max(0, a - K.warm);
            for (int p0 = w0; p0 <= b; p0 += 32) {
                float xs[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) xs[k] = p0 + k < b ? x[p0 + k] : 0.0f;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const int p = p0 + k;
                    if (p > b) break;
                    if (p == evt && p < own_hi) {            // (position b belongs to the next slice)
                        while (jl < nb && next_lo == p) {
                            plo[jl] = P;
                            ++jl;
                            next_lo = jl < nb ? lo_of(jl) : -1;
                        }
                        while (jh < nb && next_hi == p) {
                            phi[jh] = P;
                            ++jh;
                            next_hi = jh < nb ? hi_of(jh) : -1;
                        }
                        evt = next_event();
                    }
                    if (p < b) {
                        kw_step(K.cf, (double)xs[k], s1, s2, h1, h2, v);
                        if (p >= a) P += v * v;
                    }
                }
            }
        }
        // exclusive scan of the slice totals
        double incl = P;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const double off = incl - P, total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        if (nb <= 0) {
            const double m = total / (double)n;
            lufs = m <= 0.0 ? NEG_INF : -0.691 + 10.0 * log10(m);
        } else {
            for (long long j0 = 0; j0 < nb; j0 += 32) {
                const long long j = j0 + lane;
                const bool in = j < nb;
                const long long lo = in ? lo_of(j) : 0, hi = in ? hi_of(j) : 0;
                const double olo = __shfl_sync(0xffffffffu, off, lane_of(lo));
                const double ohi = __shfl_sync(0xffffffffu, off, lane_of(hi));
                if (in) plo[j] = lo < hi ? ((phi[j] + ohi) - (plo[j] + olo)) / (double)(hi - lo) : -1.0;   // lib.rs:119-120
            }
            __syncwarp();
            double gate = NEG_INF;
            for (int pass = 0; pass < 2; ++pass) {
                double sum = 0, cnt = 0;
                for (long long j = lane; j < nb; j += 32) {
                    const double m = plo[j];
                    if (!(m > 0.0)) continue;
                    const double l = -0.691 + 10.0 * log10(m);
                    const bool keep = pass == 0 ? (l >= -70.0) : (l > gate && l >= -70.0);
                    if (keep) { sum += m; cnt += 1.0; }
                }
                sum = warp_sum(sum);
                cnt = warp_sum(cnt);
                if (cnt == 0.0) { lufs = NEG_INF; break; }
                const double mean = sum / cnt;
                if (pass == 0) gate = -0.691 + 10.0 * log10(mean) - 10.0;
                else lufs = -0.691 + 10.0 * log10(mean);
            }
        }
    }
    if (lane == 0) {
        lufs_out[t] = lufs;
        gain_out[t] = pow(10.0, (kTargetLufs - lufs) / 20.0);   // lib.rs:221-222
    }
}

// ------------------------------------------------------------------ host side
static void kw_coefficients(double rate, double cf[12])          // lib.rs:13-53
{
    const double A = pow(10.0, 4.0 / 40.0);
    const double w = 2.0 * M_PI * (1500.0 / rate);
    const double alpha = sin(w) / (2.0 * M_SQRT1_2);
    const double c = cos(w);
    const double k = 2.0 * sqrt(A) * alpha;
    const double a0 = (A + 1.0) - (A - 1.0) * c + k;
    cf[0] = A * ((A + 1.0) + (A - 1.0) * c + k) / a0;
    cf[1] = -2.0 * A * ((A - 1.0) + (A + 1.0) * c) / a0;
    cf[2] = A * ((A + 1.0) + (A - 1.0) * c - k) / a0;
    cf[3] = 1.0;
    cf[4] = 2.0 * ((A - 1.0) - (A + 1.0) * c) / a0;
    cf[5] = ((A + 1.0) - (A - 1.0) * c - k) / a0;
    const double w2 = 2.0 * M_PI * (38.0 / rate);
    const double alpha2 = sin(w2) / (2.0 * 0.5);
    const double c2 = cos(w2);
    const double h0 = 1.0 + alpha2;
    cf[6] = ((1.0 + c2) / 2.0) / h0;
    cf[7] = (-(1.0 + c2)) / h0;
    cf[8] = ((1.0 + c2) / 2.0) / h0;
    cf[9] = 1.0;
    cf[10] = (-2.0 * c2) / h0;
    cf[11] = (1.0 - alpha2) / h0;
}

bool kw_config_create(int sample_rate, KwConfig* out, std::string* err)
{
    if (sample_rate <= 0) {
        if (err) *err = "sample rate must be positive";
        return false;
    }
    KwConfig K;
    memset(&K, 0, sizeof(K));
    K.rate = sample_rate;
    kw_coefficients((double)sample_rate, K.cf);
    if (sample_rate % 10 != 0) {
        // the 100 ms gating hop is not a whole number of samples: serial general-rate path (k_kw_serial); `cell` only
        // sizes the workspace then (at least one slot per gating block)
        K.general = 1;
        K.cell = std::max(1, sample_rate / 10);
        K.k_per_hop = 1;
        K.patch_cells = 1;
        {
            // memory of the slowest K-weighting pole (see below): the warm-up of a slice in k_kw_serial
            const double r = sqrt(fabs(K.cf[11]));
            int n = 1;
            while ((double)(n + 1) * (double)(n + 1) * pow(r, (double)n) > 1e-20 && n < 10000000) ++n;
            K.warm = n + 1;
        }
        *out = K;
        return true;
    }
    const int hop = sample_rate / 10;
    int k = 1;
    while (k <= hop && !(hop % k == 0 && hop / k <= 128)) ++k;
    K.k_per_hop = k;
    K.cell = hop / k;
    {
        // memory of the slowest K-weighting pole (the 38 Hz high-pass, a double pole of radius sqrt(a2))
        const double r = sqrt(fabs(K.cf[11]));
        int n = 1;
        while ((double)(n + 1) * (double)(n + 1) * pow(r, (double)n) > 1e-20 && n < 10000000) ++n;
        K.patch_cells = (n + K.cell - 1) / K.cell + 1;
    }
    // Mc = A^cell by running the zero-input response of the 4 unit states for `cell` samples
    double Mc[16];
    for (int col = 0; col < 4; ++col) {
        double s[4] = {0, 0, 0, 0};
        s[col] = 1.0;
        for (int t = 0; t < K.cell; ++t) {
            const double u = s[0];
            const double n1 = -K.cf[4] * u + s[1];
            const double n2 = -K.cf[5] * u;
            const double v = K.cf[6] * u + s[2];
            const double n3 = K.cf[7] * u - K.cf[10] * v + s[3];
            const double n4 = K.cf[8] * u - K.cf[11] * v;
            s[0] = n1; s[1] = n2; s[2] = n3; s[3] = n4;
        }
        for (int r = 0; r < 4; ++r) Mc[r * 4 + col] = s[r];
    }
    std::vector<double> pw(32 * 16);
    memcpy(&pw[0], Mc, sizeof(Mc));
    for (int p = 1; p < 32; ++p)
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double acc = 0;
                for (int q = 0; q < 4; ++q) acc += pw[(p - 1) * 16 + r * 4 + q] * Mc[q * 4 + c];
                pw[p * 16 + r * 4 + c] = acc;
            }
    // g[j] = state j samples after a unit impulse into the zero state; stored reversed: imp[t] = g[cell - 1 - t]
    const size_t imp_off = pw.size();
    pw.resize(imp_off + (size_t)K.cell * 4);
    {
        const double* cf = K.cf;
        double s[4];
        const double u0 = cf[0], v0 = cf[6] * u0;
        s[0] = cf[1] - cf[4] * u0;
        s[1] = cf[2] - cf[5] * u0;
        s[2] = cf[7] * u0 - cf[10] * v0;
        s[3] = cf[8] * u0 - cf[11] * v0;
        for (int j = 0; j < K.cell; ++j) {
            for (int r = 0; r < 4; ++r) pw[imp_off + (size_t)(K.cell - 1 - j) * 4 + r] = s[r];
            const double u = s[0];
            const double n1 = -cf[4] * u + s[1];
            const double n2 = -cf[5] * u;
            const double v = cf[6] * u + s[2];
            const double n3 = cf[7] * u - cf[10] * v + s[3];
            const double n4 = cf[8] * u - cf[11] * v;
            s[0] = n1; s[1] = n2; s[2] = n3; s[3] = n4;
        }
    }
    double* d = nullptr;
    if (cudaMalloc(&d, pw.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(d, pw.data(), pw.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        if (err) *err = "cudaMalloc failed (loudness tables)";
        return false;
    }
    K.mpow = d;
    K.imp = d + imp_off;
    *out = K;
    return true;
}

void kw_config_destroy(KwConfig* K)
{
    if (K->mpow) cudaFree((void*)K->mpow);
    K->mpow = nullptr;
    K->imp = nullptr;
}

void launch_loudness(const KwConfig& K, const SectionGeom& Gu, const SectionGeom* h_geoms,
                     const SectionGeom* d_geoms, int G, int nsec, int cells_stride, double* state, double* energy,
                     double* energy_m1, double* patch, double* lufs, double* gain, cudaStream_t st)
{
    if (nsec <= 0) return;
    if (K.general) {
        // energy[] is the per-(chunk, group) block workspace: cells_stride / G slots each
        k_kw_serial<<<(nsec * G + 3) / 4, 128, 0, st>>>(K, d_geoms, nsec, G, (long long)(cells_stride / G), energy, lufs, gain);
        return;
    }
    static bool attr = false;
    const size_t smem = (size_t)128 * (K.cell + 1) * sizeof(float);
    if (!attr) {
        cudaFuncSetAttribute(k_kw_cells<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(k_kw_cells<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr = true;
    }
    // the longest section of the launch decides the grid; CTAs past a section's end exit at once
    long long start;
    int nmax = 0;
    for (int s = 0; s < nsec; ++s) {
        int n;
        section_bounds(Gu, s, start, n);
        nmax = n > nmax ? n : nmax;
    }
    if (nmax <= 0) nmax = 1;
    bool need_patch = false;
    for (int g = 0; g < G; ++g) need_patch |= h_geoms[g].halo != Gu.halo;
    const int ncells = (nmax + K.cell - 1) / K.cell;
    dim3 grid((ncells + 127) / 128, nsec);
    k_kw_cells<0><<<grid, 128, smem, st>>>(K, Gu, cells_stride, state, energy, energy_m1);
    k_kw_scan<<<nsec, kScanWarps * 32, 0, st>>>(K, Gu, nsec, cells_stride, state);
    k_kw_cells<1><<<grid, 128, smem, st>>>(K, Gu, cells_stride, state, energy, energy_m1);
    if (need_patch) k_kw_patch<<<(nsec * G * 32 + 127) / 128, 128, 0, st>>>(K, Gu, d_geoms, nsec, G, patch);
    k_kw_gate<<<(nsec * G * 32 + 127) / 128, 128, 0, st>>>(K, Gu, d_geoms, nsec, G, cells_stride, energy, energy_m1,
                                                          patch, lufs, gain);
}

}  // namespace apd
