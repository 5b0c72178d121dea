// Internal structures shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "fft_sub.cuh"

namespace apd {

// ---------------------------------------------------------------------------
// Four-step FFT plan for a complex transform of M = N1 * N2 points, used for a
// real (negacyclic, odd-frequency) transform of N = 2M points.  See fft4.cu.
// ---------------------------------------------------------------------------
struct Fft4Plan {
    int M = 0, N1 = 0, N2 = 0;
    SubPlan col{};          // length N1 (column passes)
    SubPlan row{};          // length N2 (row passes)
    int tb_log2 = 0;        // column tile width  TB = 1 << tb_log2 (adjacent columns per CTA)
    int tr_log2 = 0;        // row tile height    TR = 1 << tr_log2 (adjacent rows per CTA)
    int threads = 256;
    size_t smem_col = 0, smem_row = 0;
};

// Where section `ci` (chunk chunk0+ci, look-back `halo`) lives inside the device-resident stream.
struct SectionGeom {
    const float* audio;     // audio[0] is stream sample `base`
    long long base;         // global index of audio[0]
    long long total;        // global index one past the last available sample
    long long chunk;        // samples per chunk C
    int chunk0;             // global index of the batch's first chunk
    int halo;               // look-back samples for this group (sliding_window * sr)
};

__host__ __device__ inline void section_bounds(const SectionGeom& g, int ci, long long& start_abs, int& n)
{
    const long long i = (long long)g.chunk0 + ci;
    start_abs = i * g.chunk - (i > 0 ? (long long)g.halo : 0LL);     // apd.py:406-412
    long long end_abs = (i + 1) * g.chunk;
    if (end_abs > g.total) end_abs = g.total;
    n = end_abs > start_abs ? (int)(end_abs - start_abs) : 0;
}

// How the CTAs of an inverse launch find their (section, clip) unit.  A launch serves every clip
// whose sliding-window group uses the same FFT shape ("shape class").
struct UnitSrc {
    const int2* list;        // explicit (ci, clip) pairs (phase 2), or nullptr for the dense layout below
    const int* list_begin;   // phase 2: this launch owns list positions [list_begin[0], list_begin[1])
    const int* shape_clips;  // dense: clip-major, clip = shape_clips[u / nb], ci = u % nb (consecutive units share
    int ns;                  //        a clip, so its spectrum rows stay in L1/L2 across the per-CTA unit loop)
    int u0;                  // first unit of this launch (offset into list / dense numbering)
    int nb;                  // dense: chunks in the batch
    int tc, tk;              // dense, tiled (tc > 0): units come in tiles of tc clips x tk chunks (clip-major inside a
                             // tile, chunk tiles fastest), so a group of tc*tk consecutive units shares tk section
                             // spectra and tc clip spectra; positions beyond the clip / chunk range are unused
};

// Returns false if launch-local unit u is not this launch's to process.
__device__ __forceinline__ bool get_unit(const UnitSrc& s, int u, int2* unit)
{
    u += s.u0;
    if (s.list) {
        if (u < s.list_begin[0] || u >= s.list_begin[1]) return false;
        *unit = s.list[u];
        return true;
    }
    if (s.tc > 0) {
        const int per_tile = s.tc * s.tk, ktiles = (s.nb + s.tk - 1) / s.tk;
        const int t = u / per_tile, i = u - t * per_tile;
        const int ct = t / ktiles, kt = t - ct * ktiles;
        const int ci = kt * s.tk + i % s.tk, k = ct * s.tc + i / s.tk;
        if (ci >= s.nb || k >= s.ns) return false;
        *unit = make_int2(ci, s.shape_clips[k]);
        return true;
    }
    *unit = s.nb > 0 ? make_int2(u % s.nb, s.shape_clips[u / s.nb]) : make_int2(u / s.ns, s.shape_clips[u % s.ns]);
    return true;
}

// Per-clip tables + per-group geometry needed to place a unit's data.
struct UnitCtx {
    const SectionGeom* geoms;        // [group] (device)
    const int* clip_group;           // [clip]
    const int* clip_len;             // [clip]
    const long long* clip_spec_off;  // [clip] offset of the clip's group spectrum inside a chunk's slab
    const float2* const* clip_spec;  // [clip] spectrum of the reversed, normalised clip
};

struct InvOut {
    // phase 1: per-unit maximum of |corr| (float bits, atomicMax), indexed ci * n_clips + clip
    unsigned int* unit_max_bits;
    int n_clips;
    // phase 2: normalised correlation written to slot (launch-local unit index), stride in floats
    float* corr;
    long long corr_stride;
    const float* self_max;   // per clip
};

bool build_plan(int M_min, Fft4Plan* plan, std::string* err);                 // picks N1, N2 >= M_min
void free_plan(Fft4Plan* plan);
long long plan_min_M_for(long long n_out);

// The sliding-window groups of one FFT shape, transformed by the same launches (device arrays of ng entries):
// transform t of a launch is section t / ng of group t % ng, so the groups of a chunk re-read its audio from L2.
struct FwdGroups {
    const int* halo;             // look-back samples of each group
    const int* gain_index;       // column of the gain matrix (global group id)
    const long long* spec_off;   // offset of the group's spectrum inside a chunk's slab
    int ng;                      // 0: a single group described by G.halo, gain column 0, offset 0
};

// scratch holds nsec * max(ng, 1) * M complex; spectra go to spec + section * spec_stride + spec_off[group].
void launch_forward(const Fft4Plan& P, const SectionGeom& G, const FwdGroups& FG, const double* gains, int gain_stride,
                    int nsec, float2* scratch, float2* spec, long long spec_stride, cudaStream_t st);
// Fused spectral multiply + inverse FFT + |.| ; write == false: per-unit max, write == true: normalised
// correlation into O.corr.  scratch holds nunits * M complex (launch-local unit index).
// desc: device scratch of corr_inv_desc_bytes(nunits) bytes (per-unit descriptors of the hot-shape kernels).
void launch_inverse(const Fft4Plan& P, const UnitCtx& C, const float2* spec, long long spec_slab,
                    const UnitSrc& U, int nunits, float2* scratch, void* desc, const InvOut& out, bool write,
                    cudaStream_t st);
// returns the number of kernels launched

// corr_inv.cu: the register-resident packed-arithmetic kernels for M = N1 x 512, N1 in {512, 576, 640}
bool corr_inv_supported(const Fft4Plan& P);
size_t corr_inv_desc_bytes(int nunits);
void launch_corr_inv(const Fft4Plan& P, const UnitCtx& C, const float2* spec, long long spec_slab, const UnitSrc& U,
                     int nunits, float2* scratch, void* desc, const InvOut& out, bool write, cudaStream_t st);

}  // namespace apd
