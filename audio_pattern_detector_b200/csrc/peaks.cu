// Candidate selection: which units can have peaks at all, then scipy-compatible
// find_peaks(height, distance) on the normalised correlation of those units.
//
// Replaces _native.find_peaks (reference lib.rs:380-396, 404-485) as called at
// audio_pattern_detector.py:516-522.
#include "internal.h"
#include "peaks.h"

namespace apd {

// ---------------------------------------------------------------------------
// k_select: a unit can only have a peak >= height if its global maximum, normalised,
// reaches height (every other sample is smaller).  Ordered compaction by one CTA so the
// selected list is sorted by (chunk, clip).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_select(const unsigned int* __restrict__ unit_max_bits, const float* __restrict__ self_max, int n_clips,
         int n_units, float height, int2* __restrict__ sel, int* __restrict__ sel_count, int capacity)
{
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int u0 = 0; u0 < n_units; u0 += blockDim.x) {
        const int u = u0 + threadIdx.x;
        bool pass = false;
        if (u < n_units) {
            const float am = __uint_as_float(unit_max_bits[u]);
            const float mc = fmaxf(self_max[u % n_clips], am);
            pass = mc > 0.0f && (am / mc) >= height;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; ++i) off += warp_tot[i];
        if (pass) {
            const int pos = off + __popc(bal & ((1u << lane) - 1));
            if (pos < capacity) sel[pos] = make_int2(u / n_clips, u % n_clips);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < (blockDim.x >> 5); ++i) t += warp_tot[i];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *sel_count = base;
}

// ---------------------------------------------------------------------------
// k_find_peaks: one CTA per selected unit (slot).
//   1. strict local maxima with plateau midpoints (lib.rs:404-428), >= height (lib.rs:431-433),
//      compacted in ascending order;
//   2. greedy tallest-first distance suppression (lib.rs:437-485) done in parallel rounds:
//      an undecided candidate that out-ranks every undecided candidate closer than `distance`
//      is exactly what the sequential greedy pass would keep next in its neighbourhood;
//      it then suppresses those neighbours.  Rank: taller first, ties -> lower index.
//   3. survivors compacted (ascending).
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool outranks(float va, int ia, float vb, int ib)
{
    return va > vb || (va == vb && ia < ib);
}

__global__ void __launch_bounds__(1024)
k_find_peaks(PeakArgs A)
{
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    __shared__ int s_any;
    const int slot = blockIdx.x;
    if (A.slot0 + slot >= *A.sel_count) return;
    const int2 unit = A.sel[A.slot0 + slot];
    long long start;
    int nsec;
    section_bounds(A.geom[A.clip_group[unit.y]], unit.x, start, nsec);
    const int L = A.clip_len[unit.y];
    const int n = nsec > 0 ? nsec + L - 1 : 0;
    const float* __restrict__ q = A.corr + (long long)slot * A.corr_stride;
    int* __restrict__ cidx = A.cand_idx + (long long)slot * A.cand_stride;
    float* __restrict__ cval = A.cand_val + (long long)slot * A.cand_stride;
    unsigned char* __restrict__ state = A.cand_state + (long long)slot * A.cand_stride;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;

    // ---- 1. local maxima >= height, ordered
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 1; i0 < n - 1; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        int pidx = -1;
        float pval = 0.0f;
        if (i < n - 1) {
            const float v = q[i];
            if (q[i - 1] < v) {
                int e = i;
                while (e + 1 < n && q[e + 1] == v) ++e;
                if (e + 1 < n && v > q[e + 1]) {
                    pidx = (i + e) >> 1;
                    pval = v;
                }
            }
        }
        const bool pass = pidx >= 0 && pval >= A.height;
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < w; ++k) off += warp_tot[k];
        if (pass) {
            const int pos = off + __popc(bal & ((1u << lane) - 1));
            cidx[pos] = pidx;
            cval[pos] = pval;
            state[pos] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < nw; ++k) t += warp_tot[k];
            s_base += t;
        }
        __syncthreads();
    }
    const int nc = s_base;
    __threadfence_block();
    __syncthreads();

    // ---- 2. distance suppression rounds (state: 0 undecided, 1 kept, 2 suppressed, 3 kept this round)
    const int dist = L;                                  // apd.py:516
    for (int round = 0; round < nc + 1; ++round) {
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += blockDim.x) {
            if (state[c] != 0) continue;
            const int ic = cidx[c];
            const float vc = cval[c];
            bool best = true;
            for (int j = c - 1; j >= 0 && ic - cidx[j] < dist && best; --j) {
                const unsigned char sj = state[j];
                if ((sj == 0 || sj == 3) && outranks(cval[j], j, vc, c)) best = false;
            }
            for (int j = c + 1; j < nc && cidx[j] - ic < dist && best; ++j) {
                const unsigned char sj = state[j];
                if ((sj == 0 || sj == 3) && outranks(cval[j], j, vc, c)) best = false;
            }
            if (best) state[c] = 3;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += blockDim.x) {
            if (state[c] != 0) continue;
            const int ic = cidx[c];
            bool sup = false;
            for (int j = c - 1; j >= 0 && ic - cidx[j] < dist && !sup; --j) sup = state[j] == 3;
            for (int j = c + 1; j < nc && cidx[j] - ic < dist && !sup; ++j) sup = state[j] == 3;
            if (sup) state[c] = 2;
            else s_any = 1;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += blockDim.x)
            if (state[c] == 3) state[c] = 1;
        const int any = s_any;
        __syncthreads();
        if (!any) break;
    }

    // ---- 3. ordered compaction of the survivors
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    int* __restrict__ peaks = A.peaks + (long long)slot * A.peak_stride;
    for (int c0 = 0; c0 < nc; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const bool keep = c < nc && state[c] == 1;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < w; ++k) off += warp_tot[k];
        if (keep) {
            const int pos = off + __popc(bal & ((1u << lane) - 1));
            if (pos < A.peak_stride) peaks[pos] = cidx[c];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < nw; ++k) t += warp_tot[k];
            s_base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        A.n_peaks[slot] = s_base < A.peak_stride ? s_base : A.peak_stride;
        A.n_cands[slot] = nc;
        if (s_base > A.peak_stride) atomicOr(A.overflow, 1);
    }
}

void launch_select(const unsigned int* unit_max_bits, const float* self_max, int n_clips, int n_units,
                   float height, int2* sel, int* sel_count, int capacity, cudaStream_t st)
{
    k_select<<<1, 1024, 0, st>>>(unit_max_bits, self_max, n_clips, n_units, height, sel, sel_count, capacity);
}

void launch_find_peaks(const PeakArgs& A, int nslots, cudaStream_t st)
{
    if (nslots <= 0) return;
    k_find_peaks<<<nslots, 1024, 0, st>>>(A);
}

}  // namespace apd
