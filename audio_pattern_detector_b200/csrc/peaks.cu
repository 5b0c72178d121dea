// Candidate selection: which units can have peaks at all, then scipy-compatible
// find_peaks(height, distance) on the normalised correlation of those units.
//
// Replaces _native.find_peaks (reference lib.rs:380-396, 404-485) as called at
// audio_pattern_detector.py:516-522.
#include "internal.h"
#include "peaks.h"

namespace apd {

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

// block-wide exclusive scan of one int per thread (1024 threads); returns the thread's offset, total in *tot
__device__ __forceinline__ int block_exscan(int v, int* warp_tot, int* tot)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    int off = 0, all = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
        const int t = warp_tot[k];
        if (k < w) off += t;
        all += t;
    }
    *tot = all;
    return off + inc - v;
}

constexpr int kPeakItems = 8;     // consecutive samples per thread per tile

__global__ void __launch_bounds__(1024)
k_find_peaks(PeakArgs A)
{
    __shared__ int warp_tot[32];
    __shared__ unsigned long long s_best[32];
    __shared__ int s_win;
    const int slot = blockIdx.x;
    if (A.slot0 + slot >= *A.sel_count) return;
    const int2 unit = A.sel[A.slot0 + slot];
    long long start;
    int nsec;
    section_bounds(A.geom[A.clip_group[unit.y]], unit.x, start, nsec);
    const int L = A.clip_len[unit.y];
    const int n = nsec > 0 ? nsec + L - 1 : 0;
    const float* __restrict__ q = A.corr + (long long)slot * A.corr_stride;
    int* cidx = A.cand_idx + (long long)slot * A.cand_stride;
    float* cval = A.cand_val + (long long)slot * A.cand_stride;
    unsigned char* state = A.cand_state + (long long)slot * A.cand_stride;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;

    // ---- 1. local maxima >= height, ordered (each thread owns kPeakItems consecutive samples of a tile).
    //         The ten samples a thread needs are loaded up front (two 16-byte loads + the two neighbours), and a
    //         tile without any candidate - nearly all of them - costs one barrier and no scan.
    int base = 0;
    const bool vec_ok = (reinterpret_cast<unsigned long long>(q) & 15ull) == 0ull;
    for (int t0 = 0; t0 < n - 1; t0 += blockDim.x * kPeakItems) {
        const int i0 = t0 + threadIdx.x * kPeakItems;
        int pidx[kPeakItems];
        float pval[kPeakItems];
        int cnt = 0;
        if (i0 < n - 1) {
            float w[kPeakItems + 2];                         // w[k + 1] = q[i0 + k], k = -1 .. kPeakItems
            if (vec_ok && i0 + kPeakItems < n) {
                const float4 a = *reinterpret_cast<const float4*>(q + i0);
                const float4 b = *reinterpret_cast<const float4*>(q + i0 + 4);
                w[0] = i0 > 0 ? q[i0 - 1] : 0.0f;
                w[kPeakItems + 1] = q[i0 + kPeakItems];
                w[1] = a.x; w[2] = a.y; w[3] = a.z; w[4] = a.w;
                w[5] = b.x; w[6] = b.y; w[7] = b.z; w[8] = b.w;
            } else {
#pragma unroll
                for (int k = 0; k < kPeakItems + 2; ++k) {
                    const int idx = i0 + k - 1;
                    w[k] = (idx >= 0 && idx < n) ? q[idx] : 0.0f;
                }
            }
            float hi = w[1];
#pragma unroll
            for (int k = 2; k <= kPeakItems; ++k) hi = fmaxf(hi, w[k]);
            if (hi >= A.height) {
#pragma unroll
                for (int k = 0; k < kPeakItems; ++k) {
                    const int i = i0 + k;
                    const float v = w[k + 1];
                    if (i >= 1 && i < n - 1 && w[k] < v && v >= A.height) {
                        if (w[k + 2] != v) {
                            if (v > w[k + 2]) { pidx[cnt] = i; pval[cnt] = v; ++cnt; }
                        } else {
                            int e = i + 1;                   // plateau: walk to its end in memory
                            while (e + 1 < n && q[e + 1] == v) ++e;
                            if (e + 1 < n && v > q[e + 1]) { pidx[cnt] = (i + e) >> 1; pval[cnt] = v; ++cnt; }
                        }
                    }
                }
            }
        }
        if (!__syncthreads_or(cnt)) continue;
        int tot;
        const int off = base + block_exscan(cnt, warp_tot, &tot);
        for (int k = 0; k < cnt; ++k) {
            cidx[off + k] = pidx[k];
            cval[off + k] = pval[k];
            state[off + k] = 0;
        }
        base += tot;
    }
    const int nc = base;
    __syncthreads();

    // ---- 2. greedy tallest-first suppression, one kept peak per round:
    //         the best-ranked undecided candidate is what the sequential pass keeps next.
    //         rank key: value bits (non-negative floats order as unsigned) then lower index first.
    const int dist = L;                                  // apd.py:516
    for (int round = 0; round <= nc; ++round) {
        unsigned long long best = 0ull;
        for (int c = threadIdx.x; c < nc; c += blockDim.x)
            if (state[c] == 0) {
                const unsigned long long key =
                    ((unsigned long long)__float_as_uint(cval[c]) << 32) | (unsigned)(0x7fffffff - c);
                best = key > best ? key : best;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t > best ? t : best;
        }
        if (lane == 0) s_best[w] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long b = 0ull;
            for (int k = 0; k < nw; ++k) b = s_best[k] > b ? s_best[k] : b;
            s_win = b ? 0x7fffffff - (int)(unsigned)(b & 0xffffffffull) : -1;
            if (b) state[s_win] = 1;
        }
        __syncthreads();
        const int win = s_win;
        if (win < 0) break;
        const int iw = cidx[win];
        // candidates are sorted by position: walk outwards from the winner in strides
        for (int c = win - 1 - (int)threadIdx.x; c >= 0; c -= blockDim.x) {
            if (iw - cidx[c] >= dist) break;
            if (state[c] == 0) state[c] = 2;
        }
        for (int c = win + 1 + threadIdx.x; c < nc; c += blockDim.x) {
            if (cidx[c] - iw >= dist) break;
            if (state[c] == 0) state[c] = 2;
        }
        __syncthreads();
    }

    // ---- 3. ordered compaction of the survivors (+ their heights)
    int* __restrict__ peaks = A.peaks + (long long)slot * A.peak_stride;
    float* __restrict__ heights = A.peak_height + (long long)slot * A.peak_stride;
    base = 0;
    for (int c0 = 0; c0 < nc; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const int keep = (c < nc && state[c] == 1) ? 1 : 0;
        int tot;
        const int pos = base + block_exscan(keep, warp_tot, &tot);
        if (keep && pos < A.peak_stride) {
            peaks[pos] = cidx[c];
            heights[pos] = cval[c];
        }
        base += tot;
    }
    if (threadIdx.x == 0) {
        A.n_peaks[slot] = base < A.peak_stride ? base : A.peak_stride;
        A.n_cands[slot] = nc;
        if (base > A.peak_stride) atomicOr(A.overflow, 1);
    }
}

void launch_find_peaks(const PeakArgs& A, int nslots, cudaStream_t st)
{
    if (nslots <= 0) return;
    k_find_peaks<<<nslots, 1024, 0, st>>>(A);
}

}  // namespace apd
