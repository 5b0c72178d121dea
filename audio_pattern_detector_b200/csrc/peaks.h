// Candidate selection + find_peaks (see peaks.cu) and Step-2 verifiers (see verify.cu).
#pragma once
#include <cuda_runtime.h>

#include "internal.h"
#include "../../include/apd_b200.h"

namespace apd {

struct PeakArgs {
    const int2* sel;              // selected (ci, clip) units, ordered
    const int* sel_count;
    int slot0;                    // slot s of this launch handles sel[slot0 + s]
    const SectionGeom* geom;      // per group (device array)
    const int* clip_group;        // per clip
    const int* clip_len;          // per clip
    float height;
    const float* corr;            // normalised correlation per slot
    long long corr_stride;
    int* cand_idx;                // per slot scratch, capacity cand_stride
    float* cand_val;
    unsigned char* cand_state;
    long long cand_stride;
    int* peaks;                   // per slot result, capacity peak_stride
    float* peak_height;           // normalised correlation at each surviving peak
    int peak_stride;
    int* n_peaks;                 // per slot
    int* n_cands;                 // per slot (local maxima >= height before the distance filter)
    int* overflow;
};

void launch_find_peaks(const PeakArgs& A, int nslots, cudaStream_t st);

// Per-clip verification data (device arrays indexed by clip).
struct ClipVerify {
    const int* clip_len;
    const int* clip_group;
    const int* strategy;              // APD_STRATEGY_*
    const float* const* self_corr;    // normalised |self correlation|, 2L-1
    const int* win_lo;                // [clip][3] Pearson window bounds into the 2L-1 curve
    const int* win_hi;
    const int* win_ds;                // [clip][3] down-sampled length (0 = unused window)
    const float* const* win_cache;    // [clip] -> concatenated clip-side down-sampled windows
    const int* is_short;
    // marker tone
    const double* tone_hz;
    const double* tone_thr;           // [clip][6]
    const int* tone_P;                // Bluestein FFT length (power of two >= 2L-1)
    const double2* const* tone_chirp_fft;   // [clip] -> FFT_P of the chirp
    const double2* const* tone_tw;          // [clip] -> e^{-2 pi i t / P}, t < P/2
    const double2* const* tone_pre;         // [clip] -> hann[n] * e^{-i pi n^2 / L}, n < L
    const double2* const* tone_post;        // [clip] -> e^{-i pi k^2 / L}, k <= L/2
};

struct VerifyArgs {
    PeakArgs pk;                  // reuses sel / geometry / corr / peaks
    ClipVerify cv;
    int chunk0;
    int sample_rate;
    const double* gains;          // [ci * n_groups + group]
    int n_groups;
    apd_candidate* out;           // ordered output, filled by k_emit
    int* out_count;
    int out_capacity;
    apd_candidate* slot_cands;    // [slot][peak_stride] scratch records
    double2* tone_scratch;        // [round item][segment][2] Bluestein ping-pong buffers
    long long tone_scratch_stride;
    int tone_scratch_slots;
};

constexpr int kMaxSlots = 1024;           // slots per phase-2 round (k_emit keeps offsets in shared memory)

void launch_verify(const VerifyArgs& A, int nslots, cudaStream_t st, long long* launches);
void launch_tone_collect(const VerifyArgs& A, int nslots, void* items, int* n_items, int item_capacity,
                         cudaStream_t st, long long* launches);
// all_segments == false: the flank transforms of items whose matched segment already fails are skipped (their
// metrics read as zero, the decision is unchanged); alive: n_items bytes of device scratch
void launch_tone_batch(const VerifyArgs& A, void* items, int* n_items_dev, int n_items, double* metrics,
                       double* stats, int round_items, int max_P, int max_L, int wl, bool all_segments,
                       unsigned char* alive, cudaStream_t st, long long* launches);
void launch_emit(const VerifyArgs& A, int nslots, cudaStream_t st, long long* launches);
void launch_tone_tables(int L, int P, const double2* tw, double2* buf0, double2* buf1, double2* chirp_fft,
                        double2* pre, double2* post, cudaStream_t st);
size_t tone_item_bytes();

}  // namespace apd
