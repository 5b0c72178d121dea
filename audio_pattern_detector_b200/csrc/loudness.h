// K0 loudness configuration (see loudness.cu).
#pragma once
#include <cuda_runtime.h>
#include <string>

#include "internal.h"

namespace apd {

struct KwConfig {
    double cf[12];        // K-weighting biquad coefficients [b_shelf | a_shelf | b_hpass | a_hpass]
    int rate;
    int cell;             // samples per scan cell (divides the 100 ms hop)
    int k_per_hop;        // cells per hop
    const double* mpow;   // device: Mc^1 .. Mc^32 (4x4 row-major each), Mc = state transition over one cell
};

bool kw_config_create(int sample_rate, KwConfig* out, std::string* err);
void kw_config_destroy(KwConfig* K);

// Loudness of sections 0..nsec-1 of geometry G.  Workspace rows are indexed sec0 + s:
// state[(sec0+s) * cells_stride * 4], energy[(sec0+s) * cells_stride], energy_m1[sec0+s].
// Results: lufs[s * out_stride], gain[s * out_stride].
void launch_loudness(const KwConfig& K, const SectionGeom& G, int nsec, int sec0, int cells_stride,
                     double* state, double* energy, double* energy_m1, double* lufs, double* gain,
                     int out_stride, cudaStream_t st);

}  // namespace apd
