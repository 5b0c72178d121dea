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
    int patch_cells;      // cells after which the filters have forgotten their start state (see k_kw_patch)
    int general;          // 1: 0.1 * rate is not a whole number of samples -> serial path (k_kw_serial)
    int warm;             // general path: samples after which the filters have forgotten their start state
    const double* mpow;   // device: Mc^1 .. Mc^32 (4x4 row-major each), Mc = state transition over one cell
    const double* imp;    // device: [cell][4] impulse-response states, reversed (same allocation as mpow)
};

bool kw_config_create(int sample_rate, KwConfig* out, std::string* err);
void kw_config_destroy(KwConfig* K);

// Loudness of every (chunk ci < nsec, group g < G) section.  Gu is the union geometry (largest halo);
// h_geoms / d_geoms are the per-group geometries on host / device.  Workspace: state[nsec*cells_stride*4],
// energy[nsec*cells_stride], energy_m1[nsec], patch[nsec*G*patch_cells].  Results: lufs/gain[ci*G + g].
void launch_loudness(const KwConfig& K, const SectionGeom& Gu, const SectionGeom* h_geoms,
                     const SectionGeom* d_geoms, int G, int nsec, int cells_stride, double* state, double* energy,
                     double* energy_m1, double* patch, double* lufs, double* gain, cudaStream_t st);

}  // namespace apd
