// Shared device/host helpers for the B200 detection kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define APD_OK 0
#define APD_ERR_INVALID 1
#define APD_ERR_CUDA 2
#define APD_ERR_UNSUPPORTED 3
#define APD_ERR_OVERFLOW 4

namespace apd {

constexpr int kWarp = 32;

// Detection constants mirrored from the reference (audio_pattern_detector.py):
constexpr float kDefaultHeight = 0.25f;          // :520
constexpr double kShortClipSeconds = 0.5;        // :36
constexpr float kMseLimit = 0.02f;               // :793 (compared in float32, NumPy 2 semantics)
constexpr double kPearsonMin = 0.90;             // :794
constexpr double kTargetLufs = -16.0;            // :171, :420

__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)   // a * conj(b)
{
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// e^{i*pi*x}
__device__ __forceinline__ float2 cispif(float x)
{
    float s, c;
    sincospif(x, &s, &c);
    return make_float2(c, s);
}

// e^{sign * 2*pi*i * k / m} for integers 0 <= k < 2^24; the fraction is formed in
// float with one rounding (k and m exact), so the phase error is <= ~4e-7 rad.
__device__ __forceinline__ float2 twiddle_frac(int k, float inv_m, float sign)
{
    return cispif(sign * 2.0f * (float)k * inv_m);
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Loudness gain + hard clip + NaN scrub of one raw sample
// (reference lib.rs:220-227 then audio_pattern_detector.py:489-490):
// f64 multiply, clamp to [-1, 1] (NaN survives the clamp), round to f32, NaN -> 0.
__device__ __forceinline__ float normalize_sample(float x, double gain)
{
    // the clamp is applied after the rounding to float32: rounding is monotonic and +-1 are exact, so
    // clamp(round(v)) == round(clamp(v)), and it keeps the compares off the float64 pipe
    const float r = (float)((double)x * gain);
    return r != r ? 0.0f : fminf(fmaxf(r, -1.0f), 1.0f);
}

}  // namespace apd
