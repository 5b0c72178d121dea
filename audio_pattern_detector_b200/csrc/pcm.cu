// Streaming ingestion: integer PCM frames -> mono float32 on the device.
//
// Replaces the host conversion the reference does per chunk in _WavFileStreamWrapper.read /
// _normalize_wav_data (match.py:393-427, audio_utils.py:60-79,132-151): int16 / 32768, int32 / 2^31 after the
// int -> float32 rounding numpy's astype does, unsigned 8-bit (x - 128) / 128, channels averaged in float32 in
// channel order.  The raw PCM
// crosses PCIe (half the bytes of float32 for 16-bit audio) and is widened next to the detection kernels.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/apd_b200.h"

namespace {

template <typename T>
__device__ __forceinline__ float pcm_sample(T v);
template <> __device__ __forceinline__ float pcm_sample<int16_t>(int16_t v) { return (float)v / 32768.0f; }
template <> __device__ __forceinline__ float pcm_sample<int32_t>(int32_t v) { return __int2float_rn(v) / 2147483648.0f; }
template <> __device__ __forceinline__ float pcm_sample<uint8_t>(uint8_t v) { return ((float)v - 128.0f) / 128.0f; }

template <typename T>
__global__ void __launch_bounds__(256)
k_pcm_to_float(const T* __restrict__ pcm, int channels, long long n_frames, float* __restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += stride) {
        const T* p = pcm + f * channels;
        float acc = pcm_sample<T>(p[0]);
        for (int c = 1; c < channels; ++c) acc += pcm_sample<T>(p[c]);          // numpy mean(axis=1): float32, in order
        out[f] = channels > 1 ? acc / (float)channels : acc;
    }
}

}  // namespace

extern "C" int apd_pcm_to_float(const void* pcm_dev, int sample_width_bytes, int channels, int64_t n_frames,
                                float* out_dev, void* cuda_stream)
{
    if (!pcm_dev || !out_dev || channels < 1 || channels > 64 || n_frames < 0) return APD_ERR_INVALID;
    if (n_frames == 0) return APD_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int blocks = (int)((n_frames + 255) / 256 < 148 * 16 ? (n_frames + 255) / 256 : 148 * 16);
    if (sample_width_bytes == 2)
        k_pcm_to_float<int16_t><<<blocks, 256, 0, st>>>((const int16_t*)pcm_dev, channels, n_frames, out_dev);
    else if (sample_width_bytes == 4)
        k_pcm_to_float<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)pcm_dev, channels, n_frames, out_dev);
    else if (sample_width_bytes == 1)
        k_pcm_to_float<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)pcm_dev, channels, n_frames, out_dev);
    else
        return APD_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? APD_OK : APD_ERR_CUDA;
}
