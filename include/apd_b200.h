/*
 * apd_b200.h -- C ABI of the B200-native detection hot path.
 *
 * Plain C entry points (pointers + sizes, no torch / C++ types) over the CUDA
 * kernels in audio_pattern_detector_b200/csrc.  The Python host
 * (audio_pattern_detector_b200/_lib.py) binds them with ctypes and passes
 * PyTorch CUDA tensor data_ptr()s and the current stream handle.
 *
 * What each entry point replaces in the reference (andrewtheguy/audio_pattern_detector,
 * paths relative to the reference root):
 *
 *   apd_create            per-clip precompute of AudioPatternDetector.__init__
 *                         (audio_pattern_detector/audio_pattern_detector.py:155-221): clip loudness +
 *                         normalisation (_native.integrated_loudness / loudness_normalize,
 *                         native-helper/src/python.rs:146-171), clip self-correlation
 *                         (fft_correlation.fft_correlate_1d, audio_pattern_detector.py:373-383), the lazily
 *                         cached clip-side Pearson windows (:822-829).
 *   apd_scan              the body of the chunk loop, for a whole range of chunks at once:
 *                         _process_chunk + _correlation_method + _verify_peak_candidate
 *                         (audio_pattern_detector.py:389-640), i.e. the FFI calls
 *                         _native.integrated_loudness, _native.loudness_normalize,
 *                         fft_correlation.fft_correlate_1d, _native.find_peaks,
 *                         _native.resample_preserve_maxima, _native.pearson_correlation
 *                         (native-helper/src/python.rs:79-181) and the numpy FFTs of
 *                         detection_utils.analyze_pure_tone_candidate (detection_utils.py:41-125).
 *   apd_stage_*           the same stages one at a time (parity tests / profiling).
 *
 * Conventions: every function returns 0 on success or an APD_ERR_* code; the message is
 * available from apd_last_error().  Device pointers are caller-owned (PyTorch tensors); the
 * library allocates only its own workspace at apd_create.  All work is enqueued on the stream
 * passed in; functions that return results to host memory synchronise that stream.
 * One context per GPU / per host thread; no global mutable state besides the last-error string.
 */
#ifndef APD_B200_H
#define APD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define APD_API __attribute__((visibility("default")))
#else
#define APD_API
#endif

#define APD_OK 0
#define APD_ERR_INVALID 1
#define APD_ERR_CUDA 2
#define APD_ERR_UNSUPPORTED 3
#define APD_ERR_OVERFLOW 4

#define APD_STRATEGY_NORMAL 0
#define APD_STRATEGY_MARKER_TONE 1

#define APD_FLAG_ACCEPT 1
#define APD_FLAG_SKIPPED 2      /* failed the +-5 sample bounds gate (audio_pattern_detector.py:534-546) */
#define APD_KIND_SHIFT 2        /* (flags >> 2) & 3 : 0 normal, 1 short clip, 2 marker tone */

typedef struct apd_ctx apd_ctx;

/* One pattern clip, host memory.  Thresholds follow .apd.toml [verification]
 * (audio_pattern_detector.py:694-705); a NaN entry selects the reference default. */
typedef struct {
    const float* samples;         /* float32 mono at the detector sample rate (un-normalised, as loaded) */
    int32_t length;
    int32_t strategy;             /* APD_STRATEGY_* */
    double tone_hz;               /* marker tone only: dominant frequency (any number, 0 Hz included, as apd.py:214-221);
                                     NaN = none known: the clip takes the normal verifier (apd.py:605-625) */
    double minimum_band_purity;
    double minimum_active_frame_ratio;
    double minimum_longest_active_run;
    double minimum_active_frame_mean_purity;
    double maximum_min_flank_purity;
    double maximum_max_flank_purity;
} apd_clip_desc;

/* One verified (or rejected) candidate peak of one (chunk x clip) unit. */
typedef struct {
    int32_t chunk;                /* global chunk index */
    int32_t clip;                 /* index into the clip list given to apd_create */
    int32_t peak;                 /* sample index of the peak in the section's 'full' correlation */
    int32_t flags;                /* APD_FLAG_* | kind << APD_KIND_SHIFT */
    float height;                 /* normalised correlation at the peak */
    float similarity_whole;       /* normal/short: mean of the 10 partition MSEs */
    float similarity_middle;      /* normal/short: mean of partitions 4..5 */
    float reserved;
    double pearson[3];            /* normal: windows 0-5, 4-6, 5-10; short: [0] only (NaN if not computed) */
    double tone[3][5];            /* tone: {detected Hz, band purity, active ratio, longest run, mean purity}
                                     for the matched segment, left flank, right flank */
} apd_candidate;

/* Per (chunk x clip) unit trace, optional. */
typedef struct {
    float absmax;                 /* max |corr| before normalisation */
    float max_choose;             /* max(self-correlation max, absmax) */
    int32_t n_out;                /* length of the 'full' correlation */
    int32_t n_peaks;              /* peaks after height + distance filtering (-1: unit had none / not selected) */
} apd_unit_trace;

APD_API const char* apd_last_error(void);

/* sample_rate must be a multiple of 10.  chunk_samples = seconds_per_chunk * sample_rate.
 * height_min <= 0 selects the default 0.25.  max_batch_chunks bounds how many chunks one
 * apd_scan call may process (workspace is sized for it). */
APD_API int apd_create(apd_ctx** out, int device, int sample_rate, int64_t chunk_samples, float height_min,
               int n_clips, const apd_clip_desc* clips, int max_batch_chunks);
APD_API int apd_destroy(apd_ctx* ctx);

/* Pattern-side results of apd_create, copied to host (for parity tests and get_config). */
APD_API int apd_clip_info(apd_ctx* ctx, int clip, int32_t* sliding_window_seconds, double* lufs, float* self_max,
                  int32_t* fft_points /* real FFT length N used for this clip's group */);
APD_API int apd_clip_normalized(apd_ctx* ctx, int clip, float* out_host /* length samples */);
APD_API int apd_clip_self_correlation(apd_ctx* ctx, int clip, float* out_host /* 2*length-1 */);

/* Scan chunks [chunk_begin, chunk_end) of a device-resident stream.
 *   audio_dev      float32 device pointer; audio_dev[0] is stream sample `base_sample`
 *   n_samples      samples available at audio_dev (the stream region [base_sample, base_sample+n_samples))
 *   The region must contain chunk_begin's look-back (max sliding window) unless chunk_begin == 0,
 *   and chunk_end-1 may be the stream's final, shorter chunk.
 * Results are appended to host arrays (capacity in entries); *n_cand receives the count.
 * Candidates are ordered by (chunk, clip, peak).  unit_trace (optional, may be NULL) must hold
 * (chunk_end-chunk_begin)*n_clips entries; section_lufs (optional) the same count of doubles. */
APD_API int apd_scan(apd_ctx* ctx, const float* audio_dev, int64_t base_sample, int64_t n_samples,
             int32_t chunk_begin, int32_t chunk_end, apd_candidate* cand_host, int32_t cand_capacity,
             int32_t* n_cand, apd_unit_trace* unit_trace_host, double* section_lufs_host, void* cuda_stream);

/* Stage-level entry points (tests / profiling).  All operate on the same chunk range as a scan
 * and leave their results in the context workspace. */
APD_API int apd_stage_loudness(apd_ctx* ctx, const float* audio_dev, int64_t base_sample, int64_t n_samples,
                       int32_t chunk_begin, int32_t chunk_end, void* cuda_stream);
APD_API int apd_stage_forward_fft(apd_ctx* ctx, void* cuda_stream);
APD_API int apd_stage_correlate_max(apd_ctx* ctx, void* cuda_stream);     /* fused multiply + inverse FFT + |.| + max */
APD_API int apd_stage_peaks_verify(apd_ctx* ctx, void* cuda_stream);
/* Full normalised correlation of one unit of the staged batch, to host (n_out floats). */
APD_API int apd_stage_unit_correlation(apd_ctx* ctx, int32_t chunk, int32_t clip, float* out_host, int32_t capacity,
                               int32_t* n_out, void* cuda_stream);
/* Marker-tone verification of one candidate `peak` (index into the 'full' correlation, i.e. match start = peak - L + 1)
 * of an already normalised section in device memory (n <= chunk_samples): the three segment metrics (match, left flank,
 * right flank) x (detected frequency, band purity, active frame ratio, longest active run, mean active purity) and
 * the accept decision.  Replaces AudioPatternDetector._verify_marker_tone (reference audio_pattern_detector.py:660-750,
 * detection_utils.py:41-142) for callers that verify a single candidate (tests/test_marker_tone_verification.py:66). */
APD_API int apd_verify_tone(apd_ctx* ctx, int32_t clip, const float* section_dev, int32_t n_samples, int32_t peak,
                            double* metrics15_host, int32_t* accept, void* cuda_stream);

/* Stage timing: when enabled, apd_scan brackets its four stages with CUDA events on the caller's
 * stream and accumulates their device times (ms): [0] loudness, [1] forward FFT,
 * [2] fused multiply + inverse FFT + |.| + max, [3] peaks + verification. */
APD_API int apd_profile(apd_ctx* ctx, int enable);
APD_API int apd_profile_read(apd_ctx* ctx, double* ms4, int reset);

/* Streaming ingestion (row N1): integer PCM frames, already on the device, to mono float32 with the reference's
 * arithmetic (_WavFileStreamWrapper.read / _normalize_wav_data, match.py:393-427, audio_utils.py:60-79,132-151):
 * int16 / 32768, int32 / 2^31 or (uint8 - 128) / 128, channels averaged in float32.  sample_width_bytes is 1, 2 or
 * 4, channels 1..64.
 * Needs no context; returns APD_ERR_* without setting apd_last_error(). */
APD_API int apd_pcm_to_float(const void* pcm_dev, int sample_width_bytes, int channels, int64_t n_frames,
                     float* out_dev, void* cuda_stream);

/* FFT resampler (row N2): `batch` independent signals of n_in float32 samples each (signal b starts at
 * in_dev + b*in_stride) -> n_out samples each (out_dev + b*out_stride), with the arithmetic of _native.resample
 * (native-helper/src/python.rs:106-116 -> resample_1d, native-helper/src/lib.rs:235-275), which the reference applies
 * per chunk read in _WavFileStreamWrapper.read (match.py:421-423) via audio_utils.resample_audio
 * (audio_utils.py:154-171): float64 complex FFT of length n_in, the (N+1)/2 positive and (N-1)/2 negative bins of
 * N = min(n_in, n_out) kept, inverse FFT of length n_out, scale 1/n_in, rounded to float32.  Any lengths up to
 * 2^28 (7-smooth lengths as mixed-radix passes, others through Bluestein).  n_in == 0 gives zeros, n_in == n_out
 * a copy.  The caller owns the float64 workspace (apd_resample_workspace_bytes; 0 for the trivial cases).
 * Needs no context; returns APD_ERR_* without setting apd_last_error(). */
APD_API int apd_resample_workspace_bytes(int64_t n_in, int64_t n_out, int32_t batch, int64_t* bytes);
APD_API int apd_resample(const float* in_dev, int64_t n_in, int64_t in_stride, float* out_dev, int64_t n_out,
                 int64_t out_stride, int32_t batch, void* workspace_dev, int64_t workspace_bytes, void* cuda_stream);

/* Introspection for the bench: algorithmic byte counts and kernel launch counter. */
APD_API int64_t apd_launch_count(apd_ctx* ctx);
/* Phase-2 work since the last reset: out4 = {selected units (write-back inverse + find_peaks), candidate records,
 * marker-tone work items, sub-batches}. */
APD_API int apd_work_counters(apd_ctx* ctx, int64_t* out4, int reset);
APD_API int apd_unit_n_out(apd_ctx* ctx, int32_t chunk, int32_t clip, int64_t total_samples, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* APD_B200_H */
