"""B200-native ``AudioPatternDetector`` -- drop-in for the reference class of the same name.

Same constructor / ``find_clip_in_audio`` / ``get_config`` surface, exceptions and
result layout as the reference (audio_pattern_detector/audio_pattern_detector.py:84-371),
but the per-chunk work -- loudness normalisation, FFT cross-correlation against every
clip, peak picking and the Step-2 verifiers -- runs as batched CUDA kernels behind the
C ABI in include/apd_b200.h.  The host keeps only what the reference also does in
Python: argument validation, the chunk read loop, timestamp arithmetic in the
reference's exact order of float operations (:439-456, :585) and the per-chunk
ordering of callbacks (:324-327).

There is no CPU fallback: constructing a detector needs a CUDA device and the built
``libapd_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import os
import sys
from collections.abc import Callable
from dataclasses import dataclass
from typing import Any, Optional, TypedDict

import numpy as np
from numpy.typing import NDArray

from . import _lib
from .audio_clip import AudioClip, AudioStream
from .audio_utils import DEFAULT_TARGET_SAMPLE_RATE
from .detection_utils import get_pure_tone_frequency

logger = logging.getLogger(__name__)

DEFAULT_SECONDS_PER_CHUNK = 60            # reference :33
SHORT_CLIP_DURATION_THRESHOLD = 0.5       # reference :36
MARKER_TONE_STRATEGY = "marker_tone"      # reference :38
DEFAULT_BATCH_CHUNKS = 32

PatternDetectedCallback = Callable[[str, float], None]


class ClipData(TypedDict):
    """Per-clip data of the reference class (:42-48).  Here it lives on the device (apd_create); ``clip_info()``
    copies it back in this layout for inspection and parity tests."""
    clip: NDArray[np.float32]
    clip_name: str
    sliding_window: int
    correlation_clip: NDArray[np.float32]
    correlation_clip_absolute_max: "np.floating[Any]"


class ClipCache(TypedDict):
    """The reference's lazily filled cache of down-sampled clip curves (:51-54); computed eagerly on the device."""
    downsampled_correlation_clips: dict[str, NDArray[np.float32]]
    downsampled_pearson_windows: dict[str, list[NDArray[np.float32]]]


class ClipConfig(TypedDict):
    duration_seconds: float
    sliding_window_seconds: int


class DetectorConfig(TypedDict):
    default_seconds_per_chunk: int
    min_chunk_size_seconds: int
    sample_rate: int
    clips: dict[str, ClipConfig]


@dataclass
class Candidate:
    """One verified or rejected peak, with the scores the verifier computed (device results)."""
    chunk: int
    clip: str
    peak: int
    kind: str                 # "normal" | "short" | "tone"
    accept: bool
    skipped: bool
    height: float
    similarity_whole: float
    similarity_middle: float
    pearson: tuple[float, float, float]
    tone: tuple[tuple[float, ...], ...]
    timestamp: float


# memory image of apd_candidate (include/apd_b200.h), 176 bytes
CAND_DTYPE = np.dtype([("chunk", "<i4"), ("clip", "<i4"), ("peak", "<i4"), ("flags", "<i4"), ("height", "<f4"),
                       ("similarity_whole", "<f4"), ("similarity_middle", "<f4"), ("reserved", "<f4"),
                       ("pearson", "<f8", (3,)), ("tone", "<f8", (3, 5))])
assert CAND_DTYPE.itemsize == C.sizeof(_lib.Candidate)


class ScanResult:
    """Detections of a scan.  ``records`` is the raw candidate table (CAND_DTYPE, ordered by chunk, clip,
    peak) with a parallel ``timestamps`` array; ``candidates`` materialises :class:`Candidate` objects on
    first use (the hot path never does)."""

    def __init__(self, peak_times: dict[str, list[float]], events: list[tuple[float, str]],
                 records: "NDArray[Any]", timestamps: "NDArray[np.float64]", clip_names: list[str],
                 unit_trace: Optional[dict[tuple[int, str], dict[str, Any]]], total_time: float) -> None:
        self.peak_times = peak_times
        self.events = events                      # callback order
        self.records = records
        self.timestamps = timestamps
        self.unit_trace = unit_trace
        self.total_time = total_time
        self._clip_names = clip_names
        self._candidates: Optional[list[Candidate]] = None

    @property
    def n_candidates(self) -> int:
        return int(self.records.shape[0])

    @property
    def clip_names(self) -> list[str]:
        return list(self._clip_names)

    @property
    def candidates(self) -> list[Candidate]:
        if self._candidates is None:
            out = []
            for r, t in zip(self.records, self.timestamps):
                flags = int(r["flags"])
                out.append(Candidate(
                    chunk=int(r["chunk"]), clip=self._clip_names[int(r["clip"])], peak=int(r["peak"]),
                    kind=_lib.KINDS[(flags >> _lib.KIND_SHIFT) & 3], accept=bool(flags & _lib.FLAG_ACCEPT),
                    skipped=bool(flags & _lib.FLAG_SKIPPED), height=float(r["height"]),
                    similarity_whole=float(r["similarity_whole"]), similarity_middle=float(r["similarity_middle"]),
                    pearson=tuple(float(v) for v in r["pearson"]),
                    tone=tuple(tuple(float(v) for v in seg) for seg in r["tone"]), timestamp=float(t)))
            self._candidates = out
        return self._candidates


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("audio_pattern_detector_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


class AudioPatternDetector:

    def __init__(self, audio_clips: list[AudioClip], debug_mode: bool = False,
                 seconds_per_chunk: int | None = DEFAULT_SECONDS_PER_CHUNK,
                 target_sample_rate: int | None = None, debug_dir: str = "./tmp",
                 height_min: float | None = None, *, device: int | None = None,
                 max_batch_chunks: int | None = None, stream_read_chunks: int | None = None) -> None:
        self.audio_clips = audio_clips
        self.debug_mode = debug_mode
        self.debug_dir = debug_dir
        self.height_min = height_min
        self.normalize = True
        self.target_sample_rate = target_sample_rate if target_sample_rate is not None else DEFAULT_TARGET_SAMPLE_RATE
        sr = self.target_sample_rate

        seen: set[str] = set()
        longest = 0
        for clip in audio_clips:                                           # reference :105-115
            if clip.name in seen:
                raise ValueError(f"clip {clip.name} needs to be unique")
            if clip.sample_rate != sr:
                raise ValueError(f"clip {clip.name} needs to be {sr} sample rate")
            seen.add(clip.name)
            longest = max(longest, len(clip.audio))

        if seconds_per_chunk is None or seconds_per_chunk < 1:             # reference :117-120
            seconds_per_chunk = math.ceil(longest / sr) * 2
            logger.warning("seconds_per_chunk is not set or less than 1, setting it to longest clip * 2 seconds, "
                           f"which is {seconds_per_chunk} seconds")
        self._min_chunk_size = 0
        for clip in audio_clips:                                           # reference :122-137
            secs = len(clip.audio) / sr
            sw = math.ceil(secs)
            self._min_chunk_size = max(self._min_chunk_size, sw * 2)
            if seconds_per_chunk < sw * 2:
                raise ValueError(f"seconds_per_chunk {seconds_per_chunk} is too small for clip '{clip.name}' "
                                 f"(duration: {secs:.2f}s, sliding_window: {sw}s, minimum chunk size: {sw * 2}s)")
        self.seconds_per_chunk = seconds_per_chunk
        if seconds_per_chunk != 60:                                        # reference :141-143
            logger.warning(f"seconds_per_chunk {seconds_per_chunk} is not 60 seconds, turning off debug mode "
                           "because it was made for 60 seconds only")
            self.debug_mode = False

        self._clip_strategies: dict[str, str | None] = {}
        self._clip_strategy_params: dict[str, dict[str, Any]] = {}
        self._tone_frequencies: dict[str, float] = {}
        self._clip_lengths = [len(c.audio) for c in audio_clips]
        self._sliding_windows = [math.ceil(n / sr) for n in self._clip_lengths]
        self._sw_arr = np.asarray(self._sliding_windows, dtype=np.int64)
        self._clip_seconds_arr = np.asarray([n / sr for n in self._clip_lengths], dtype=np.float64)
        self._chunk_samples = int(seconds_per_chunk * sr)
        self._chunk_size = self._chunk_samples * 4                         # reference :224

        for clip, sw in zip(audio_clips, self._sliding_windows):
            secs = len(clip.audio) / sr
            if sw != secs:                                                 # reference :162-163
                print(f"adjusted sliding_window from {secs} to {sw} for {clip.name}", file=sys.stderr)

        # ---- hand the clips to the device (pattern-side precompute happens in apd_create)
        torch = _torch()
        self._device = torch.cuda.current_device() if device is None else int(device)
        self._max_batch = int(max_batch_chunks or os.environ.get("APD_B200_BATCH_CHUNKS", DEFAULT_BATCH_CHUNKS))
        # find_clip_in_audio reads this many chunks per device scan: several sub-batches, so that apd_scan can overlap
        # its phases; a caller that wants the reference's per-chunk callback latency on a live stream passes 1
        self._stream_read_chunks = int(stream_read_chunks) if stream_read_chunks else 4 * self._max_batch
        descs = (_lib.ClipDesc * len(audio_clips))()
        self._keepalive: list[NDArray[np.float32]] = []
        nan = float("nan")
        for d, clip in zip(descs, audio_clips):
            a = np.ascontiguousarray(clip.audio, dtype=np.float32)
            self._keepalive.append(a)
            d.samples = a.ctypes.data_as(C.POINTER(C.c_float))
            d.length = a.size
            d.strategy = _lib.STRATEGY_NORMAL
            d.tone_hz = nan                                                # none: normal verifier (reference :605-625)
            for k in ("minimum_band_purity", "minimum_active_frame_ratio", "minimum_longest_active_run",
                      "minimum_active_frame_mean_purity", "maximum_min_flank_purity", "maximum_max_flank_purity"):
                setattr(d, k, nan)
            self._clip_strategies[clip.name] = clip.strategy
            self._clip_strategy_params[clip.name] = dict(clip.strategy_params)
            if clip.strategy == MARKER_TONE_STRATEGY:                      # reference :214-221
                d.strategy = _lib.STRATEGY_MARKER_TONE
                freq = clip.strategy_params.get("dominant_frequency_hz")
                if freq is None:
                    freq = self._fallback_tone_frequency(a)
                if freq is not None:
                    self._tone_frequencies[clip.name] = float(freq)
                    d.tone_hz = float(freq)
                ver = clip.strategy_params.get("verification", {})
                if isinstance(ver, dict):                                  # reference :694-705
                    for k, v in ver.items():
                        if hasattr(d, k):
                            setattr(d, k, float(int(v)) if k == "minimum_longest_active_run" else float(v))
        ctx = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.apd_create(C.byref(ctx), self._device, sr, self._chunk_samples,
                                float(height_min) if height_min is not None else float("nan"),
                                len(audio_clips), descs, self._max_batch), "apd_create")
        self._ctx = ctx
        self._max_halo = max(self._sliding_windows, default=0) * sr
        self._dev_buf = None
        self._pinned = None
        self._cand_buf = None
        self._cand_cap_boost = 1

    # ------------------------------------------------------------------ helpers
    def _fallback_tone_frequency(self, raw: NDArray[np.float32]) -> Optional[float]:
        """Reference :217-219: a marker-tone clip without a declared frequency gets the dominant frequency of its
        loudness-NORMALISED samples (gain to -16 LUFS, clamped to +-1 -- the clamp can add harmonics).  The
        normalisation is the device's (apd_create on the clip alone, read back with apd_clip_normalized)."""
        sr = self.target_sample_rate
        L = _lib.lib()
        desc = (_lib.ClipDesc * 1)()
        desc[0].samples = raw.ctypes.data_as(C.POINTER(C.c_float))
        desc[0].length = raw.size
        desc[0].strategy = _lib.STRATEGY_NORMAL
        ctx = C.c_void_p()
        chunk = max(self._chunk_samples, 2 * math.ceil(raw.size / sr) * sr)
        _lib.check(L.apd_create(C.byref(ctx), self._device, sr, chunk, float("nan"), 1, desc, 1), "apd_create")
        try:
            norm = np.empty(raw.size, dtype=np.float32)
            _lib.check(L.apd_clip_normalized(ctx, 0, norm.ctypes.data_as(C.POINTER(C.c_float))), "apd_clip_normalized")
        finally:
            L.apd_destroy(ctx)
        return get_pure_tone_frequency(norm, sr)

    def __del__(self) -> None:
        ctx = getattr(self, "_ctx", None)
        if ctx:
            try:
                _lib.lib().apd_destroy(ctx)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._ctx = None

    def close(self) -> None:
        self.__del__()

    def get_config(self) -> DetectorConfig:
        """reference :226-246."""
        clips: dict[str, ClipConfig] = {}
        for clip, sw in zip(self.audio_clips, self._sliding_windows):
            clips[clip.name] = {"duration_seconds": round(len(clip.audio) / self.target_sample_rate, 6),
                                "sliding_window_seconds": sw}
        return {"default_seconds_per_chunk": DEFAULT_SECONDS_PER_CHUNK, "min_chunk_size_seconds": self._min_chunk_size,
                "sample_rate": self.target_sample_rate, "clips": clips}

    def clip_info(self, index: int) -> dict[str, Any]:
        """Pattern-side device results (loudness, self-correlation max, FFT size) for tests/inspection."""
        sw, lufs, smax, nfft = C.c_int32(), C.c_double(), C.c_float(), C.c_int32()
        _lib.check(_lib.lib().apd_clip_info(self._ctx, index, C.byref(sw), C.byref(lufs), C.byref(smax),
                                            C.byref(nfft)), "apd_clip_info")
        n = self._clip_lengths[index]
        norm = np.empty(n, dtype=np.float32)
        cc = np.empty(2 * n - 1, dtype=np.float32)
        _lib.check(_lib.lib().apd_clip_normalized(self._ctx, index, norm.ctypes.data_as(C.POINTER(C.c_float))), "clip")
        _lib.check(_lib.lib().apd_clip_self_correlation(self._ctx, index, cc.ctypes.data_as(C.POINTER(C.c_float))),
                   "clip")
        return {"sliding_window": sw.value, "lufs": lufs.value, "self_max": smax.value, "fft_points": nfft.value,
                "normalized": norm, "self_correlation": cc}

    def enable_profiling(self, on: bool = True) -> None:
        _lib.check(_lib.lib().apd_profile(self._ctx, 1 if on else 0), "apd_profile")

    def stage_times_ms(self, reset: bool = True) -> dict[str, float]:
        ms = (C.c_double * 4)()
        _lib.check(_lib.lib().apd_profile_read(self._ctx, ms, 1 if reset else 0), "apd_profile_read")
        return dict(zip(("loudness", "forward_fft", "correlate_max", "peaks_verify"), ms))

    def time_correlate_stage(self, audio_dev: Any) -> tuple[float, int]:
        """Device time (ms, CUDA events on the current stream) of the fused correlate + max stage alone over a
        whole device-resident stream: per sub-batch the loudness and forward-FFT stages are run untimed, then
        apd_stage_correlate_max is bracketed by events with nothing else on the GPU.  Returns (ms, launches)."""
        torch = _torch()
        L = _lib.lib()
        n = audio_dev.numel()
        C_ = self._chunk_samples
        n_chunks = (n + C_ - 1) // C_
        total_ms, launches = 0.0, 0
        with torch.cuda.device(self._device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for c0 in range(0, n_chunks, self._max_batch):
                c1 = min(n_chunks, c0 + self._max_batch)
                _lib.check(L.apd_stage_loudness(self._ctx, C.c_void_p(audio_dev.data_ptr()), 0, n, c0, c1, st), "stage")
                _lib.check(L.apd_stage_forward_fft(self._ctx, st), "stage")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0 = self.launch_count()
                e0.record()
                _lib.check(L.apd_stage_correlate_max(self._ctx, st), "stage")
                e1.record()
                e1.synchronize()
                total_ms += e0.elapsed_time(e1)
                launches += self.launch_count() - l0
        return total_ms, launches

    def launch_count(self) -> int:
        return int(_lib.lib().apd_launch_count(self._ctx))

    def work_counters(self, reset: bool = True) -> dict[str, int]:
        """Phase-2 work since the last reset (what the write-back inverse, find_peaks and the verifiers ran on)."""
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib().apd_work_counters(self._ctx, out, 1 if reset else 0), "apd_work_counters")
        return dict(zip(("selected_units", "candidate_records", "tone_items", "sub_batches"), (int(v) for v in out)))

    def unit_n_out(self, chunk: int, clip_index: int, total_samples: int) -> int:
        n = C.c_int32()
        _lib.check(_lib.lib().apd_unit_n_out(self._ctx, chunk, clip_index, total_samples, C.byref(n)), "unit_n_out")
        return n.value

    # ------------------------------------------------------------------ single-candidate verifier (reference :642-750)
    def tone_candidate_metrics(self, clip_name: str, audio_section: "NDArray[np.float32]", peak: int
                               ) -> tuple[bool, tuple[tuple[float, ...], ...]]:
        """Device marker-tone verification of one candidate of an already normalised section (apd_verify_tone):
        (accept, ((detected_frequency, overall_band_purity, active_frame_ratio, longest_active_run,
        active_frame_mean_purity) for the matched segment, the left flank and the right flank))."""
        torch = _torch()
        idx = [c.name for c in self.audio_clips].index(clip_name)
        sec = np.ascontiguousarray(audio_section, dtype=np.float32)
        metrics = (C.c_double * 15)()
        accept = C.c_int32()
        with torch.cuda.device(self._device):
            dev = torch.from_numpy(sec).cuda()
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().apd_verify_tone(self._ctx, idx, C.c_void_p(dev.data_ptr()), sec.size, int(peak),
                                                  metrics, C.byref(accept), st), "apd_verify_tone")
        m = tuple(tuple(float(metrics[s * 5 + j]) for j in range(5)) for s in range(3))
        return bool(accept.value), m

    def _verify_marker_tone(self, clip_name: str, audio_section: "NDArray[np.float32]", peak: int, clip_length: int,
                            dominant_frequency: float, sr: int, section_ts: str) -> bool:
        """Signature of the reference's verifier (:660-668).  The clip's length, frequency and thresholds were handed
        to the device at construction; arguments that disagree with them cannot be honoured and raise."""
        idx = [c.name for c in self.audio_clips].index(clip_name)
        if clip_length != self._clip_lengths[idx] or sr != self.target_sample_rate:
            raise ValueError("_verify_marker_tone: clip_length / sr differ from the detector's clip")
        if not math.isclose(float(dominant_frequency), self._tone_frequencies.get(clip_name, float("nan")), rel_tol=1e-12):
            raise ValueError("_verify_marker_tone: dominant_frequency differs from the detector's clip")
        return self.tone_candidate_metrics(clip_name, audio_section, peak)[0]

    # ------------------------------------------------------------------ timestamps
    def _timestamps(self, rec: "NDArray[Any]") -> "NDArray[np.float64]":
        """reference :585 then :440-451, same order of float64 operations, for a whole candidate table."""
        sr = self.target_sample_rate
        clip = rec["clip"]
        chunk = rec["chunk"].astype(np.int64)
        t = rec["peak"].astype(np.float64) / sr
        t = t - np.where(chunk > 0, self._sw_arr[clip], 0).astype(np.float64)
        t = t + (chunk * int(self.seconds_per_chunk)).astype(np.float64)
        t = t - self._clip_seconds_arr[clip]
        return np.where(t >= 0, t, 0.0)

    # ------------------------------------------------------------------ device scan of a chunk range
    def _scan_batch(self, dev_ptr: int, base_sample: int, n_samples: int, chunk_begin: int, chunk_end: int,
                    want_trace: bool) -> tuple["NDArray[Any]", Optional[dict]]:
        """apd_scan over chunks [chunk_begin, chunk_end); returns (candidate table, unit trace).  The device lists are
        sized for the average case of a sub-batch and for the worst case of one chunk (a unit keeps up to N_out / L
        peaks, e.g. a 0.15 s tone clip inside minutes of steady tone): when a scan reports a workspace overflow the
        range is scanned again in halves, so dense material costs time, never detections."""
        try:
            return self._scan_once(dev_ptr, base_sample, n_samples, chunk_begin, chunk_end, want_trace)
        except _lib.ApdError as e:
            if e.code != 4:
                raise
            if chunk_end - chunk_begin <= 1:
                if self._cand_cap_boost >= 1 << 10:
                    raise
                self._cand_cap_boost *= 4                      # host-side record buffer: grow and retry
                self._cand_buf = None
                return self._scan_batch(dev_ptr, base_sample, n_samples, chunk_begin, chunk_end, want_trace)
        mid = (chunk_begin + chunk_end) // 2
        ra, ta = self._scan_batch(dev_ptr, base_sample, n_samples, chunk_begin, mid, want_trace)
        rb, tb = self._scan_batch(dev_ptr, base_sample, n_samples, mid, chunk_end, want_trace)
        if ta is not None and tb is not None:
            ta.update(tb)
        return np.concatenate([ra, rb]), ta

    def _scan_once(self, dev_ptr: int, base_sample: int, n_samples: int, chunk_begin: int, chunk_end: int,
                   want_trace: bool) -> tuple["NDArray[Any]", Optional[dict]]:
        torch = _torch()
        L = _lib.lib()
        nb = chunk_end - chunk_begin
        ncl = len(self.audio_clips)
        cap = max(4096, nb * ncl * 4) * self._cand_cap_boost
        if self._cand_buf is None or len(self._cand_buf) < cap:
            self._cand_buf = (_lib.Candidate * cap)()
        cands = self._cand_buf
        cap = len(cands)
        n = C.c_int32(0)
        trace = (_lib.UnitTrace * (nb * ncl))() if want_trace else None
        lufs = (C.c_double * (nb * ncl))() if want_trace else None
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(L.apd_scan(self._ctx, C.c_void_p(dev_ptr), base_sample, n_samples, chunk_begin, chunk_end,
                              cands, cap, C.byref(n), trace, lufs, C.c_void_p(stream)), "apd_scan")
        rec = np.frombuffer(cands, dtype=CAND_DTYPE, count=n.value).copy()
        tr = None
        if want_trace:
            tr = {}
            for ci in range(nb):
                for p in range(ncl):
                    u = trace[ci * ncl + p]
                    tr[(chunk_begin + ci, self.audio_clips[p].name)] = {
                        "absmax": u.absmax, "max_choose": u.max_choose, "n_out": u.n_out,
                        "n_peaks": u.n_peaks, "lufs": lufs[ci * ncl + p]}
        return rec, tr

    def _dump_debug(self, rec: "NDArray[Any]") -> None:
        """``--debug`` score dump (row N4; reference :563-582, :793-804, :880-895): per (chunk, clip) with peaks,
        ``{debug_dir}/debug/cross_correlation_{clip}/{index}_{section_ts}.txt`` holding the find_peaks result, the
        seconds of the verified candidates and their similarity / Pearson scores.  (The reference's PNG graphs and
        WAV excerpts are debug tooling outside the hot path and are not produced.)"""
        import json
        from .audio_utils import seconds_to_time
        sr = self.target_sample_rate
        names = [c.name for c in self.audio_clips]
        keys = np.stack([rec["chunk"], rec["clip"]], axis=1) if rec.size else np.zeros((0, 2), dtype=np.int32)
        start = 0
        while start < rec.shape[0]:
            end = start
            while end < rec.shape[0] and (keys[end] == keys[start]).all():
                end += 1
            index, clip = int(keys[start][0]), int(keys[start][1])
            peaks, seconds, similarities = [], [], []
            for r in rec[start:end]:
                flags = int(r["flags"])
                kind = (flags >> _lib.KIND_SHIFT) & 3
                peaks.append(int(r["peak"]))
                if kind == 2 or flags & _lib.FLAG_SKIPPED:            # tone clips / bounds gate: no similarity entry
                    continue
                seconds.append(int(r["peak"]) / sr)
                whole, middle = float(r["similarity_whole"]), float(r["similarity_middle"])
                sim = whole if kind == 1 else min(whole, middle)
                parts = {"whole": whole, "middle": middle}
                pr = [float(v) for v in r["pearson"]]
                if np.isnan(pr[0]):                                    # rejected on similarity before Pearson
                    similarities.append((sim, parts, None))
                    continue
                wins = [(0, 10)] if kind == 1 else [(0, 5), (4, 6), (5, 10)]
                best = max(range(len(wins)), key=lambda i: (pr[i], -i))
                d = {"pearson_r": pr[0 if kind == 1 else 1], "best_window_left": float(wins[best][0]),
                     "best_window_right": float(wins[best][1])}
                d.update({f"pearson_w{a}_{b}": pr[i] for i, (a, b) in enumerate(wins)})
                similarities.append((sim, parts, d))
            section_ts = seconds_to_time(seconds=index * self.seconds_per_chunk, include_decimals=False)
            out_dir = f"{self.debug_dir}/debug/cross_correlation_{names[clip]}"
            os.makedirs(out_dir, exist_ok=True)
            with open(f"{out_dir}/{index}_{section_ts}.txt", "w") as f:
                print(json.dumps({"peaks": peaks, "seconds": seconds, "similarities": similarities}, indent=2), file=f)
            start = end

    def _emit_batch(self, rec: "NDArray[Any]", ts: "NDArray[np.float64]",
                    peak_times: Optional[dict[str, list[float]]], events: list[tuple[float, str]],
                    on_pattern_detected: Optional[PatternDetectedCallback]) -> None:
        """Accepted candidates -> results, in the reference's order (:303-327): per chunk the clips in list
        order (the table is ordered by chunk, clip, peak), then a stable sort by timestamp inside the chunk."""
        if self.debug_mode:
            self._dump_debug(rec)
        acc = np.flatnonzero(rec["flags"] & _lib.FLAG_ACCEPT)
        if acc.size == 0:
            return
        names = [c.name for c in self.audio_clips]
        clip = rec["clip"][acc]
        t = ts[acc]
        if peak_times is not None:
            for k, tv in zip(clip.tolist(), t.tolist()):
                peak_times[names[k]].append(tv)
        order = np.lexsort((t, rec["chunk"][acc]))          # stable: ties keep (clip, peak) order
        for k, tv in zip(clip[order].tolist(), t[order].tolist()):
            events.append((tv, names[k]))
            if on_pattern_detected:
                on_pattern_detected(names[k], tv)

    # ------------------------------------------------------------------ public scanning API
    def scan_array(self, audio: "NDArray[np.float32] | Any", on_pattern_detected: Optional[PatternDetectedCallback] = None,
                   collect_trace: bool = False, chunk_range: Optional[tuple[int, int]] = None,
                   base_sample: int = 0, total_samples: Optional[int] = None, pcm_channels: int = 1) -> ScanResult:
        """Scan a whole in-memory stream (numpy array or CUDA float32 tensor).

        B200-side extension of the reference API: the stream is made device resident once and
        scanned ``max_batch_chunks`` chunks per launch sequence.

        ``audio`` may also be interleaved integer PCM on the host (int16 / int32 / uint8 array or tensor, ``pcm_channels``
        channels): the frames are copied as they are and widened to mono float32 on the device with the
        reference's arithmetic (row N1).

        Sharded use (sharding.py): ``audio`` is a slab of a longer stream starting at stream sample
        ``base_sample`` (it must contain the look-back halo of ``chunk_range[0]``), ``chunk_range``
        the global chunk indices this call scans and ``total_samples`` the length of the whole
        stream (only its final chunk is shorter)."""
        torch = _torch()
        devname = f"cuda:{self._device}"
        host = None
        pcm_width = 0                     # > 0: host holds interleaved integer PCM frames (int16 / int32)
        if isinstance(audio, np.ndarray):
            if audio.dtype in (np.int16, np.int32, np.uint8):
                pcm_width = audio.dtype.itemsize
                host = torch.from_numpy(np.ascontiguousarray(audio).reshape(-1))
            else:
                host = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        elif not audio.is_cuda:
            if audio.dtype in (torch.int16, torch.int32, torch.uint8):
                pcm_width = audio.element_size()
                host = audio.contiguous().reshape(-1)
            else:
                host = audio.to(dtype=torch.float32).contiguous()
        if host is not None:
            # host input: copied segment by segment on a copy stream, overlapped with the scan of the
            # previous segment (pinned host memory copies at full PCIe rate); integer PCM is widened on the
            # device (apd_pcm_to_float), so only the raw frames cross PCIe
            if pcm_width and host.numel() % pcm_channels:
                raise ValueError("PCM input length is not a multiple of the channel count")
            dev = torch.empty(host.numel() // (pcm_channels if pcm_width else 1), dtype=torch.float32, device=devname)
        else:
            dev = audio.to(device=devname, dtype=torch.float32).contiguous()
        n = dev.numel()
        C_ = self._chunk_samples
        end_sample = base_sample + n                    # one past the last stream sample held
        if total_samples is not None and total_samples < end_sample:
            raise ValueError("total_samples is smaller than the slab")
        n_chunks = (end_sample + C_ - 1) // C_
        first, last = (0, n_chunks) if chunk_range is None else chunk_range
        if not 0 <= first <= last <= n_chunks:
            raise ValueError(f"chunk_range {chunk_range} outside the slab's chunks [0, {n_chunks})")
        if total_samples is not None and last < n_chunks and last * C_ > end_sample:
            raise ValueError("slab ends inside a chunk that is not the stream's last")
        peak_times: dict[str, list[float]] = {c.name: [] for c in self.audio_clips}
        events: list[tuple[float, str]] = []
        all_rec: list[Any] = []
        all_ts: list[Any] = []
        trace: Optional[dict] = {} if collect_trace else None
        with torch.cuda.device(self._device):
            # one apd_scan per segment: the C side cuts a segment into sub-batches of max_batch_chunks chunks
            # and overlaps phase 2 of one sub-batch with phase 1 of the next
            if host is not None:
                # a short first segment (its copy is the only one not hidden), then segments that double in length
                # up to a cap: the copy of segment i + 1 runs while segment i is scanned, and a pinned H2D copy moves
                # a chunk about twice as fast as the device scans one, so doubling keeps every later copy hidden
                cap = self._max_batch * int(os.environ.get("APD_B200_HOST_SEG", 8))
                bounds, step = [first], self._max_batch
                while bounds[-1] < last:
                    bounds.append(min(last, bounds[-1] + step))
                    step = min(2 * step, cap)
            else:
                bounds = [first, last]
            if last <= first:
                bounds = [first]
            copied = 0                                   # slab samples already enqueued for copy
            copy_stream = torch.cuda.Stream() if host is not None else None
            ready: list[Any] = []

            raw_dev = None                               # device staging of raw PCM frames (largest segment)
            if pcm_width:
                longest = max(b - a for a, b in zip(bounds[:-1], bounds[1:])) if len(bounds) > 1 else 0
                raw_dev = torch.empty(min(n, longest * C_ + self._max_halo) * pcm_channels, dtype=host.dtype,
                                      device=devname)

            def enqueue_copy(upto_chunk: int) -> None:
                nonlocal copied
                hi = min(n, upto_chunk * C_ - base_sample)
                if hi > copied:
                    with torch.cuda.stream(copy_stream):
                        if pcm_width:
                            k = (hi - copied) * pcm_channels
                            raw_dev[:k].copy_(host[copied * pcm_channels:hi * pcm_channels], non_blocking=True)
                            self._pcm_to_float(raw_dev, pcm_width, pcm_channels, hi - copied, dev[copied:], copy_stream)
                        else:
                            dev[copied:hi].copy_(host[copied:hi], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    ready.append(ev)
                    copied = hi

            if host is not None and len(bounds) > 1:
                enqueue_copy(bounds[1])
            for si in range(len(bounds) - 1):
                c0, c1 = bounds[si], bounds[si + 1]
                if host is not None:
                    torch.cuda.current_stream().wait_event(ready[si])
                    if si + 2 < len(bounds):
                        enqueue_copy(bounds[si + 2])
                rec, tr = self._scan_batch(dev.data_ptr(), base_sample, n, c0, c1, collect_trace)
                ts = self._timestamps(rec)
                all_rec.append(rec)
                all_ts.append(ts)
                if trace is not None and tr:
                    trace.update(tr)
                self._emit_batch(rec, ts, peak_times, events, on_pattern_detected)
        total = 0.0
        for i in range(first, last):                                        # reference :301
            total += (min((i + 1) * C_, end_sample) - i * C_) / self.target_sample_rate
        records = np.concatenate(all_rec) if all_rec else np.zeros(0, dtype=CAND_DTYPE)
        stamps = np.concatenate(all_ts) if all_ts else np.zeros(0, dtype=np.float64)
        return ScanResult(peak_times, events, records, stamps, [c.name for c in self.audio_clips], trace, total)

    def _pcm_to_float(self, raw_dev: Any, sampwidth: int, channels: int, frames: int, out_dev: Any, stream: Any) -> None:
        rc = _lib.lib().apd_pcm_to_float(C.c_void_p(raw_dev.data_ptr()), sampwidth, channels, frames,
                                         C.c_void_p(out_dev.data_ptr()), C.c_void_p(stream.cuda_stream))
        if rc != _lib.APD_OK:
            raise ValueError(f"unsupported PCM format: {sampwidth * 8}-bit, {channels} channel(s)")

    def _pcm_chunks_fit(self, src: Any) -> bool:
        sr, C_ = self.target_sample_rate, self._chunk_samples
        in_sr = int(getattr(src, "pcm_sample_rate", None) or sr)
        if in_sr == sr:
            return True
        Cin = int(C_ * in_sr / sr)                                            # frames per chunk read, match.py:398
        return Cin >= 1 and int(Cin * sr / in_sr) == C_

    def _find_clip_in_pcm(self, src: Any, peak_times: Optional[dict[str, list[float]]],
                          on_pattern_detected: Optional[PatternDetectedCallback]
                          ) -> tuple[dict[str, list[float]] | None, float]:
        """Chunk loop over a raw-PCM source: ``src.read_pcm(frames) -> bytes`` and ``src.pcm_format = (sample
        width in bytes, channels)``.  Same batches, look-back and callback order as find_clip_in_audio.

        ``src.pcm_sample_rate`` (optional) is the rate of the frames when it differs from the detector's: every
        chunk read is then resampled on its own, on the device, exactly as the reference resamples what each
        ``read`` returns (match.py:395-423: ``int(chunk * in_rate / rate)`` frames in, ``int(len * rate / in_rate)``
        samples out, row N2)."""
        torch = _torch()
        sampwidth, channels = src.pcm_format
        np_dt, t_dt = {1: (np.uint8, torch.uint8), 2: (np.int16, torch.int16), 4: (np.int32, torch.int32)}[sampwidth]
        sr, C_ = self.target_sample_rate, self._chunk_samples
        in_sr = int(getattr(src, "pcm_sample_rate", None) or sr)
        resampling = in_sr != sr
        Cin = int(C_ * in_sr / sr) if resampling else C_                      # frames per chunk read, match.py:398
        if resampling and (Cin < 1 or int(Cin * sr / in_sr) != C_):
            raise ValueError(f"cannot cut a {in_sr} Hz source into chunks of {C_} samples at {sr} Hz")
        events: list[tuple[float, str]] = []
        total_time = 0.0
        dev = f"cuda:{self._device}"
        cap = self._max_halo + self._stream_read_chunks * C_
        from concurrent.futures import ThreadPoolExecutor

        from .resample import resample_into
        per_read = self._stream_read_chunks
        with torch.cuda.device(self._device), ThreadPoolExecutor(1) as pool:
            stream = torch.cuda.current_stream()
            fbuf = torch.empty(cap, dtype=torch.float32, device=dev)
            fin = torch.empty(per_read * Cin, dtype=torch.float32, device=dev) if resampling else None
            pins = [torch.empty(per_read * Cin * channels, dtype=t_dt).pin_memory() for _ in range(2)]
            raw_dev = torch.empty_like(pins[0], device=dev)

            def read_batch(which: int, chunks: int) -> int:
                """One large read of up to `chunks` chunks into pinned buffer `which`; returns the frames read.
                Runs on the reader thread, overlapped with the device scan of the previous batch (apd_scan
                releases the GIL)."""
                if hasattr(src, "readinto_pcm"):                              # straight into pinned memory
                    return src.readinto_pcm(pins[which].numpy(), chunks * Cin)
                # a pipe may hand out fewer frames than asked for before its end: read on until the batch is full or
                # the source is exhausted (a short batch means end of stream below)
                frame = sampwidth * channels
                data = b""
                while len(data) < chunks * Cin * frame:
                    more = src.read_pcm(chunks * Cin - len(data) // frame)
                    if not more:
                        break
                    data += more
                got = len(data) // frame
                if got:
                    pins[which].numpy()[:got * channels] = np.frombuffer(data, dtype=np_dt, count=got * channels)
                return got

            # the first read is the only one the device waits for: keep it short, then double up to per_read
            n_halo, chunk_index, which = 0, 0, 0
            want = min(per_read, self._max_batch)
            pending = pool.submit(read_batch, which, want)
            while True:
                frames = pending.result()
                if frames == 0:
                    break
                last = frames < want * Cin                                   # a short read ends the stream
                cur = which
                if not last:
                    which ^= 1
                    want = min(per_read, 2 * want)
                    pending = pool.submit(read_batch, which, want)           # next batch while this one is scanned
                raw_dev[:frames * channels].copy_(pins[cur][:frames * channels], non_blocking=True)
                if resampling:
                    self._pcm_to_float(raw_dev, sampwidth, channels, frames, fin, stream)
                    full, rem = divmod(frames, Cin)
                    tail = int(rem * sr / in_sr)                             # audio_utils.py:170
                    if full:
                        resample_into(fin, Cin, fbuf[n_halo:], C_, full, stream)
                    if tail:
                        resample_into(fin[full * Cin:], rem, fbuf[n_halo + full * C_:], tail, 1, stream)
                    new = full * C_ + tail
                else:
                    self._pcm_to_float(raw_dev, sampwidth, channels, frames, fbuf[n_halo:], stream)
                    new = frames
                if new == 0:
                    break
                n_chunks = (new + C_ - 1) // C_
                for k in range(n_chunks):                                    # reference :301, chunk by chunk
                    total_time += (min((k + 1) * C_, new) - k * C_) / sr
                n_tot = n_halo + new
                c0, c1 = chunk_index, chunk_index + n_chunks
                rec, _ = self._scan_batch(fbuf.data_ptr(), c0 * C_ - n_halo, n_tot, c0, c1, False)
                self._emit_batch(rec, self._timestamps(rec), peak_times, events, on_pattern_detected)
                keep = min(self._max_halo, n_tot)
                fbuf[:keep] = fbuf[n_tot - keep:n_tot].clone()               # look-back stays on the device
                n_halo = keep
                chunk_index = c1
                if last:
                    break
        return peak_times, total_time

    def find_clip_in_audio(self, audio_stream: AudioStream,
                           on_pattern_detected: PatternDetectedCallback | None = None,
                           accumulate_results: bool = True) -> tuple[dict[str, list[float]] | None, float]:
        """reference :248-371.  Reads ``seconds_per_chunk`` at a time like the reference, but hands
        up to ``max_batch_chunks`` chunks to the device per scan; callbacks still arrive in the
        reference's order (chunk by chunk, sorted by timestamp inside a chunk)."""
        if audio_stream.sample_rate != self.target_sample_rate:
            raise ValueError(f"full_streaming_audio_clip {audio_stream.name} needs to be "
                             f"{self.target_sample_rate} sample rate")
        torch = _torch()
        sr = self.target_sample_rate
        C_ = self._chunk_samples
        peak_times: dict[str, list[float]] | None = (
            {c.name: [] for c in self.audio_clips} if accumulate_results else None)
        events: list[tuple[float, str]] = []
        total_time = 0.0
        src = audio_stream.audio_stream
        if getattr(src, "pcm_format", None) is not None and self._pcm_chunks_fit(src):
            # raw integer PCM (e.g. match._WavFileStreamWrapper on a 16/32-bit WAV): the frames cross PCIe as they are
            # and are widened -- and, at another rate, resampled chunk by chunk -- on the device (rows N1, N2).  A source
            # rate that does not cut into whole device chunks (e.g. 44 101 Hz) takes the float read() path below, where
            # the wrapper converts and resamples on the host as the reference does (match.py:393-427)
            return self._find_clip_in_pcm(src, peak_times, on_pattern_detected)
        halo = np.zeros(0, dtype=np.float32)          # tail of the previous batch (look-back)
        chunk_index = 0
        eof = False
        cap = self._max_halo + self._stream_read_chunks * C_
        with torch.cuda.device(self._device):
            if self._dev_buf is None:
                self._dev_buf = torch.empty(cap, dtype=torch.float32, device=f"cuda:{self._device}")
                self._pinned = torch.empty(cap, dtype=torch.float32).pin_memory()
            while not eof:
                parts: list[NDArray[np.float32]] = []
                while len(parts) < self._stream_read_chunks:
                    data = src.read(self._chunk_size)
                    # a raw / unbuffered source may return fewer bytes than asked for before its end: keep reading until
                    # the chunk is full or read() returns b'' (the reference, :296-299, would scan each short read as a
                    # chunk of its own with misplaced timestamps; buffered sources never get here)
                    while data and len(data) < self._chunk_size:
                        more = src.read(self._chunk_size - len(data))
                        if not more:
                            break
                        data += more
                    if not data:
                        eof = True
                        break
                    chunk = np.frombuffer(data[:len(data) - len(data) % 4], dtype="float32")
                    total_time += len(chunk) / sr                           # reference :301
                    parts.append(chunk)
                    if len(chunk) != C_:
                        eof = True                                          # the stream's final, shorter chunk
                        break
                if not parts:
                    break
                new = np.concatenate(parts) if len(parts) > 1 else parts[0]
                n_halo = halo.size
                n_tot = n_halo + new.size
                host = self._pinned[:n_tot].numpy()
                host[:n_halo] = halo
                host[n_halo:] = new
                self._dev_buf[:n_tot].copy_(self._pinned[:n_tot], non_blocking=True)
                c0, c1 = chunk_index, chunk_index + len(parts)
                rec, _ = self._scan_batch(self._dev_buf.data_ptr(), c0 * C_ - n_halo, n_tot, c0, c1, False)
                self._emit_batch(rec, self._timestamps(rec), peak_times, events, on_pattern_detected)
                keep = min(self._max_halo, n_tot)
                halo = host[n_tot - keep:n_tot].copy()
                chunk_index = c1
        return peak_times, total_time
