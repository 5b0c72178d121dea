"""CPU oracle for the two-step detection hot path -- TEST INFRASTRUCTURE ONLY.

A numpy/scipy + C (oracle/native_oracle.c) restatement of what the reference
computes per (chunk x pattern) unit, written as plain functions that return a
full trace (gain, maxima, candidate peaks, every verification score) so the
CUDA path can be compared stage by stage.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

File:line citations are into /root/reference/audio_pattern_detector/
(apd.py = audio_pattern_detector.py, du.py = detection_utils.py,
au.py = audio_utils.py) and native-helper/src/lib.rs.

Third-party arithmetic not in the reference tree: ``fft-correlation==0.0.5``
(pyproject.toml:9), a Rust wheel providing ``fft_correlate_1d(a, b,
mode='full')``.  Its published contract is scipy.signal.correlate(..., 'full')
semantics returning float32; that is what :func:`correlate_full` restates
(float64 FFT, rounded once to float32).  Exact-sample parity at that boundary
is pinned by the reference's end-to-end goldens only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Optional

import numpy as np
import scipy.fft as sfft

from . import native

TARGET_LUFS = -16.0
DEFAULT_HEIGHT = 0.25                # apd.py:520
SHORT_CLIP_SECONDS = 0.5             # apd.py:36
MSE_LIMIT = 0.02                     # apd.py:793
PEARSON_MIN = 0.90                   # apd.py:794
TONE_DEFAULTS = {                    # apd.py:698-705
    "minimum_band_purity": 0.95,
    "minimum_active_frame_ratio": 0.80,
    "minimum_longest_active_run": 9,
    "minimum_active_frame_mean_purity": 0.92,
    "maximum_min_flank_purity": 0.25,
    "maximum_max_flank_purity": 0.65,
}


# --------------------------------------------------------------------------- #
# Step 1 pieces
# --------------------------------------------------------------------------- #
def correlate_full(a: np.ndarray, b: np.ndarray, precision: str = "f64") -> np.ndarray:
    """out[k] = sum_n a[n + k - (len(b)-1)] * b[n], k in [0, len(a)+len(b)-1); float32.

    Restates the call ``fft_correlate_1d(section, clip, mode='full')`` at
    apd.py:376,491 (scipy 'full' convention: peak index = match_start + L - 1,
    consistent with apd.py:650)."""
    a = np.asarray(a)
    b = np.asarray(b)
    n = a.size + b.size - 1
    if a.size == 0 or b.size == 0:
        return np.zeros(max(n, 0), dtype=np.float32)
    dt = np.float64 if precision == "f64" else np.float32
    nfft = sfft.next_fast_len(n, real=True)
    fa = sfft.rfft(a.astype(dt, copy=False), nfft)
    fb = sfft.rfft(b[::-1].astype(dt, copy=False), nfft)
    fa *= fb
    out = sfft.irfft(fa, nfft)[:n]
    return out.astype(np.float32)


def section_block_size(n_samples: int, sr: int) -> float:
    """apd.py:416-417 / :169: block = T if T < 0.5 else 0.4."""
    t = n_samples / sr
    return t if t < 0.5 else 0.4


def normalize_section(section: np.ndarray, sr: int) -> tuple[np.ndarray, float]:
    """apd.py:414-420 then the NaN scrub of apd.py:489-490.  Returns (f32 section, lufs)."""
    lufs = native.integrated_loudness(section, sr, section_block_size(section.size, sr))
    out = native.loudness_normalize(section, lufs, TARGET_LUFS)
    np.nan_to_num(out, copy=False, nan=0.0)
    return out, lufs


# --------------------------------------------------------------------------- #
# Pattern-side precompute (apd.py:155-221)
# --------------------------------------------------------------------------- #
@dataclass
class ClipState:
    name: str
    clip: np.ndarray                 # loudness-normalised f32
    length: int
    sliding_window: int
    self_corr: np.ndarray            # |corr(clip, clip)| / max, f32, len 2L-1
    self_max: np.float32
    tone_hz: Optional[float] = None
    tone_thresholds: dict = field(default_factory=dict)
    pearson_cache: Optional[list] = None


def prepare_clip(name: str, audio: np.ndarray, sr: int, strategy: Optional[str] = None,
                 strategy_params: Optional[dict] = None, precision: str = "f64") -> ClipState:
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    n = audio.size
    seconds = n / sr
    lufs = native.integrated_loudness(audio, sr, seconds if seconds < SHORT_CLIP_SECONDS else 0.4)
    clip = native.loudness_normalize(audio, lufs, TARGET_LUFS)          # apd.py:169-171
    cc = np.abs(correlate_full(clip, clip, precision))                   # apd.py:373-383
    smax = np.max(cc)
    cc = cc / smax
    st = ClipState(name=name, clip=clip, length=n, sliding_window=math.ceil(seconds),
                   self_corr=cc.astype(np.float32), self_max=np.float32(smax))
    if strategy == "marker_tone":                                         # apd.py:214-221
        params = dict(strategy_params or {})
        f0 = params.get("dominant_frequency_hz")
        if f0 is None:
            f0 = pure_tone_frequency(clip, sr)
        if f0 is not None:
            st.tone_hz = float(f0)
        ver = params.get("verification", {})
        st.tone_thresholds = dict(ver) if isinstance(ver, dict) else {}
    return st


def pure_tone_frequency(audio: np.ndarray, sr: int) -> Optional[float]:
    """du.py:19-38 (init-time only)."""
    mag = np.abs(np.fft.rfft(audio))
    freqs = np.fft.rfftfreq(len(audio), d=1 / sr)
    k = int(np.argmax(mag))
    if mag[k] == 0.0:
        return None
    peaks, _ = native.find_peaks(np.ascontiguousarray(mag / mag[k]), prominence=0.05)
    f = float(freqs[k])
    if len(peaks) == 1 and math.isclose(freqs[peaks[0]], f, rel_tol=0.01):
        return f
    return None


# --------------------------------------------------------------------------- #
# Step 2: normal / short-clip verifier (apd.py:752-902)
# --------------------------------------------------------------------------- #
def centered_slice(arr: np.ndarray, width: int, mid: int) -> np.ndarray:
    """au.py:177-191: arr[mid - floor(w/2) : mid + ceil(w/2)], zero padded, exact width."""
    lo = int(mid - math.floor(width / 2))
    hi = int(mid + math.ceil(width / 2))
    out = np.zeros(hi - lo, dtype=arr.dtype)
    s0, s1 = max(lo, 0), min(hi, arr.size)
    if s1 > s0:
        out[s0 - lo:s1 - lo] = arr[s0:s1]
    return out


def pearson_windows(short: bool) -> tuple[list[tuple[int, int, int]], int]:
    """apd.py:808-820."""
    base = 101
    if short:
        return [(0, 10, round(base * 10 / 2))], 0
    return [(0, 5, round(base * 5 / 2)), (4, 6, base), (5, 10, round(base * 5 / 2))], 1


def verify_normal(st: ClipState, corr: np.ndarray, peak: int, sr: int) -> dict:
    """Returns a score record; 'accept' is the reference's decision."""
    cc = st.self_corr
    width = cc.size
    sl = centered_slice(corr, width, peak)                               # apd.py:622
    sl = sl / np.max(sl)                                                  # apd.py:623 (f32)
    short = st.length / sr < SHORT_CLIP_SECONDS                           # apd.py:628
    ps = width // 10                                                      # apd.py:777
    parts = np.array([np.mean((cc[i * ps:(i + 1) * ps] - sl[i * ps:(i + 1) * ps]) ** 2)
                      for i in range(10)], dtype=np.float32)              # apd.py:779-783
    middle = np.mean(parts[4:6])
    whole = np.mean(parts)
    sim = whole if short else min(whole, middle)                          # apd.py:788-791
    rec: dict[str, Any] = {"kind": "short" if short else "normal", "peak": int(peak),
                           "similarity_whole": float(whole), "similarity_middle": float(middle),
                           "similarity": float(sim), "pearson": None, "accept": False}
    if sim > MSE_LIMIT:                                                   # apd.py:796 (f32 compare)
        return rec
    wins, center = pearson_windows(short)
    if st.pearson_cache is None:                                          # apd.py:822-829
        st.pearson_cache = [native.resample_preserve_maxima(
            np.ascontiguousarray(cc[round(width * wl / 10):round(width * wr / 10)]), ds)
            for wl, wr, ds in wins]
    rs = []
    for wi, (wl, wr, ds) in enumerate(wins):                              # apd.py:835-840
        lo, hi = round(width * wl / 10), round(width * wr / 10)
        d = native.resample_preserve_maxima(np.ascontiguousarray(sl[lo:hi]), ds)
        rs.append(native.pearson_correlation(st.pearson_cache[wi], d))
    rec["pearson"] = rs
    rec["pearson_center"] = rs[center]
    rec["accept"] = bool(rs[center] >= PEARSON_MIN)                       # apd.py:846,897
    return rec


# --------------------------------------------------------------------------- #
# Step 2: marker-tone verifier (apd.py:642-750, du.py:41-142)
# --------------------------------------------------------------------------- #
def padded_segment(x: np.ndarray, start: int, length: int) -> np.ndarray:
    """du.py:128-142."""
    out = np.zeros(length, dtype=np.float32)
    s0, s1 = max(start, 0), min(start + length, x.size)
    if s1 > s0:
        out[s0 - start:s1 - start] = x[s0:s1]
    return out


def tone_metrics(seg: np.ndarray, sr: int, f0: float) -> dict:
    """du.py:41-125; all float64 (np.hanning is f64)."""
    z = {"detected_frequency": 0.0, "overall_band_purity": 0.0, "active_frame_ratio": 0.0,
         "longest_active_run": 0, "active_frame_mean_purity": 0.0}
    n = seg.size
    if n == 0:
        return z
    band = max(40.0, f0 * 0.08)
    lock = max(20.0, f0 * 0.04)
    spec = np.abs(np.fft.rfft(seg * np.hanning(n)))
    freqs = np.fft.rfftfreq(n, d=1 / sr)
    z["detected_frequency"] = float(freqs[int(np.argmax(spec))])
    total = float(np.sum(spec ** 2))
    if total == 0.0:
        return z
    z["overall_band_purity"] = float(np.sum(spec[np.abs(freqs - f0) <= band] ** 2)) / total
    wl = max(int(round(0.025 * sr)), 32)
    hop = max(wl // 2, 1)
    win = np.hanning(wl)
    ffreqs = np.fft.rfftfreq(wl, d=1 / sr)
    fband = np.abs(ffreqs - f0) <= band
    frames = active = run = best = 0
    purities = []
    for s in range(0, n - wl, hop):                                       # du.py:87
        fs = np.abs(np.fft.rfft(seg[s:s + wl] * win))
        e = float(np.sum(fs ** 2))
        if e == 0.0:
            run = 0
            continue
        frames += 1
        fdom = float(ffreqs[int(np.argmax(fs))])
        pur = float(np.sum(fs[fband] ** 2)) / e
        if math.isclose(fdom, f0, abs_tol=lock) and pur >= 0.55:          # du.py:102-105
            active += 1
            run += 1
            best = max(best, run)
            purities.append(pur)
        else:
            run = 0
    z["active_frame_ratio"] = active / frames if frames else 0.0
    z["longest_active_run"] = best
    z["active_frame_mean_purity"] = float(np.mean(purities)) if purities else 0.0
    return z


def verify_tone(st: ClipState, section: np.ndarray, peak: int, sr: int) -> dict:
    L = st.length
    f0 = st.tone_hz
    assert f0 is not None
    start = peak - L + 1                                                  # apd.py:650
    m = tone_metrics(padded_segment(section, start, L), sr, f0)
    left = tone_metrics(padded_segment(section, start - L, L), sr, f0)
    right = tone_metrics(padded_segment(section, start + L, L), sr, f0)
    th = {**TONE_DEFAULTS, **st.tone_thresholds}
    lo = min(left["overall_band_purity"], right["overall_band_purity"])
    hi = max(left["overall_band_purity"], right["overall_band_purity"])
    ok = math.isclose(m["detected_frequency"], f0, rel_tol=0.05)          # apd.py:707
    ok = ok and (m["overall_band_purity"] >= float(th["minimum_band_purity"])
                 and m["active_frame_ratio"] >= float(th["minimum_active_frame_ratio"])
                 and m["longest_active_run"] >= int(th["minimum_longest_active_run"])
                 and m["active_frame_mean_purity"] >= float(th["minimum_active_frame_mean_purity"])
                 and lo <= float(th["maximum_min_flank_purity"])
                 and hi <= float(th["maximum_max_flank_purity"]))         # apd.py:717-724
    return {"kind": "tone", "peak": int(peak), "match": m, "left": left, "right": right,
            "accept": bool(ok)}


# --------------------------------------------------------------------------- #
# One (chunk x pattern) unit and the chunk loop
# --------------------------------------------------------------------------- #
def process_unit(st: ClipState, section_raw: np.ndarray, sr: int, height: Optional[float] = None,
                 precision: str = "f64", normalized: Optional[tuple] = None, keep_corr: bool = False) -> dict:
    """Everything apd.py:414-420 + :466-587 does for one section/clip pair."""
    if normalized is None:
        section, lufs = normalize_section(section_raw, sr)
    else:
        section, lufs = normalized
    corr = np.abs(correlate_full(section, st.clip, precision))            # apd.py:491
    absmax = np.max(corr) if corr.size else np.float32(0)
    max_choose = max(st.self_max, absmax)                                 # apd.py:493
    corr /= max_choose                                                    # apd.py:494
    h = DEFAULT_HEIGHT if height is None else height
    peaks, _ = native.find_peaks(corr, height=h, distance=st.length)      # apd.py:516-522
    half = st.self_corr.size // 2
    cands = []
    accepted = []
    for pk in peaks.tolist():
        if pk + half > corr.size + 5 or pk - half < -5:                   # apd.py:534-546
            cands.append({"kind": "skipped", "peak": int(pk), "accept": False})
            continue
        if st.tone_hz is not None:                                        # apd.py:605-620
            rec = verify_tone(st, section, pk, sr)
        else:
            rec = verify_normal(st, corr, pk, sr)
        rec["height"] = float(corr[pk])
        cands.append(rec)
        if rec["accept"]:
            accepted.append(int(pk))
    out = {"lufs": lufs, "absmax": float(absmax), "max_choose": float(max_choose),
           "n_out": int(corr.size), "peaks": [int(p) for p in peaks], "candidates": cands,
           "accepted": accepted}
    if keep_corr:
        out["corr"] = corr
    return out


def peak_to_timestamp(peak: int, sr: int, subtract_seconds: int, index: int, spc: int, L: int) -> float:
    """apd.py:585 then :440-451, in the reference's order of float operations."""
    t = peak / sr
    t = t - subtract_seconds
    t = t + (index * spc)
    t = t - (L / sr)
    return t if t >= 0 else 0


class OracleDetector:
    """Chunk loop of apd.py:248-371 over an in-memory float32 stream."""

    def __init__(self, clips: list[dict], sr: int = 8000, seconds_per_chunk: Optional[int] = 60,
                 height_min: Optional[float] = None, precision: str = "f64", keep_corr: bool = False) -> None:
        self.keep_corr = keep_corr
        self.sr = sr
        self.height_min = height_min
        self.precision = precision
        self.states = [prepare_clip(c["name"], c["audio"], sr, c.get("strategy"),
                                    c.get("strategy_params"), precision) for c in clips]
        maxlen = max((s.length for s in self.states), default=0)
        if seconds_per_chunk is None or seconds_per_chunk < 1:            # apd.py:117-119
            seconds_per_chunk = math.ceil(maxlen / sr) * 2
        for s in self.states:                                             # apd.py:125-136
            if seconds_per_chunk < 2 * s.sliding_window:
                raise ValueError(f"seconds_per_chunk {seconds_per_chunk} is too small for clip '{s.name}'")
        self.spc = seconds_per_chunk
        self.chunk_samples = int(seconds_per_chunk * sr)

    def run(self, audio: np.ndarray, on_unit: Optional[Callable[[int, ClipState, dict], None]] = None,
            chunk_range: Optional[tuple[int, int]] = None) -> tuple[dict, list, float]:
        """Returns (peak_times by clip, ordered events [(t, name, chunk, peak)], total seconds).

        chunk_range=(a, b) processes chunks a..b-1 only (each still sees its look-back halo);
        used to shard the CPU baseline across processes."""
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        C, sr = self.chunk_samples, self.sr
        n_chunks = (audio.size + C - 1) // C
        a, b = (0, n_chunks) if chunk_range is None else chunk_range
        times: dict[str, list] = {s.name: [] for s in self.states}
        events = []
        total = 0.0
        for i in range(n_chunks):                                         # apd.py:301 (running float sum)
            total += (min((i + 1) * C, audio.size) - i * C) / sr
        for i in range(a, min(b, n_chunks)):
            lo, hi = i * C, min((i + 1) * C, audio.size)
            norm_cache: dict[int, tuple] = {}
            chunk_events = []
            for st in self.states:
                if i > 0:                                                 # apd.py:406-412
                    # previous_chunk[int(-sw*sr):] -- previous chunk is always a full chunk
                    sub = st.sliding_window
                    start = lo - min(int(sub * sr), C)
                else:
                    sub, start = 0, lo
                if start not in norm_cache:
                    norm_cache[start] = normalize_section(audio[start:hi], sr)
                tr = process_unit(st, audio[start:hi], sr, self.height_min, self.precision,
                                  normalized=norm_cache[start], keep_corr=self.keep_corr)
                tr["section_start"] = start
                if on_unit is not None:
                    on_unit(i, st, tr)
                for pk in tr["accepted"]:
                    t = peak_to_timestamp(pk, sr, sub, i, self.spc, st.length)
                    times[st.name].append(t)
                    chunk_events.append((t, st.name, i, pk))
            chunk_events.sort(key=lambda e: e[0])                         # apd.py:324-327 (stable)
            events.extend(chunk_events)
        return times, events, total
