"""Host-side pieces of the reference's detection_utils.py that are NOT on the GPU path.

The per-candidate tone metrics (detection_utils.py:41-125) run in csrc/verify.cu; what is
left here is the result type and the init-time fallback ``get_pure_tone_frequency``
(detection_utils.py:19-38), used only when a marker_tone clip declares no frequency.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from numpy.typing import NDArray


@dataclass(frozen=True)
class PureToneMetrics:                                   # detection_utils.py:8-16
    detected_frequency: float
    overall_band_purity: float
    active_frame_ratio: float
    longest_active_run: int
    active_frame_mean_purity: float


def _prominent_peaks(x: NDArray[np.floating], min_prominence: float) -> list[int]:
    """Strict local maxima (plateau midpoints) whose prominence >= min_prominence
    (reference native-helper/src/lib.rs:404-428, 496-523)."""
    n = len(x)
    out: list[int] = []
    i = 1
    while i < n - 1:
        if x[i - 1] < x[i]:
            left = i
            while i + 1 < n and x[i] == x[i + 1]:
                i += 1
            if i + 1 < n and x[i] > x[i + 1]:
                out.append((left + i) // 2)
        i += 1
    keep = []
    for p in out:
        v = x[p]
        lmin = v
        for j in range(p - 1, -1, -1):
            lmin = min(lmin, x[j])
            if x[j] > v:
                break
        rmin = v
        for j in range(p + 1, n):
            rmin = min(rmin, x[j])
            if x[j] > v:
                break
        if v - max(lmin, rmin) >= min_prominence:
            keep.append(p)
    return keep


def get_pure_tone_frequency(audio_data: NDArray[np.float32], sample_rate: int) -> float | None:
    mag = np.abs(np.fft.rfft(audio_data))
    freqs = np.fft.rfftfreq(len(audio_data), d=1 / sample_rate)
    k = int(np.argmax(mag))
    if mag[k] == 0.0:
        return None
    peaks = _prominent_peaks((mag / mag[k]).astype(np.float32), np.float32(0.05))
    f = float(freqs[k])
    if len(peaks) == 1 and math.isclose(freqs[peaks[0]], f, rel_tol=0.01):
        return f
    return None


def max_distance(sorted_data: list[float]) -> float:     # detection_utils.py:145-151
    return max((b - a for a, b in zip(sorted_data, sorted_data[1:])), default=0)
