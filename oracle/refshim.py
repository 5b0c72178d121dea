"""Import the UNMODIFIED reference Python package from /root/reference -- golden generation only.

The reference cannot run as shipped in this image: its Rust extension
(``audio_pattern_detector._native``, native-helper/) cannot be compiled (no
cargo/rustc) and two wheels are absent (``fft-correlation==0.0.5``,
``andrew-utils``; pyproject.toml:9,11).  This module installs three stand-ins
in ``sys.modules`` and then imports the reference's own pure-Python
orchestration, which is what oracle/make_golden.py runs to produce
tests/golden/.  It only works in the build container (/root/reference does not
exist on the GPU box) and nothing under tests/ -m gpu, smoke() or bench.py
uses it.
"""
from __future__ import annotations

import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def seconds_to_time(seconds: float, include_decimals: bool = True) -> str:
    """Stand-in for andrew_utils.seconds_to_time: HH:MM:SS[.mmm] (README.md:89-93)."""
    ms_total = int(round(float(seconds) * 1000)) if include_decimals else int(seconds) * 1000
    h, rem = divmod(ms_total, 3600_000)
    m, rem = divmod(rem, 60_000)
    s, ms = divmod(rem, 1000)
    return f"{h:02d}:{m:02d}:{s:02d}.{ms:03d}" if include_decimals else f"{h:02d}:{m:02d}:{s:02d}"


def _fft_correlate_1d(a, b, mode: str = "full"):
    """Stand-in for fft_correlation.fft_correlate_1d: scipy's FFT correlation, float32 out."""
    from scipy.signal import correlate
    a32 = np.asarray(a, dtype=np.float32)
    b32 = np.asarray(b, dtype=np.float32)
    return correlate(a32, b32, mode=mode, method="fft").astype(np.float32)


def _resample(data, num_samples: int):
    """Stand-in for _native.resample (lib.rs:235-275): scipy.signal.resample spectrum slicing."""
    x = np.asarray(data, dtype=np.float64)
    n, m = x.size, int(num_samples)
    if n == 0 or m == 0:
        return np.zeros(m, dtype=np.float32)
    if n == m:
        return x.astype(np.float32)
    X = np.fft.fft(x)
    nc = min(n, m)
    pos, neg = (nc + 1) // 2, (nc - 1) // 2
    Y = np.zeros(m, dtype=np.complex128)
    Y[:pos] = X[:pos]
    if neg:
        Y[m - neg:] = X[n - neg:]
    return (np.fft.ifft(Y).real * (m / n)).astype(np.float32)


def install() -> types.ModuleType:
    """Install stand-ins and return the imported reference package."""
    from . import native

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    m = types.ModuleType("fft_correlation")
    m.fft_correlate_1d = _fft_correlate_1d
    sys.modules["fft_correlation"] = m

    m = types.ModuleType("andrew_utils")
    m.seconds_to_time = seconds_to_time
    sys.modules["andrew_utils"] = m

    m = types.ModuleType("audio_pattern_detector._native")
    for name in ("find_peaks", "resample_preserve_maxima", "integrated_loudness",
                 "loudness_normalize", "pearson_correlation"):
        setattr(m, name, getattr(native, name))
    m.resample = _resample
    sys.modules["audio_pattern_detector._native"] = m

    import audio_pattern_detector  # the reference package
    audio_pattern_detector._native = m
    return audio_pattern_detector
