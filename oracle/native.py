"""ctypes front-end for oracle/native_oracle.c -- TEST INFRASTRUCTURE ONLY.

Mirrors the call surface of the reference's PyO3 module
``audio_pattern_detector._native`` (native-helper/src/python.rs:79-181) so the
oracle detector and the reference-import shim can both use it.  Never imported
by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Any

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_native.so")


def build(force: bool = False) -> str:
    """Compile native_oracle.c with gcc (seconds).  Returns the .so path."""
    src = os.path.join(_HERE, "native_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "liboracle_native.so"], check=True)
    return _SO


_lib: Any = None


def lib() -> Any:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_k_weighting.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        L.oracle_k_weighting.restype = None
        L.oracle_integrated_loudness.argtypes = [f32p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_double]
        L.oracle_integrated_loudness.restype = ctypes.c_double
        L.oracle_loudness_normalize.argtypes = [f32p, ctypes.c_size_t, ctypes.c_double, ctypes.c_double, f32p]
        L.oracle_loudness_normalize.restype = None
        L.oracle_resample_preserve_maxima.argtypes = [f32p, ctypes.c_size_t, ctypes.c_size_t, f32p]
        L.oracle_resample_preserve_maxima.restype = ctypes.c_size_t
        L.oracle_local_maxima.argtypes = [f32p, ctypes.c_int64, i64p, ctypes.c_int64]
        L.oracle_local_maxima.restype = ctypes.c_int64
        L.oracle_find_peaks.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                        ctypes.c_int, ctypes.c_int64, i64p]
        L.oracle_find_peaks.restype = ctypes.c_int64
        L.oracle_pearson.argtypes = [f32p, f32p, ctypes.c_size_t]
        L.oracle_pearson.restype = ctypes.c_double
        _lib = L
    return _lib


def _f32(data: Any) -> np.ndarray:
    """python.rs:16-40: contiguous 1-D f32 borrowed, f64 converted, else TypeError."""
    if not isinstance(data, np.ndarray) or data.ndim != 1:
        raise TypeError("data must be a contiguous 1D numpy.ndarray with dtype float32 or float64")
    if data.dtype not in (np.float32, np.float64):
        raise TypeError("data must be a contiguous 1D numpy.ndarray with dtype float32 or float64")
    if not data.flags["C_CONTIGUOUS"]:
        raise TypeError("data must be a contiguous 1D numpy array")
    return data if data.dtype == np.float32 else data.astype(np.float32)


def _p(a: np.ndarray) -> Any:
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def k_weighting_coefficients(rate: float) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    c = (ctypes.c_double * 12)()
    lib().oracle_k_weighting(float(rate), c)
    v = np.array(list(c), dtype=np.float64)
    return v[0:3], v[3:6], v[6:9], v[9:12]


def integrated_loudness(data: Any, sample_rate: int, block_size: float = 0.4) -> float:
    x = _f32(data)
    return float(lib().oracle_integrated_loudness(_p(x), x.size, int(sample_rate), float(block_size)))


def loudness_normalize(data: Any, current_lufs: float, target_lufs: float) -> np.ndarray:
    x = _f32(data)
    out = np.empty(x.size, dtype=np.float32)
    lib().oracle_loudness_normalize(_p(x), x.size, float(current_lufs), float(target_lufs), _p(out))
    return out


def resample(data: Any, num_samples: int) -> np.ndarray:
    """FFT resampling, restating resample_1d (native-helper/src/lib.rs:235-275; binding python.rs:106-116):
    full complex float64 FFT, keep the (N+1)//2 lowest positive and (N-1)//2 lowest negative bins of
    N = min(n, num_samples) -- an even N loses its Nyquist bin, unlike scipy.signal.resample -- un-normalised
    inverse FFT of length num_samples, times 1/n, rounded to float32."""
    x = _f32(data)
    n, m = x.size, int(num_samples)
    if n == 0 or m == 0:                                 # lib.rs:237-239
        return np.zeros(m, dtype=np.float32)
    if n == m:                                           # lib.rs:240-242
        return x.copy()
    spectrum = np.fft.fft(x.astype(np.float64))          # lib.rs:247-251
    nc = min(n, m)
    pos, neg = (nc + 1) // 2, (nc - 1) // 2              # lib.rs:258-260
    new = np.zeros(m, dtype=np.complex128)
    new[:pos] = spectrum[:pos]
    if neg > 0:
        new[m - neg:] = spectrum[n - neg:]
    y = np.fft.ifft(new, norm="forward")                 # un-normalised inverse, lib.rs:268-269
    return (y.real * (1.0 / n)).astype(np.float32)       # lib.rs:273-274


def resample_preserve_maxima(data: Any, num_samples: int) -> np.ndarray:
    if num_samples == 0:
        raise ValueError("num_samples must be greater than 0")
    x = _f32(data)
    out = np.empty(num_samples, dtype=np.float32)
    got = lib().oracle_resample_preserve_maxima(_p(x), x.size, int(num_samples), _p(out))
    if got != num_samples:
        raise ValueError(f"downsampled curve length {got} not equal to num_samples {num_samples}")
    return out


def local_maxima(data: Any) -> np.ndarray:
    x = _f32(data)
    out = np.empty(x.size // 2 + 1, dtype=np.int64)
    m = lib().oracle_local_maxima(_p(x), x.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), out.size)
    return out[:m].copy()


def find_peaks(data: Any, *, height: float | None = None, distance: int | None = None,
               prominence: float | None = None) -> tuple[np.ndarray, dict]:
    """height/distance in C; prominence (init-time only in the reference, du.py:32) via scipy."""
    x = _f32(data)
    out = np.empty(x.size // 2 + 1, dtype=np.int64)
    m = lib().oracle_find_peaks(_p(x), x.size, height is not None,
                                np.float32(0.0 if height is None else height),
                                distance is not None, 0 if distance is None else int(distance),
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if m < 0:
        raise MemoryError("oracle_find_peaks")
    peaks = out[:m].copy()
    if prominence is not None and peaks.size:
        from scipy.signal import peak_prominences
        prom = peak_prominences(x, peaks)[0]
        peaks = peaks[prom >= np.float32(prominence)]
    return peaks, {}


def pearson_correlation(x: Any, y: Any) -> float:
    a, b = _f32(x), _f32(y)
    if a.size != b.size:
        raise ValueError("arrays must have the same length")
    return float(lib().oracle_pearson(_p(a), _p(b), a.size))
