"""ctypes binding of include/apd_b200.h (the C ABI over the CUDA kernels).

There is no CPU fallback: if the shared library is missing this module raises at
import of :func:`lib`, and every entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapd_b200.so")

APD_OK = 0
ERR_NAMES = {1: "invalid argument", 2: "CUDA error", 3: "unsupported configuration", 4: "workspace overflow"}
STRATEGY_NORMAL, STRATEGY_MARKER_TONE = 0, 1
FLAG_ACCEPT, FLAG_SKIPPED, KIND_SHIFT = 1, 2, 2
KINDS = ("normal", "short", "tone")


class ClipDesc(C.Structure):
    _fields_ = [("samples", C.POINTER(C.c_float)), ("length", C.c_int32), ("strategy", C.c_int32),
                ("tone_hz", C.c_double), ("minimum_band_purity", C.c_double),
                ("minimum_active_frame_ratio", C.c_double), ("minimum_longest_active_run", C.c_double),
                ("minimum_active_frame_mean_purity", C.c_double), ("maximum_min_flank_purity", C.c_double),
                ("maximum_max_flank_purity", C.c_double)]


class Candidate(C.Structure):
    _fields_ = [("chunk", C.c_int32), ("clip", C.c_int32), ("peak", C.c_int32), ("flags", C.c_int32),
                ("height", C.c_float), ("similarity_whole", C.c_float), ("similarity_middle", C.c_float),
                ("reserved", C.c_float), ("pearson", C.c_double * 3), ("tone", (C.c_double * 5) * 3)]


class UnitTrace(C.Structure):
    _fields_ = [("absmax", C.c_float), ("max_choose", C.c_float), ("n_out", C.c_int32), ("n_peaks", C.c_int32)]


# name -> (restype, argtypes); must list every APD_API symbol of include/apd_b200.h
PROTOTYPES: dict[str, tuple[Any, list[Any]]] = {
    "apd_last_error": (C.c_char_p, []),
    "apd_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_float, C.c_int,
                             C.POINTER(ClipDesc), C.c_int]),
    "apd_destroy": (C.c_int, [C.c_void_p]),
    "apd_clip_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "apd_clip_normalized": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "apd_clip_self_correlation": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "apd_scan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                           C.POINTER(Candidate), C.c_int32, C.POINTER(C.c_int32), C.POINTER(UnitTrace),
                           C.POINTER(C.c_double), C.c_void_p]),
    "apd_stage_loudness": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "apd_stage_forward_fft": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_stage_correlate_max": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_stage_peaks_verify": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_stage_unit_correlation": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int32,
                                             C.POINTER(C.c_int32), C.c_void_p]),
    "apd_verify_tone": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                  C.POINTER(C.c_int32), C.c_void_p]),
    "apd_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "apd_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "apd_pcm_to_float": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "apd_resample_workspace_bytes": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    "apd_resample": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                               C.c_void_p, C.c_int64, C.c_void_p]),
    "apd_launch_count": (C.c_int64, [C.c_void_p]),
    "apd_work_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "apd_unit_n_out": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int32)]),
}

_lib: Any = None


def lib() -> Any:
    """Load libapd_b200.so once.  Raises (never falls back) if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -m audio_pattern_detector_b200.build); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class ApdError(RuntimeError):
    code = 0            # status code of include/apd_b200.h (4: a workspace list overflowed)


def check(rc: int, what: str) -> None:
    if rc != APD_OK:
        msg = lib().apd_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        err = ApdError(f"{what}: {ERR_NAMES.get(rc, rc)}: {msg}")
        err.code = rc
        raise err
