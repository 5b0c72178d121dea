"""Helpers to replay tests/golden/ (written by oracle/make_golden.py)."""
from __future__ import annotations

import functools
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def fixtures():
    return np.load(os.path.join(GOLDEN, "fixtures.npz"))


@functools.lru_cache(maxsize=None)
def load_json(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def fixture_audio(run: dict) -> np.ndarray:
    pcm = fixtures()["wav:" + run["wav"]]
    return pcm.astype(np.float32) / 32768.0


def fixture_clips(run: dict) -> list[dict]:
    fx = fixtures()
    return [{"name": m["name"], "audio": fx[f"clip{run['sr']}:{m['name']}"], "strategy": m["strategy"],
             "strategy_params": m["strategy_params"]} for m in run["clips"]]


def synthetic_inputs(run: dict):
    """Regenerate a synthetic golden case's inputs and check the recorded checksum."""
    import hashlib
    case = run["case"]
    if case.get("kind") == "beeps_in_silence":
        sr = case["sr"]
        t = np.linspace(0, 0.23, int(sr * 0.23), endpoint=False)
        beep = np.sin(2 * np.pi * 1000.0 * t).astype(np.float32)
        audio = np.zeros(int(sr * 31.0), dtype=np.float32)
        for s in run["starts"]:
            i = int(s * sr)
            audio[i:i + beep.size] = beep
        clips = [{"name": "test_beep", "audio": beep, "strategy": None, "strategy_params": {}}]
    else:
        from audio_pattern_detector_b200 import workloads as W
        clips = W.make_patterns(case["n_patterns"], case["sr"], case["pattern_seed"], case["min_s"], case["max_s"])
        spc = case["spc"]
        if spc is None:
            spc = int(np.ceil(max(p["audio"].size for p in clips) / case["sr"])) * 2
        audio, _ = W.make_stream(case["seconds"], clips, case["sr"], case["stream_seed"], case["plants"], spc)
    got = hashlib.sha256(np.ascontiguousarray(audio).tobytes()).hexdigest()[:16]
    assert got == run["audio_sha"], "synthetic generator drifted from the committed golden"
    return clips, audio


def rel_close(a: float, b: float, rel: float = 1e-4, abs_: float = 1e-9) -> bool:
    return abs(a - b) <= max(rel * max(abs(a), abs(b)), abs_)


def peaks_match(got: list, ref: list, corr, L: int, tol: float = 3e-5) -> tuple[bool, bool]:
    """(ok, exact).  ok also when the lists differ only by rounding-level ties: a pair of
    different indices inside one suppression neighbourhood (< L apart) whose normalised
    correlation heights agree to `tol` relative -- float32 FFT rounding decides those, and
    neither the reference's stand-in FFT nor ours is authoritative there."""
    if list(got) == list(ref):
        return True, True
    if len(got) != len(ref):
        return False, False
    for g, r in zip(got, ref):
        if g == r:
            continue
        if abs(g - r) >= L:
            return False, False
        a, b = float(corr[g]), float(corr[r])
        if abs(a - b) > tol * max(abs(a), abs(b)):
            return False, False
    return True, False
