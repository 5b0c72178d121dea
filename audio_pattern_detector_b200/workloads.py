"""Deterministic synthetic "radio" workloads.

Used by bench.py, the parity tests and oracle/make_golden.py so that all three
see the same inputs.  Generation is numpy ``RandomState`` (a frozen stream, so
the same seed gives the same samples on every box) and is not part of the
timed region anywhere.

A stream is a band-limited noise bed (white N(0, 0.1^2) through a one-pole
low-pass) with pattern clips planted at seeded offsets; some offsets are forced
to straddle chunk boundaries.  Patterns are a mix of noise-like "jingles",
chirps (some shorter than the 0.5 s short-clip threshold) and marker-tone sines
declared the way an ``.apd.toml`` ``source = "sine"`` clip is
(reference pattern_config.py:106-108: float32 phase).
"""
from __future__ import annotations

import math
from typing import Any, Optional

import numpy as np


def _one_pole(x: np.ndarray, a: float) -> np.ndarray:
    from scipy.signal import lfilter
    return lfilter([1.0 - a], [1.0, -a], x).astype(np.float32)


def make_patterns(n: int, sr: int = 8000, seed: int = 1, min_s: float = 0.3, max_s: float = 10.0,
                  tone_every: int = 8, chirp_every: int = 8) -> list[dict[str, Any]]:
    """n patterns with lengths linspace(min_s, max_s, n) seconds."""
    rs = np.random.RandomState(seed)
    out: list[dict[str, Any]] = []
    lengths = np.linspace(min_s, max_s, n) if n > 1 else np.array([min_s])
    for i, sec in enumerate(lengths):
        ns = int(round(float(sec) * sr))
        kind = "jingle"
        if tone_every and i % tone_every == tone_every - 1:
            kind = "tone"
        elif chirp_every and i % chirp_every == 0:
            kind = "chirp"
        if kind == "tone":
            f0 = float(rs.uniform(900.0, 1200.0))
            t = np.arange(ns, dtype=np.float32) / np.float32(sr)
            audio = (0.9 * np.sin(2 * np.pi * f0 * t)).astype(np.float32)
            out.append({"name": f"p{i:03d}_tone", "audio": audio, "strategy": "marker_tone",
                        "strategy_params": {"dominant_frequency_hz": f0}})
        elif kind == "chirp":
            f_a, f_b = float(rs.uniform(300, 900)), float(rs.uniform(1500, 3200))
            t = np.arange(ns, dtype=np.float64) / sr
            ph = 2 * np.pi * (f_a * t + (f_b - f_a) * t * t / (2 * max(float(sec), 1e-9)))
            audio = (0.8 * np.sin(ph) * np.hanning(ns)).astype(np.float32)
            out.append({"name": f"p{i:03d}_chirp", "audio": audio, "strategy": None, "strategy_params": {}})
        else:
            w = rs.standard_normal(ns).astype(np.float32)
            w = _one_pole(w, float(rs.uniform(0.2, 0.8)))
            env = 0.6 + 0.4 * np.sin(2 * np.pi * rs.uniform(0.5, 3.0) * np.arange(ns) / sr + rs.uniform(0, 6.28))
            audio = (0.3 * w / (np.std(w) + 1e-12) * env).astype(np.float32)
            out.append({"name": f"p{i:03d}_jingle", "audio": audio, "strategy": None, "strategy_params": {}})
    return out


def make_stream(seconds: float, patterns: list[dict[str, Any]], sr: int = 8000, seed: int = 0,
                plants_per_pattern: int = 2, chunk_seconds: Optional[int] = 60,
                bed_sigma: float = 0.1, gains: tuple[float, ...] = (1.0, 0.5)
                ) -> tuple[np.ndarray, list[tuple[str, int, float]]]:
    """Returns (float32 stream, [(pattern name, start sample, gain)])."""
    rs = np.random.RandomState(seed)
    n = int(round(seconds * sr))
    bed = rs.standard_normal(n).astype(np.float32) * np.float32(bed_sigma)
    audio = _one_pole(bed, 0.5)
    plants: list[tuple[str, int, float]] = []
    occupied: list[tuple[int, int]] = []
    C = int(chunk_seconds * sr) if chunk_seconds else 0
    for pi, p in enumerate(patterns):
        L = p["audio"].size
        if L + 2 >= n:
            continue
        for k in range(plants_per_pattern):
            for _attempt in range(50):
                if C and k == 0 and n > C + L and (pi % 2 == 0):
                    # straddle a chunk boundary: start inside the last L samples of a chunk
                    b = int(rs.randint(1, max(2, n // C))) * C
                    start = b - int(rs.randint(1, L))
                else:
                    start = int(rs.randint(0, n - L - 1))
                if start < 0 or start + L >= n:
                    continue
                if any(start < e + L and s - L < start + L for s, e in occupied):
                    continue
                break
            else:
                continue
            g = float(gains[(pi + k) % len(gains)])
            duck = 0.02 if p.get("strategy") == "marker_tone" else 0.1
            pad = L if p.get("strategy") == "marker_tone" else 0   # quiet flanks for tone clips
            a0, a1 = max(0, start - pad), min(n, start + L + pad)
            audio[a0:a1] *= np.float32(duck)
            audio[start:start + L] += np.float32(g) * p["audio"]
            occupied.append((a0, a1))
            plants.append((p["name"], start, g))
    return audio.astype(np.float32), plants


def describe(patterns: list[dict[str, Any]], sr: int) -> dict[str, Any]:
    ls = [p["audio"].size for p in patterns]
    return {"n_patterns": len(patterns), "min_len": min(ls), "max_len": max(ls),
            "sliding_windows": sorted({math.ceil(l / sr) for l in ls})}


def make_stream_device(seconds: float, patterns: list[dict[str, Any]], sr: int = 8000, seed: int = 0,
                       plants_per_pattern: int = 24, chunk_seconds: int = 60, bed_sigma: float = 0.1,
                       gains: tuple[float, ...] = (1.0, 0.5), device: str = "cuda"):
    """Same recipe as :func:`make_stream` but generated on the GPU (config 3-5 scale: 24 h = 691 M samples).

    Returns (float32 CUDA tensor, plants).  The noise bed comes from torch's device generator, so it
    is reproducible per seed on a given torch build; plant offsets come from numpy RandomState."""
    import torch
    n = int(round(seconds * sr))
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    audio = torch.empty(n, dtype=torch.float32, device=device)
    block = 1 << 26
    taps = torch.tensor([0.5 ** (k + 1) for k in range(32)], dtype=torch.float32, device=device).flip(0).view(1, 1, -1)
    prev = torch.zeros(31, dtype=torch.float32, device=device)
    for a in range(0, n, block):
        b = min(n, a + block)
        w = torch.randn(b - a, generator=gen, dtype=torch.float32, device=device) * bed_sigma
        x = torch.cat([prev, w])
        audio[a:b] = torch.nn.functional.conv1d(x.view(1, 1, -1), taps).view(-1)   # one-pole low-pass, a = 0.5
        prev = x[-31:].clone()
    rs = np.random.RandomState(seed + 1000003)
    C = int(chunk_seconds * sr)
    plants: list[tuple[str, int, float]] = []
    taken = np.zeros((n + C - 1) // C + 1, dtype=np.int32)   # coarse occupancy: at most a few plants per chunk
    spans: list[tuple[int, int]] = []
    for pi, p in enumerate(patterns):
        L = p["audio"].size
        if L + 2 >= n:
            continue
        pt = torch.from_numpy(p["audio"]).to(device)
        for k in range(plants_per_pattern):
            for _attempt in range(50):
                if k % 6 == 0 and n > C + L:
                    b0 = int(rs.randint(1, max(2, n // C))) * C
                    start = b0 - int(rs.randint(1, L))
                else:
                    start = int(rs.randint(0, n - L - 1))
                if start < 0 or start + L >= n:
                    continue
                pad = L if p.get("strategy") == "marker_tone" else 0
                a0, a1 = max(0, start - pad), min(n, start + L + pad)
                c0, c1 = a0 // C, a1 // C
                if taken[c0:c1 + 1].max() >= 2:
                    continue
                if any(a0 < e and s < a1 for s, e in spans[-4096:] if abs(s - a0) < 4 * C):
                    continue
                break
            else:
                continue
            taken[c0:c1 + 1] += 1
            spans.append((a0, a1))
            g = float(gains[(pi + k) % len(gains)])
            duck = 0.02 if p.get("strategy") == "marker_tone" else 0.1
            audio[a0:a1] *= duck
            audio[start:start + L] += g * pt
            plants.append((p["name"], start, g))
    return audio, plants


def _block_noise(block: int, block_size: int, seed: int, bed_sigma: float, device: str):
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed((seed * 1000003 + block) & 0x7FFFFFFFFFFF)
    return torch.randn(block_size, generator=gen, dtype=torch.float32, device=device) * bed_sigma


def make_stream_slab_device(seconds_total: float, patterns: list[dict[str, Any]], sr: int, seed: int,
                            plants_per_pattern: int, chunk_seconds: int, lo: int, hi: int, bed_sigma: float = 0.1,
                            gains: tuple[float, ...] = (1.0, 0.5), device: str = "cuda"):
    """Samples [lo, hi) of ONE logical stream of ``seconds_total`` seconds, generated on the device so that every
    process of a sharded scan (any world size) sees the same stream: the noise bed is drawn per fixed 4 Mi-sample
    block from a generator seeded by (seed, block index), the one-pole low-pass carries across blocks through the
    previous block's last 31 noise samples, and the plant list is planned for the whole stream on the host and applied
    where it overlaps the slab.  Returns (float32 CUDA tensor of hi - lo samples, plants of the whole stream)."""
    import torch
    n = int(round(seconds_total * sr))
    lo, hi = max(0, int(lo)), min(n, int(hi))
    bs = 1 << 22
    out = torch.empty(max(0, hi - lo), dtype=torch.float32, device=device)
    taps = torch.tensor([0.5 ** (k + 1) for k in range(32)], dtype=torch.float32, device=device).flip(0).view(1, 1, -1)
    for b in range(lo // bs, (hi + bs - 1) // bs if hi > lo else 0):
        w = _block_noise(b, bs, seed, bed_sigma, device)
        prev = _block_noise(b - 1, bs, seed, bed_sigma, device)[-31:] if b > 0 else torch.zeros(31, device=device)
        y = torch.nn.functional.conv1d(torch.cat([prev, w]).view(1, 1, -1), taps).view(-1)
        a, e = max(lo, b * bs), min(hi, (b + 1) * bs)
        out[a - lo:e - lo] = y[a - b * bs:e - b * bs]
    rs = np.random.RandomState(seed + 1000003)
    C = int(chunk_seconds * sr)
    plants: list[tuple[str, int, float]] = []
    taken = np.zeros((n + C - 1) // C + 1, dtype=np.int32)
    spans: list[tuple[int, int]] = []
    for pi, p in enumerate(patterns):
        L = p["audio"].size
        if L + 2 >= n:
            continue
        pt = None
        for k in range(plants_per_pattern):
            for _attempt in range(50):
                if k % 6 == 0 and n > C + L:
                    b0 = int(rs.randint(1, max(2, n // C))) * C
                    start = b0 - int(rs.randint(1, L))
                else:
                    start = int(rs.randint(0, n - L - 1))
                if start < 0 or start + L >= n:
                    continue
                pad = L if p.get("strategy") == "marker_tone" else 0
                a0, a1 = max(0, start - pad), min(n, start + L + pad)
                c0, c1 = a0 // C, a1 // C
                if taken[c0:c1 + 1].max() >= 2:
                    continue
                if any(a0 < e and s_ < a1 for s_, e in spans[-4096:] if abs(s_ - a0) < 4 * C):
                    continue
                break
            else:
                continue
            taken[c0:c1 + 1] += 1
            spans.append((a0, a1))
            g = float(gains[(pi + k) % len(gains)])
            plants.append((p["name"], start, g))
            if a1 <= lo or a0 >= hi:
                continue
            if pt is None:
                pt = torch.from_numpy(p["audio"]).to(device)
            duck = 0.02 if p.get("strategy") == "marker_tone" else 0.1
            out[max(a0, lo) - lo:min(a1, hi) - lo] *= duck
            s0, s1 = max(start, lo), min(start + L, hi)
            if s1 > s0:
                out[s0 - lo:s1 - lo] += g * pt[s0 - start:s1 - start]
    return out, plants
