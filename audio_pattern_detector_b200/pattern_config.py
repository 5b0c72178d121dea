"""`.apd.toml` pattern files (reference pattern_config.py:34-220).

Config parsing is host logic; it stays bit-identical to the reference because the
clip samples it produces feed the correlation (the sine is synthesised with a
float32 time base, reference pattern_config.py:106-108).
"""
from __future__ import annotations

import base64
import binascii
import tomllib
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np
from numpy.typing import NDArray

from .audio_utils import load_wav_from_bytes, resample_audio

APD_EXTENSION = ".apd.toml"
VALID_STRATEGIES = frozenset({"marker_tone"})
VALID_CLIP_SOURCES = frozenset({"sine", "wav_base64"})
VALID_VERIFICATION_THRESHOLDS = frozenset({
    "minimum_band_purity", "minimum_active_frame_ratio", "minimum_longest_active_run",
    "minimum_active_frame_mean_purity", "maximum_min_flank_purity", "maximum_max_flank_purity"})
_CLIP_FIELDS = {"sine": frozenset({"frequency_hz", "duration_seconds", "amplitude"}),
                "wav_base64": frozenset({"data"})}
_VERIFICATION_FIELDS = VALID_VERIFICATION_THRESHOLDS | {"strategy", "dominant_frequency_hz"}
_TOP_LEVEL_FIELDS = frozenset({"description", "clip", "verification"})


@dataclass(frozen=True)
class PatternConfig:
    strategy: str
    strategy_params: dict[str, Any]
    audio: NDArray[np.float32]


def _need(table: dict[str, Any], key: str, kinds: type | tuple[type, ...], where: str) -> Any:
    if key not in table:
        raise ValueError(f"{where}: missing required field '{key}'")
    v = table[key]
    if not isinstance(v, kinds):
        names = kinds.__name__ if isinstance(kinds, type) else "/".join(k.__name__ for k in kinds)
        raise ValueError(f"{where}: field '{key}' must be {names}, got {type(v).__name__}")
    return v


def _reject_unknown(table: dict[str, Any], allowed: frozenset[str] | set[str], where: str, what: str) -> None:
    extra = sorted(set(table) - set(allowed))
    if extra:
        raise ValueError(f"{where}: unknown {what} field(s): {extra}. Valid fields: {sorted(allowed)}")


def _sine(clip: dict[str, Any], sr: int, where: str) -> NDArray[np.float32]:
    f = float(_need(clip, "frequency_hz", (int, float), where))
    dur = float(_need(clip, "duration_seconds", (int, float), where))
    amp = float(clip.get("amplitude", 0.9))
    if f <= 0:
        raise ValueError(f"{where}: frequency_hz must be positive, got {f}")
    if dur <= 0:
        raise ValueError(f"{where}: duration_seconds must be positive, got {dur}")
    if not f * 2 < sr:
        raise ValueError(f"{where}: frequency_hz {f} exceeds Nyquist ({sr / 2}) for sample_rate {sr}")
    t = np.arange(int(round(dur * sr)), dtype=np.float32) / np.float32(sr)
    return (amp * np.sin(2 * np.pi * f * t)).astype(np.float32)


def _wav_base64(clip: dict[str, Any], sr: int, where: str) -> NDArray[np.float32]:
    text = "".join(_need(clip, "data", str, where).split())
    try:
        blob = base64.b64decode(text, validate=True)
    except binascii.Error as e:
        raise ValueError(f"{where}: invalid base64 in [clip].data: {e}") from e
    audio, src_sr = load_wav_from_bytes(blob, name=where)
    return resample_audio(audio, src_sr, sr) if src_sr != sr else audio


def load_apd_file(path: str | Path, sample_rate: int) -> PatternConfig:
    where = str(path)
    with open(path, "rb") as fh:
        try:
            doc = tomllib.load(fh)
        except tomllib.TOMLDecodeError as e:
            raise ValueError(f"{where}: invalid TOML: {e}") from e
    extra = sorted(set(doc) - _TOP_LEVEL_FIELDS)
    if extra:
        raise ValueError(f"{where}: unknown top-level field(s): {extra}. Valid fields: {sorted(_TOP_LEVEL_FIELDS)} "
                         "(note: 'strategy' moved into [verification] in the v2 schema)")
    clip = _need(doc, "clip", dict, where)
    kind = _need(clip, "source", str, where)
    if kind not in VALID_CLIP_SOURCES:
        raise ValueError(f"{where}: unknown [clip].source '{kind}'. Valid sources: {sorted(VALID_CLIP_SOURCES)}")
    extra = sorted(set(clip) - _CLIP_FIELDS[kind] - {"source"})
    if extra:                                                            # reference pattern_config.py:88-93,114-119
        raise ValueError(f"{where}: unknown [clip] field(s) for source='{kind}': {extra}. "
                         f"Valid fields: {sorted(_CLIP_FIELDS[kind])}")
    audio = _sine(clip, sample_rate, where) if kind == "sine" else _wav_base64(clip, sample_rate, where)

    ver = _need(doc, "verification", dict, where)
    _reject_unknown(ver, _VERIFICATION_FIELDS, where, "[verification]")
    strategy = _need(ver, "strategy", str, where)
    if strategy not in VALID_STRATEGIES:
        raise ValueError(f"{where}: unknown strategy '{strategy}'. Valid strategies: {sorted(VALID_STRATEGIES)}")
    params: dict[str, Any] = {}
    if "dominant_frequency_hz" in ver:
        params["dominant_frequency_hz"] = float(_need(ver, "dominant_frequency_hz", (int, float), where))
    elif kind == "sine":
        params["dominant_frequency_hz"] = float(clip["frequency_hz"])
    thresholds: dict[str, float | int] = {}
    for key in sorted(set(ver) & VALID_VERIFICATION_THRESHOLDS):
        if key == "minimum_longest_active_run":
            thresholds[key] = int(_need(ver, key, int, where))
        else:
            thresholds[key] = float(_need(ver, key, (int, float), where))
    if thresholds:
        params["verification"] = thresholds
    return PatternConfig(strategy=strategy, strategy_params=params, audio=audio)
