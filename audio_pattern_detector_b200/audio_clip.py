"""Pattern clip / stream value types (reference audio_clip.py:17-102)."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Protocol

import numpy as np
from numpy.typing import NDArray

from .audio_utils import DEFAULT_TARGET_SAMPLE_RATE, load_wav_from_bytes, load_wave_file, resample_audio
from .pattern_config import APD_EXTENSION, load_apd_file


class ReadableStream(Protocol):
    def read(self, size: int, /) -> bytes: ...


@dataclass(frozen=True)
class AudioClip:
    name: str
    audio: NDArray[np.float32]
    sample_rate: int
    strategy: str | None = None                       # "marker_tone" for .apd.toml clips
    strategy_params: dict[str, Any] = field(default_factory=dict)

    @staticmethod
    def from_audio_file(clip_path: str | Path, sample_rate: int | None = None) -> "AudioClip":
        sr = DEFAULT_TARGET_SAMPLE_RATE if sample_rate is None else sample_rate
        p = str(clip_path)
        if p.lower().endswith(APD_EXTENSION):
            cfg = load_apd_file(clip_path, sample_rate=sr)
            return AudioClip(name=Path(p[:-len(APD_EXTENSION)]).name, audio=cfg.audio, sample_rate=sr,
                             strategy=cfg.strategy, strategy_params=cfg.strategy_params)
        return AudioClip(name=Path(clip_path).stem, audio=load_wave_file(p, expected_sample_rate=sr), sample_rate=sr)

    @staticmethod
    def from_wav_bytes(wav_bytes: bytes, name: str, sample_rate: int | None = None) -> "AudioClip":
        sr = DEFAULT_TARGET_SAMPLE_RATE if sample_rate is None else sample_rate
        audio, src_sr = load_wav_from_bytes(wav_bytes, name)
        if src_sr != sr:
            audio = resample_audio(audio, src_sr, sr)
        return AudioClip(name=name, audio=audio, sample_rate=sr)

    def clip_length_seconds(self) -> float:
        return len(self.audio) / self.sample_rate


@dataclass(frozen=True)
class AudioStream:
    name: str
    audio_stream: ReadableStream      # raw float32 mono PCM bytes at sample_rate
    sample_rate: int
