"""Host-side audio helpers: WAV decode, FFT resampling, ffmpeg pipes.

Mirrors the public helpers of the reference's audio_utils.py (file:line in each
docstring).  Pattern-clip decode is outside the detection hot path, so this is plain numpy
host code; the streaming side of rows N1/N2 runs on the
device (csrc/pcm.cu, csrc/resample.cu, resample.py).
"""
from __future__ import annotations

import io
import math
import subprocess
import sys
import wave
from contextlib import contextmanager
from typing import IO, Any, Iterator

import numpy as np
from numpy.typing import NDArray

DEFAULT_TARGET_SAMPLE_RATE = 8000      # reference audio_utils.py:13

_ffmpeg_available: bool | None = None


def is_ffmpeg_available() -> bool:
    """reference audio_utils.py:19-39."""
    global _ffmpeg_available
    if _ffmpeg_available is None:
        try:
            subprocess.run(["ffmpeg", "-version"], capture_output=True, check=True)
            _ffmpeg_available = True
        except (subprocess.CalledProcessError, FileNotFoundError):
            _ffmpeg_available = False
    return _ffmpeg_available


def pcm_to_float32(raw: bytes, sampwidth: int, channels: int) -> NDArray[np.float32]:
    """Integer PCM frames -> mono float32 in [-1, 1) (reference audio_utils.py:60-79,132-151)."""
    if sampwidth == 2:
        x = np.frombuffer(raw, dtype=np.int16).astype(np.float32) / 32768.0
    elif sampwidth == 4:
        x = np.frombuffer(raw, dtype=np.int32).astype(np.float32) / 2147483648.0
    elif sampwidth == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif sampwidth == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(b[:, 2] >= 0x80, v - (1 << 24), v).astype(np.int32) << 8
        x = v.astype(np.float32) / 2147483648.0
    else:
        raise ValueError(f"Unsupported sample width {sampwidth}")
    if channels > 1:
        x = x.reshape(-1, channels).mean(axis=1).astype(np.float32)
    return x


def _open_wav(src: "str | IO[bytes]", what: str) -> tuple[NDArray[np.float32], int]:
    try:
        with wave.open(src, "rb") as w:
            sr, ch, sw = w.getframerate(), w.getnchannels(), w.getsampwidth()
            raw = w.readframes(w.getnframes())
    except Exception as e:  # noqa: BLE001 - same contract as the reference: any decode error -> ValueError
        raise ValueError(f"Failed to read WAV data from {what}: {e}") from e
    try:
        return pcm_to_float32(raw, sw, ch), sr
    except ValueError as e:
        raise ValueError(f"{e} in {what}") from e


def load_wav_file(file_path: str) -> tuple[NDArray[np.float32], int]:
    """reference audio_utils.py:82-95."""
    return _open_wav(file_path, f"file {file_path}")


def load_wav_from_bytes(wav_bytes: bytes, name: str = "bytes") -> tuple[NDArray[np.float32], int]:
    """reference audio_utils.py:98-114."""
    return _open_wav(io.BytesIO(wav_bytes), name)


def resample_fft(x: NDArray[np.floating[Any]], num: int) -> NDArray[np.float32]:
    """scipy.signal.resample-style FFT resampling (reference native-helper/src/lib.rs:235-275)."""
    x64 = np.asarray(x, dtype=np.float64)
    n = x64.size
    if n == 0 or num == 0:
        return np.zeros(num, dtype=np.float32)
    if n == num:
        return x64.astype(np.float32)
    X = np.fft.fft(x64)
    keep = min(n, num)
    pos, neg = (keep + 1) // 2, (keep - 1) // 2
    Y = np.zeros(num, dtype=np.complex128)
    Y[:pos] = X[:pos]
    if neg:
        Y[num - neg:] = X[n - neg:]
    return (np.fft.ifft(Y, norm="forward").real * (1.0 / n)).astype(np.float32)     # lib.rs:268-274


def resample_audio(audio: NDArray[np.float32], orig_sr: int, target_sr: int) -> NDArray[np.float32]:
    """reference audio_utils.py:154-171."""
    if orig_sr == target_sr:
        return audio
    return resample_fft(audio, int(len(audio) * target_sr / orig_sr))


def slicing_with_zero_padding(array: Any, width: int, middle_index: int) -> NDArray[Any]:
    """Centred fixed-width slice, zero padded (reference audio_utils.py:177-191)."""
    a = np.asarray(array)
    lo = int(middle_index - math.floor(width / 2))
    hi = int(middle_index + math.ceil(width / 2))
    out = np.zeros(hi - lo, dtype=a.dtype)
    s0, s1 = max(lo, 0), min(hi, a.size)
    if s1 > s0:
        out[s0 - lo:s1 - lo] = a[s0:s1]
    return out


def load_wave_file(file_path: str, expected_sample_rate: int) -> NDArray[np.float32]:
    """reference audio_utils.py:194-228."""
    if file_path.lower().endswith(".wav"):
        data, sr = load_wav_file(file_path)
        return resample_audio(data, sr, expected_sample_rate) if sr != expected_sample_rate else data
    if not is_ffmpeg_available():
        raise ValueError(f"ffmpeg not available and file {file_path} is not a WAV file. "
                         "Install ffmpeg or use WAV files for patterns.")
    with ffmpeg_get_float32_pcm(file_path, target_sample_rate=expected_sample_rate, ac=1) as out:
        return np.frombuffer(out.read(), dtype=np.float32)


@contextmanager
def ffmpeg_get_float32_pcm(full_audio_path: str, target_sample_rate: int | None = None, ac: int | None = None,
                           from_stdin: bool = False, input_format: str | None = None) -> Iterator[IO[bytes]]:
    """Decode anything ffmpeg reads to a float32 PCM pipe (reference audio_utils.py:239-291)."""
    cmd = ["ffmpeg"]
    if from_stdin:
        if input_format:
            cmd += ["-f", input_format]
        cmd += ["-i", "pipe:0"]
    else:
        cmd += ["-i", full_audio_path]
    cmd += ["-f", "f32le", "-acodec", "pcm_f32le"]
    if ac is not None:
        cmd += ["-ac", str(ac)]
    if target_sample_rate is not None:
        cmd += ["-ar", str(target_sample_rate)]
    cmd += ["-loglevel", "error", "pipe:"]
    proc = None
    try:
        proc = subprocess.Popen(cmd, stdin=sys.stdin.buffer if from_stdin else None, stdout=subprocess.PIPE)
        assert proc.stdout is not None
        yield proc.stdout
        if proc.wait() != 0:
            raise ValueError(f"ffmpeg command failed with return code {proc.returncode}")
    finally:
        if proc is not None and proc.stdout is not None:
            proc.stdout.close()


def write_wav_file(filepath: str, audio_data: NDArray[np.float32], sample_rate: int) -> None:
    """Mono float32 samples -> a file in whatever container ffmpeg picks from the name (reference
    audio_utils.py:294-322); ``ValueError`` when ffmpeg fails."""
    cmd = ["ffmpeg", "-y", "-f", "f32le", "-ar", str(sample_rate), "-ac", "1", "-i", "pipe:", "-loglevel", "error",
           filepath]
    proc = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.DEVNULL)
    proc.communicate(input=np.ascontiguousarray(audio_data, dtype=np.float32).tobytes())
    if proc.returncode != 0:
        raise ValueError(f"ffmpeg write failed with return code {proc.returncode}")


def get_audio_duration(audio_path: str) -> float | None:
    """Duration in seconds from ffprobe, ``None`` when the source has none, e.g. a live stream (reference
    audio_utils.py:324-352)."""
    import json
    res = subprocess.run(["ffprobe", "-v", "error", "-show_entries", "format=duration", "-of", "json", audio_path],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise ValueError(f"ffprobe failed: {res.stderr}")
    duration = json.loads(res.stdout).get("format", {}).get("duration")
    return None if duration is None else float(duration)


def seconds_to_time(seconds: float, include_decimals: bool = True) -> str:
    """HH:MM:SS[.mmm].  Stands in for andrew_utils.seconds_to_time (a dependency absent from the
    reference tree; its format is pinned only by README.md:89-93 "00:00:05.500")."""
    ms_total = int(round(float(seconds) * 1000)) if include_decimals else int(seconds) * 1000
    h, rem = divmod(ms_total, 3600_000)
    m, rem = divmod(rem, 60_000)
    s, ms = divmod(rem, 1000)
    return f"{h:02d}:{m:02d}:{s:02d}.{ms:03d}" if include_decimals else f"{h:02d}:{m:02d}:{s:02d}"
