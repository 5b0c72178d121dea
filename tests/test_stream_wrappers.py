"""Host logic of the streaming wrappers (rows N1 / N2), no GPU needed: which WAV formats take the raw-PCM path, and
that the multi-threaded positional reads of readinto_pcm return exactly the file's data chunk."""
import io
import struct
import sys
import types
import wave

import numpy as np
import pytest

from audio_pattern_detector_b200 import match
from audio_pattern_detector_b200.match import _WavFileStreamWrapper, _WavStdinStreamWrapper


def write_wav(path, pcm: np.ndarray, rate: int, channels: int = 1) -> None:
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(pcm.dtype.itemsize)
        w.setframerate(rate)
        w.writeframes(pcm.tobytes())


@pytest.mark.parametrize("dtype,rate,channels,expect", [
    (np.int16, 8000, 1, (2, 1)), (np.int16, 16000, 2, (2, 2)), (np.int32, 8000, 1, (4, 1)), (np.uint8, 11025, 1, (1, 1)),
])
def test_pcm_path_selection(tmp_path, dtype, rate, channels, expect):
    pcm = np.zeros(100 * channels, dtype=dtype)
    path = tmp_path / "a.wav"
    write_wav(path, pcm, rate, channels)
    w = _WavFileStreamWrapper(str(path), 8000)
    assert w.pcm_format == expect and w.pcm_sample_rate == rate and w.needs_resample == (rate != 8000)
    w.close()


def test_24_bit_wav_keeps_the_host_path(tmp_path):
    path = tmp_path / "b.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(3)
        w.setframerate(8000)
        w.writeframes(bytes(300))
    w = _WavFileStreamWrapper(str(path), 8000)
    assert w.pcm_format is None
    with pytest.raises(ValueError, match="Unsupported WAV sample width"):      # reference match.py:411
        w.read(400)
    w.close()


def test_missing_file_is_a_value_error():
    with pytest.raises(ValueError, match="Failed to read WAV file"):
        _WavFileStreamWrapper("/nonexistent/x.wav", 8000)


@pytest.mark.parametrize("frames", [0, 1, 4097, 6_000_001])
def test_readinto_pcm_returns_the_data_chunk(tmp_path, monkeypatch, frames):
    """Reads larger than 8 MB are split over positional-read threads; the pieces must line up, the read sizes need
    not divide the file, and a LIST chunk in front of the data chunk is skipped."""
    monkeypatch.setattr(match, "_READ_THREADS", 3)
    rs = np.random.RandomState(frames % 97)
    pcm = rs.randint(-32768, 32767, size=frames * 2).astype(np.int16)          # stereo
    path = tmp_path / "c.wav"
    write_wav(path, pcm, 8000, 2)
    raw = bytearray(path.read_bytes())
    # splice a LIST chunk (odd size, padded) between fmt and data
    at = raw.index(b"data")
    extra = b"LIST" + struct.pack("<I", 5) + b"hello" + b"\x00"
    raw[at:at] = extra
    raw[4:8] = struct.pack("<I", len(raw) - 8)
    path.write_bytes(bytes(raw))
    w = _WavFileStreamWrapper(str(path), 8000)
    assert w._wav.getnframes() == frames
    buf = np.zeros(2_500_000 * 2, dtype=np.int16)                              # 10 MB per read -> threaded
    got = []
    while True:
        n = w.readinto_pcm(buf, 2_500_000)
        if n == 0:
            break
        got.append(buf[:n * 2].copy())
    w.close()
    out = np.concatenate(got) if got else np.zeros(0, np.int16)
    assert out.size == pcm.size and np.array_equal(out, pcm)


def test_stdin_wrapper_formats(monkeypatch):
    def stdin_with(sampwidth: int, fmt_float: bool = False):
        blob = io.BytesIO()
        if fmt_float:
            data = np.linspace(-1, 1, 50, dtype=np.float32).tobytes()
            hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
                "<IHHIIHH", 16, 3, 1, 8000, 8000 * 4, 4, 32) + b"data" + struct.pack("<I", len(data))
            blob.write(hdr + data)
        else:
            with wave.open(blob, "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(sampwidth)
                w.setframerate(8000)
                w.writeframes(np.arange(50, dtype=np.int16 if sampwidth == 2 else np.int32).tobytes())
        monkeypatch.setattr(sys, "stdin", types.SimpleNamespace(buffer=io.BytesIO(blob.getvalue())))
        return _WavStdinStreamWrapper(8000)

    w = stdin_with(2)
    assert w.pcm_format == (2, 1)
    assert np.frombuffer(w.read_pcm(10), np.int16).tolist() == list(range(10))
    assert np.frombuffer(w.read(40), np.float32).tolist() == [k / 32768.0 for k in range(10, 20)]   # reference :318-325
    w = stdin_with(4)
    assert w.pcm_format == (4, 1)
    w = stdin_with(0, fmt_float=True)
    assert w.pcm_format is None and len(w.read(40)) == 40
