"""Row N1 (streaming ingestion): integer PCM widened on the device must reproduce the reference's host conversion
(_WavFileStreamWrapper.read / _normalize_wav_data) bit for bit, and the PCM paths must give the same detections."""
import ctypes as C
import wave

import numpy as np
import pytest

from tests.golden_util import load_json, synthetic_inputs
from tests.gpu_compare import make_detector

pytestmark = pytest.mark.gpu


def device_convert(pcm: np.ndarray, channels: int) -> np.ndarray:
    import torch
    from audio_pattern_detector_b200 import _lib
    t = torch.from_numpy(np.ascontiguousarray(pcm).reshape(-1)).cuda()
    frames = t.numel() // channels
    out = torch.empty(frames, dtype=torch.float32, device="cuda")
    rc = _lib.lib().apd_pcm_to_float(C.c_void_p(t.data_ptr()), pcm.dtype.itemsize, channels, frames,
                                     C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    return out.cpu().numpy()


@pytest.mark.parametrize("dtype,width", [(np.int16, 2), (np.int32, 4), (np.uint8, 1)])
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_device_conversion_is_bit_exact(dtype, width, channels):
    from audio_pattern_detector_b200.audio_utils import pcm_to_float32
    rs = np.random.RandomState(7)
    info = np.iinfo(dtype)
    pcm = rs.randint(info.min, info.max, size=50_001 * channels, dtype=np.int64).astype(dtype)
    pcm[:4 * channels] = np.array([info.min, info.max, 0, -1] * channels).astype(dtype)
    want = pcm_to_float32(pcm.tobytes(), width, channels)
    got = device_convert(pcm, channels)
    assert got.dtype == np.float32 and np.array_equal(got, want)


def quantised_case():
    run = [r for r in load_json("synthetic_runs.json") if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    pcm = np.clip(np.round(audio * 32768.0), -32768, 32767).astype(np.int16)
    return clips, pcm


def test_scan_array_accepts_pcm16():
    clips, pcm = quantised_case()
    det = make_detector(clips, 8000, 10, max_batch_chunks=1)          # several host segments
    want = det.scan_array(pcm.astype(np.float32) / 32768.0)
    got = det.scan_array(pcm)
    assert got.peak_times == want.peak_times and got.events == want.events
    assert sum(len(v) for v in want.peak_times.values()) > 0
    stereo = np.repeat(pcm, 2)                                        # identical channels: the mean is the sample
    assert det.scan_array(stereo, pcm_channels=2).peak_times == want.peak_times


def test_wav_stream_wrapper_takes_the_pcm_path(tmp_path):
    from audio_pattern_detector_b200.audio_clip import AudioStream
    from audio_pattern_detector_b200.match import _WavFileStreamWrapper
    clips, pcm = quantised_case()
    path = tmp_path / "stream.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(8000)
        w.writeframes(np.repeat(pcm, 2).tobytes())
    det = make_detector(clips, 8000, 10, max_batch_chunks=3)
    want = det.scan_array(pcm.astype(np.float32) / 32768.0)
    wrapper = _WavFileStreamWrapper(str(path), 8000)
    assert wrapper.pcm_format == (2, 2)
    seen = []
    times, total = det.find_clip_in_audio(AudioStream(name="s", audio_stream=wrapper, sample_rate=8000),
                                          on_pattern_detected=lambda n, t: seen.append((t, n)))
    wrapper.close()
    assert times == want.peak_times and seen == want.events and total == want.total_time
    # the float path of the same wrapper (what the reference does on the host) agrees
    wrapper = _WavFileStreamWrapper(str(path), 8000)
    wrapper.pcm_format = None
    times2, _ = det.find_clip_in_audio(AudioStream(name="s", audio_stream=wrapper, sample_rate=8000))
    wrapper.close()
    assert times2 == times


def test_wav_stdin_wrapper_takes_the_pcm_path(monkeypatch):
    """A 16-bit WAV arriving on stdin (match --from-stdin-wav style): the frames are widened on the device."""
    import io
    import sys
    import types
    from audio_pattern_detector_b200.audio_clip import AudioStream
    from audio_pattern_detector_b200.match import _WavStdinStreamWrapper
    clips, pcm = quantised_case()
    blob = io.BytesIO()
    with wave.open(blob, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(8000)
        w.writeframes(pcm.tobytes())
    monkeypatch.setattr(sys, "stdin", types.SimpleNamespace(buffer=io.BytesIO(blob.getvalue())))
    wrapper = _WavStdinStreamWrapper(8000)
    assert wrapper.pcm_format == (2, 1)
    det = make_detector(clips, 8000, 10, max_batch_chunks=1)              # live pipe: chunk by chunk
    want = det.scan_array(pcm.astype(np.float32) / 32768.0)
    seen = []
    times, total = det.find_clip_in_audio(AudioStream(name="stdin", audio_stream=wrapper, sample_rate=8000),
                                          on_pattern_detected=lambda n, t: seen.append((t, n)))
    assert times == want.peak_times and seen == want.events and total == want.total_time


def test_eight_bit_wav_stream(tmp_path):
    """8-bit unsigned WAV: widened on the device like the reference's host conversion ((x - 128) / 128)."""
    from audio_pattern_detector_b200.audio_clip import AudioStream
    from audio_pattern_detector_b200.match import _WavFileStreamWrapper
    clips, pcm = quantised_case()
    u8 = np.clip(np.round(pcm.astype(np.float32) / 256.0) + 128, 0, 255).astype(np.uint8)
    path = tmp_path / "u8.wav"
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(1)
        w.setframerate(8000)
        w.writeframes(u8.tobytes())
    det = make_detector(clips, 8000, 10, max_batch_chunks=3)
    want = det.scan_array((u8.astype(np.float32) - 128.0) / 128.0)
    wrapper = _WavFileStreamWrapper(str(path), 8000)
    assert wrapper.pcm_format == (1, 1)
    times, _ = det.find_clip_in_audio(AudioStream(name="u8", audio_stream=wrapper, sample_rate=8000))
    wrapper.close()
    assert times == want.peak_times
    assert det.scan_array(u8).peak_times == want.peak_times
