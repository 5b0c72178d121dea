"""Row N2 (FFT resampler on the device): apd_resample against the oracle's restatement of the reference's
resample_1d (native-helper/src/lib.rs:235-275), and the resampling WAV stream path against the reference-style host
path of the same wrapper.

Tolerance: both sides compute in float64 and round once to float32, so outputs are identical except where the
float64 value sits within ~1e-13 of a float32 rounding boundary: at most 1 float32 ulp of the signal's peak, on a
vanishing fraction of samples."""
import os
import wave

import numpy as np
import pytest

from oracle import native
from tests.golden_util import load_json, synthetic_inputs
from tests.gpu_compare import make_detector

pytestmark = pytest.mark.gpu


def device_resample(x: np.ndarray, m: int) -> np.ndarray:
    import torch
    from audio_pattern_detector_b200.resample import resample
    out = resample(torch.from_numpy(np.ascontiguousarray(x)).cuda(), m)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def check(got: np.ndarray, want: np.ndarray) -> None:
    assert got.dtype == np.float32 and got.shape == want.shape
    if want.size == 0:
        return
    ulp = float(np.max(np.abs(want))) * 2.0 ** -23 + 1e-30
    assert float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) <= ulp
    assert np.count_nonzero(got != want) <= max(2, want.size // 1000)


# 7-smooth and not, odd and even, up and down, every radix (25, 16, 8, 7, 5, 4, 3, 2, 1), tiny lengths
PAIRS = [(4, 4), (0, 5), (2, 0), (8, 4), (3, 6), (5, 10), (160, 80), (1, 7), (7, 1), (2, 3), (11, 13), (13, 11),
         (1000, 999), (4410, 800), (800, 4410), (2 * 3 * 5 * 7 * 16 * 25, 2 * 3 * 5 * 7 * 8), (16000, 8000),
         (44100, 8000), (8000, 44100), (22050, 16000), (10007, 5003), (5003, 10007), (65536, 48000), (12345, 12346),
         (48000 * 3, 8000 * 3), (99991, 18139),
         # even -> even runs at half length (packed real transforms); halves that are prime go through Bluestein
         (2, 4), (4, 2), (6, 2), (2, 6), (20014, 10006), (10006, 20014), (6, 10006), (9998, 8000), (8000, 9998)]


@pytest.mark.parametrize("n,m", PAIRS)
def test_matches_oracle(n, m):
    x = (np.random.default_rng(n * 31 + m).standard_normal(n) * 0.3).astype(np.float32)
    check(device_resample(x, m), native.resample(x, m))


def test_reference_kats():                               # lib.rs:855-890
    out = device_resample(np.array([1, 2, 3, 4], np.float32), 4)
    assert np.allclose(out, [1, 2, 3, 4], atol=1e-5)
    assert device_resample(np.zeros(0, np.float32), 5).tolist() == [0.0] * 5
    sine = np.sin(2.0 * np.float32(np.pi) * np.arange(8, dtype=np.float32) / np.float32(8)).astype(np.float32)
    out = device_resample(sine, 4)
    assert abs(out[0]) < 0.1 and abs(out[1] - 1.0) < 0.1


def test_batched_chunks_match_single_calls():
    import torch
    from audio_pattern_detector_b200.resample import resample
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((5, 32000)) * 0.2).astype(np.float32)
    got = resample(torch.from_numpy(x).cuda(), 16000).cpu().numpy()
    for b in range(5):
        check(got[b], native.resample(x[b], 16000))
    x = (rng.standard_normal((3, 4999)) * 0.2).astype(np.float32)              # Bluestein both ways, batched
    got = resample(torch.from_numpy(x).cuda(), 2503).cpu().numpy()
    for b in range(3):
        check(got[b], native.resample(x[b], 2503))


@pytest.mark.parametrize("n,m", [(960000, 480000), (2646000, 480000), (959999, 479999)])
def test_chunk_sized(n, m):
    """One 60 s chunk of a 16 kHz / 44.1 kHz source -> 8 kHz (BASELINE chunk shape), and an odd-length final chunk."""
    x = (np.random.default_rng(n).standard_normal(n) * 0.25).astype(np.float32)
    check(device_resample(x, m), native.resample(x, m))


def test_properties_at_full_size():
    """Size-independent properties on a 60 s, 48 kHz chunk: linearity, and an odd-length signal survives an
    up-sampling round trip (no Nyquist bin to lose)."""
    import torch
    from audio_pattern_detector_b200.resample import resample
    g = torch.Generator(device="cuda").manual_seed(11)
    n, m = 2880001, 480000
    a = torch.randn(n, device="cuda", generator=g) * 0.2
    b = torch.randn(n, device="cuda", generator=g) * 0.2
    ra, rb, rab = resample(a, m), resample(b, m), resample(a + 2.0 * b, m)
    assert float((rab - (ra + 2.0 * rb)).abs().max()) < 2e-6
    x = a[:479999].contiguous()
    back = resample(resample(x, 960001), 479999)
    assert float((back - x).abs().max()) < 1e-6
    assert torch.equal(resample(x, 479999), x)


def write_wav(path, pcm: np.ndarray, rate: int) -> None:
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(pcm.tobytes())


@pytest.mark.parametrize("in_rate", [16000, 11025])
def test_resampling_wav_stream_matches_the_host_path(tmp_path, in_rate):
    """A WAV at another rate than the detector's: every chunk read is resampled on the device
    (_find_clip_in_pcm) and must give what the reference-style host path of the same wrapper gives (read():
    numpy float64 FFT per chunk, match.py:395-423), including the odd-length final chunk."""
    from audio_pattern_detector_b200.audio_clip import AudioStream
    from audio_pattern_detector_b200.match import _WavFileStreamWrapper
    run = [r for r in load_json("synthetic_runs.json") if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    up = native.resample(audio, int(len(audio) * in_rate / 8000))[:-137]       # ragged tail
    pcm = np.clip(np.round(up * 32768.0), -32768, 32767).astype(np.int16)
    path = tmp_path / "other_rate.wav"
    write_wav(path, pcm, in_rate)
    det = make_detector(clips, 8000, 10, max_batch_chunks=3)

    host = _WavFileStreamWrapper(str(path), 8000)
    assert host.needs_resample
    host.pcm_format = None
    seen_h = []
    times_h, total_h = det.find_clip_in_audio(AudioStream(name="s", audio_stream=host, sample_rate=8000),
                                              on_pattern_detected=lambda n, t: seen_h.append((t, n)))
    host.close()

    dev = _WavFileStreamWrapper(str(path), 8000)
    assert dev.pcm_format == (2, 1) and dev.pcm_sample_rate == in_rate
    seen_d = []
    times_d, total_d = det.find_clip_in_audio(AudioStream(name="s", audio_stream=dev, sample_rate=8000),
                                              on_pattern_detected=lambda n, t: seen_d.append((t, n)))
    dev.close()
    assert sum(len(v) for v in times_h.values()) > 0
    assert times_d == times_h and seen_d == seen_h and total_d == total_h


@pytest.mark.parametrize("frames", [1, 5, 160001])
def test_resampling_stream_edge_lengths(tmp_path, frames):
    """Files shorter than one chunk read: a single frame resamples to nothing (the reference's read() then returns
    b"" = end of stream), a handful of frames to a 2-sample final chunk, and one chunk plus one frame to a full chunk
    and nothing."""
    from audio_pattern_detector_b200.audio_clip import AudioStream
    from audio_pattern_detector_b200.match import _WavFileStreamWrapper
    run = [r for r in load_json("synthetic_runs.json") if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    up = native.resample(audio[:80000], 160000)
    pcm = np.clip(np.round(np.resize(up, frames) * 32768.0), -32768, 32767).astype(np.int16)
    path = tmp_path / "short.wav"
    write_wav(path, pcm, 16000)
    det = make_detector(clips, 8000, 10, max_batch_chunks=2)
    out = []
    for device_path in (False, True):
        w = _WavFileStreamWrapper(str(path), 8000)
        if not device_path:
            w.pcm_format = None
        out.append(det.find_clip_in_audio(AudioStream(name="s", audio_stream=w, sample_rate=8000)))
        w.close()
    assert out[0] == out[1]
    assert out[1][1] == {1: 0.0, 5: 2 / 8000, 160001: 10.0}[frames]


RESAMPLE_RUNS = load_json("resample_runs.json")


@pytest.mark.parametrize("run", RESAMPLE_RUNS, ids=lambda r: f"{os.path.basename(r['wav'])}/{r['spc']}")
def test_resampling_stream_matches_the_reference_golden(tmp_path, run):
    """End to end against the UNMODIFIED reference: its match_pattern on the 16 kHz fixtures with an 8 kHz detector
    (every chunk read resampled, match.py:395-423) gave tests/golden/resample_runs.json (oracle/make_golden_resample.py);
    here the same WAV goes through the device PCM + resampler path."""
    from audio_pattern_detector_b200.audio_clip import AudioStream
    from audio_pattern_detector_b200.match import _WavFileStreamWrapper
    from tests.golden_util import fixture_clips, fixtures
    meta = [r for r in load_json("fixture_runs.json") if r["sr"] == 8000 and len(r["clips"]) >= 6][0]
    clips = fixture_clips(meta)
    assert sorted(c["name"] for c in clips) == sorted(run["timestamps"])
    path = tmp_path / "in16k.wav"
    write_wav(path, fixtures()["wav:" + run["wav"]], run["wav_rate"])
    det = make_detector(clips, 8000, run["spc"], max_batch_chunks=2)
    w = _WavFileStreamWrapper(str(path), 8000)
    assert w.pcm_format == (2, 1) and w.pcm_sample_rate == 16000
    seen = []
    times, total = det.find_clip_in_audio(AudioStream(name="s", audio_stream=w, sample_rate=8000),
                                          on_pattern_detected=lambda n, t: seen.append([n, t]))
    w.close()
    assert times == run["timestamps"] and seen == run["events"] and total == run["total_time"]
