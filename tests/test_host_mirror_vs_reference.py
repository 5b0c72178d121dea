"""The host-side mirror modules (clip / WAV loading, .apd.toml patterns, the WAV stream wrapper's float path, slicing)
against the unmodified reference modules imported from /root/reference (oracle/refshim.py supplies stand-ins for the
three wheels that cannot be installed offline).  Same inputs, bit-identical outputs.  Skipped where the reference tree
is absent (the GPU box)."""
import glob
import importlib
import os

import numpy as np
import pytest

BASE = "/root/reference/sample_audios"
pytestmark = pytest.mark.skipif(not os.path.isdir(BASE), reason="reference tree not present")


@pytest.fixture(scope="module")
def mods():
    from oracle import refshim
    refshim.install()
    names = ("audio_clip", "audio_utils", "pattern_config", "detection_utils", "match")
    ref = {m: importlib.import_module(f"audio_pattern_detector.{m}") for m in names}
    own = {m: importlib.import_module(f"audio_pattern_detector_b200.{m}") for m in names}
    return ref, own


def same(a, b):
    if isinstance(a, np.ndarray):
        return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    return a == b


@pytest.mark.parametrize("sr", [8000, 16000])
def test_pattern_clips_load_identically(mods, sr):
    ref, own = mods
    files = sorted(glob.glob(BASE + "/clips/*"))
    assert len(files) >= 6
    for f in files:
        a = ref["audio_clip"].AudioClip.from_audio_file(f, sample_rate=sr)
        b = own["audio_clip"].AudioClip.from_audio_file(f, sample_rate=sr)
        assert a.name == b.name and a.strategy == b.strategy and a.strategy_params == b.strategy_params, f
        assert same(a.audio, b.audio), f
        assert ref["detection_utils"].get_pure_tone_frequency(a.audio, sr) == \
            own["detection_utils"].get_pure_tone_frequency(b.audio, sr), f


def test_wav_loading_and_resampling(mods):
    ref, own = mods
    files = sorted(glob.glob(BASE + "/*.wav") + glob.glob(BASE + "/test_16khz/*.wav"))[:6]
    for f in files:
        a, sa = ref["audio_utils"].load_wav_file(f)
        b, sb = own["audio_utils"].load_wav_file(f)
        assert sa == sb and same(a, b), f
        for sr in (8000, 16000):
            assert same(ref["audio_utils"].load_wave_file(f, sr), own["audio_utils"].load_wave_file(f, sr)), (f, sr)


def test_wav_stream_wrapper_float_path(mods):
    """read(): what find_clip_in_audio consumes chunk by chunk, with per-chunk resampling when the rates differ."""
    ref, own = mods
    files = sorted(glob.glob(BASE + "/*.wav"))[:3] + sorted(glob.glob(BASE + "/test_16khz/*.wav"))[:2]
    for f in files:
        for sr in (8000, 16000):
            wa = ref["match"]._WavFileStreamWrapper(f, sr)
            wb = own["match"]._WavFileStreamWrapper(f, sr)
            assert wa.needs_resample == wb.needs_resample and wa.input_sample_rate == wb.input_sample_rate
            while True:
                a, b = wa.read(4 * sr * 3), wb.read(4 * sr * 3)
                assert a == b, (f, sr)
                if not a:
                    break
            wa.close()
            wb.close()


def test_slicing_with_zero_padding(mods):
    ref, own = mods
    rng = np.random.default_rng(0)
    for _ in range(200):
        arr = rng.standard_normal(int(rng.integers(1, 40))).astype(np.float32)
        w, m = int(rng.integers(1, 30)), int(rng.integers(-5, 45))
        assert same(ref["audio_utils"].slicing_with_zero_padding(arr, w, m),
                    own["audio_utils"].slicing_with_zero_padding(arr, w, m)), (arr.size, w, m)


def _capture(fn):
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn()
    return buf.getvalue()


@pytest.mark.parametrize("fmt", ["both", "ms", "formatted"])
def test_jsonl_events_are_identical(mods, fmt):
    """match.py:524-565: pattern_detected / end events, the per-clip suppression of consecutive equal milliseconds,
    non-ASCII clip names."""
    ref, own = mods
    events = [("cbs_news", 25.89875), ("cbs_news", 25.8990), ("天空下的彩虹intro", 13.848), ("cbs_news", 25.8994),
              ("cbs_news", 3725.0004), ("x", 0.0), ("x", 0.0004), ("x", 0.0005), ("x", 86399.9996)]

    def run(m):
        cb = m._make_jsonl_callback(fmt)
        for n, t in events:
            cb(n, t)
        m._emit_jsonl_end(7325.4321, fmt)
        m._emit_jsonl("start", source="日本語.wav")

    assert _capture(lambda: run(ref["match"])) == _capture(lambda: run(own["match"]))


def _wav_header(fmt_tag=1, channels=1, rate=8000, bits=16, extra=b"", data=b"\x00" * 8):
    import struct
    fmt = struct.pack("<HHIIHH", fmt_tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_wav_header_validation_is_identical(mods):
    """match.py:215-283: accepted formats and every rejection message."""
    import io
    import struct
    ref, own = mods
    junk = b"LIST" + struct.pack("<I", 4) + b"abcd"
    cases = [_wav_header(), _wav_header(bits=32), _wav_header(fmt_tag=3, bits=32), _wav_header(extra=junk),
             _wav_header(bits=8), _wav_header(fmt_tag=3, bits=64), _wav_header(fmt_tag=2), _wav_header(channels=2),
             _wav_header(rate=16000), b"RIFX" + _wav_header()[4:], _wav_header()[:8] + b"WAVX" + _wav_header()[12:],
             _wav_header()[:20], _wav_header()[:12], b""]
    for blob in cases:
        out = []
        for m in (ref["match"], own["match"]):
            try:
                out.append(("ok", m._validate_wav_header(io.BytesIO(blob), 8000)))
            except ValueError as e:
                out.append(("ValueError", str(e)))
        assert out[0] == out[1], blob[:40]


def test_multiplexed_pattern_protocol_is_identical(mods, monkeypatch):
    """match.py:30-96: [uint32 n] then [uint32 name_len][name][uint32 data_len][wav] per pattern."""
    import io
    import struct
    import sys
    import types
    import wave
    ref, own = mods

    def wav_bytes(rate):
        b = io.BytesIO()
        with wave.open(b, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(rate)
            w.writeframes((np.sin(np.arange(1600) * 0.3) * 12000).astype(np.int16).tobytes())
        return b.getvalue()

    def frame(items):
        out = struct.pack("<I", len(items))
        for name, blob in items:
            nb = name.encode("utf-8")
            out += struct.pack("<I", len(nb)) + nb + struct.pack("<I", len(blob)) + blob
        return out

    good = frame([("beep", wav_bytes(8000)), ("彩虹", wav_bytes(16000))])
    streams = [good, good[:30], frame([]), frame([("a", wav_bytes(8000)), ("a", wav_bytes(8000))]),
               struct.pack("<I", 1) + struct.pack("<I", 0), b"\x01\x00"]
    for blob in streams:
        out = []
        for m in (ref["match"], own["match"]):
            monkeypatch.setattr(sys, "stdin", types.SimpleNamespace(buffer=io.BytesIO(blob)))
            try:
                clips = m._read_patterns_from_multiplexed_stdin(8000)
                out.append([(c.name, c.audio.tobytes(), c.sample_rate) for c in clips])
            except ValueError as e:
                out.append(("ValueError", str(e)))
        assert out[0] == out[1], blob[:24]


def test_detector_argument_validation_is_identical(mods):
    """apd.py:105-137: duplicate names, wrong clip sample rate, chunk too small -- the same ValueError text, raised
    before anything touches the device (so this runs without a GPU)."""
    ref, own = mods
    rdet = importlib.import_module("audio_pattern_detector.audio_pattern_detector").AudioPatternDetector
    odet = importlib.import_module("audio_pattern_detector_b200.audio_pattern_detector").AudioPatternDetector
    tone = (0.3 * np.sin(np.arange(20000) * 0.7)).astype(np.float32)

    def clips(m, spec):
        return [m["audio_clip"].AudioClip(name=n, audio=tone[:k].copy(), sample_rate=sr) for n, k, sr in spec]

    bad = [
        ([("a", 8000, 8000), ("a", 4000, 8000)], dict()),                                  # duplicate name
        ([("a", 8000, 16000)], dict()),                                                     # clip at another rate
        ([("a", 8000, 8000)], dict(target_sample_rate=16000)),
        ([("a", 20000, 8000)], dict(seconds_per_chunk=5)),                                  # 2.5 s clip needs >= 6 s
        ([("ok", 8000, 8000), ("long", 16001, 8000)], dict(seconds_per_chunk=4)),
    ]
    for spec, kw in bad:
        out = []
        for m, det in ((ref, rdet), (own, odet)):
            with pytest.raises(ValueError) as e:
                det(audio_clips=clips(m, spec), **kw)
            out.append(str(e.value))
        assert out[0] == out[1], (spec, kw)


def test_match_pattern_argument_errors_are_identical(mods, tmp_path):
    """match.py:120-149: missing audio, missing pattern, no patterns, unknown pattern type -- before any device work."""
    ref, own = mods
    audio = sorted(glob.glob(BASE + "/*.wav"))[0]
    clip = sorted(glob.glob(BASE + "/clips/*.wav"))[0]
    empty = tmp_path / "empty_dir"
    empty.mkdir()
    cases = [("/nonexistent/audio.wav", [clip]), (audio, ["/nonexistent/clip.wav"]), (audio, [])]
    for src, pats in cases:
        out = []
        for m in (ref["match"], own["match"]):
            with pytest.raises(ValueError) as e:
                m.match_pattern(src, pats)
            out.append(str(e.value))
        assert out[0] == out[1], (src, pats)
