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
