"""The reference's OWN detector test-suites (tests/ref_suite/, vendored verbatim) against the CUDA detector, through
the import alias of tests/ref_alias_runner.py: sliding-window / chunk-loop timestamps, short-clip routing and the
seeded-noise false-positive test, the marker-tone verifier's known answers, the real-data regression set and the
get_config contract.  ffmpeg is not in the image: the tests that pipe audio through it are deselected by name."""
import os
import shutil
import subprocess
import sys
import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SUITE = os.path.join(HERE, "ref_suite")

# test file -> tests that need the ffmpeg binary (reference tests/test_sliding_window.py:410-505, test_detector_api.py:64-430)
FFMPEG_TESTS = {
    "test_sliding_window.py": ["test_rthk_beep_detection_with_small_chunks", "test_cbs_news_detection_with_multiple_chunks"],
    "test_short_clip.py": [],
    "test_marker_tone_verification.py": [],
    "test_real_data_regressions.py": [],
    "test_detector_api.py": ["test_callback", "test_accumulate_results"],
}


def _write_wav(path, pcm, sr):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


@pytest.fixture(scope="module")
def sample_tree(tmp_path_factory):
    """sample_audios/ as the reference's tests expect it, rebuilt from tests/golden/fixtures.npz + ref_suite/clips."""
    root = tmp_path_factory.mktemp("refsuite")
    fx = np.load(os.path.join(HERE, "golden", "fixtures.npz"))
    for k in fx.files:
        if k.startswith("wav:") and not k.startswith("wav:test_16khz"):
            _write_wav(os.path.join(root, "sample_audios", k[4:]), fx[k], 8000)
    clips = os.path.join(root, "sample_audios", "clips")
    os.makedirs(clips, exist_ok=True)
    # the two WAV clips were PCM16: float32 sample * 32768 is the exact integer again
    for name in ("cbs_news", "天空下的彩虹intro"):
        pcm = np.rint(fx[f"clip8000:{name}"].astype(np.float64) * 32768.0).astype(np.int16)
        _write_wav(os.path.join(clips, name + ".wav"), pcm, 8000)
    for f in os.listdir(os.path.join(SUITE, "clips")):
        shutil.copy(os.path.join(SUITE, "clips", f), os.path.join(clips, f))
    return str(root)


@pytest.mark.parametrize("name", sorted(FFMPEG_TESTS))
def test_reference_suite_passes_on_the_cuda_detector(name, sample_tree):
    args = [sys.executable, os.path.join(HERE, "ref_alias_runner.py"), "-q", "--tb=short", "-p", "no:cacheprovider",
            "--rootdir", sample_tree, os.path.join(SUITE, name)]
    if FFMPEG_TESTS[name]:
        args += ["-k", " and ".join(f"not {t}" for t in FFMPEG_TESTS[name])]
    res = subprocess.run(args, capture_output=True, text=True, cwd=sample_tree, timeout=1800)
    out = res.stdout + res.stderr
    print(out[-3000:])
    assert res.returncode == 0 and " passed" in out and " failed" not in out, out[-6000:]
