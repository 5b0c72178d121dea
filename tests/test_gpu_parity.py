"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed goldens.

Bars (BASELINE.json north_star): identical detections (clip names + integer peak sample
indices -> bit-identical timestamps), scores within 1e-4 relative of the float64 CPU path.
"""
import numpy as np
import pytest

from tests.golden_util import fixture_audio, fixture_clips, load_json, synthetic_inputs
from tests.gpu_compare import compare_with_oracle, make_detector

pytestmark = pytest.mark.gpu

FIXTURE_RUNS = load_json("fixture_runs.json")
SYN_RUNS = load_json("synthetic_runs.json")


@pytest.mark.parametrize("run", FIXTURE_RUNS, ids=lambda r: f"{r['wav']}@{r['sr']}/{r.get('requested_spc', 60)}")
def test_fixtures_match_oracle_and_reference_golden(run):
    clips, audio = fixture_clips(run), fixture_audio(run)
    out = compare_with_oracle(clips, audio, run["sr"], run.get("requested_spc", 60))
    if out["tie_units"] == 0:
        # and the unmodified reference's own result on this fixture (tests/golden/fixture_runs.json)
        assert out["result"].peak_times == run["timestamps"]
        assert [[n, t] for t, n in out["result"].events] == run["events"]


@pytest.mark.parametrize("run", SYN_RUNS, ids=lambda r: r["case"]["id"])
def test_synthetic_streams_match_oracle(run):
    clips, audio = synthetic_inputs(run)
    case = run["case"]
    out = compare_with_oracle(clips, audio, case["sr"], case["spc"], case.get("height_min"), max_batch_chunks=4)
    if out["tie_units"] == 0:
        assert out["result"].peak_times == run["timestamps"]


def test_batching_does_not_change_results():
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    ref = None
    for b in (1, 3, 16):
        det = make_detector(clips, 8000, 10, max_batch_chunks=b)
        res = det.scan_array(audio)
        got = (res.peak_times, res.events)
        if ref is None:
            ref = got
        assert got == ref


def test_streaming_api_equals_array_scan():
    import io
    from audio_pattern_detector_b200.audio_clip import AudioStream
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    det = make_detector(clips, 8000, 10, max_batch_chunks=3)
    want = det.scan_array(audio)
    seen = []
    times, total = det.find_clip_in_audio(
        AudioStream(name="s", audio_stream=io.BytesIO(audio.tobytes()), sample_rate=8000),
        on_pattern_detected=lambda n, t: seen.append((t, n)))
    assert times == want.peak_times and total == want.total_time and seen == want.events
    none_times, _ = det.find_clip_in_audio(
        AudioStream(name="s", audio_stream=io.BytesIO(audio.tobytes()), sample_rate=8000), accumulate_results=False)
    assert none_times is None


def test_pattern_side_precompute_matches_oracle():
    from oracle.detector import prepare_clip
    run = FIXTURE_RUNS[0]
    clips = fixture_clips(run)
    det = make_detector(clips, run["sr"], 60)
    for i, c in enumerate(clips):
        st = prepare_clip(c["name"], c["audio"], run["sr"], c["strategy"], c["strategy_params"])
        info = det.clip_info(i)
        assert info["sliding_window"] == st.sliding_window
        assert abs(info["lufs"] - native_lufs(c["audio"], run["sr"])) < 1e-9
        np.testing.assert_allclose(info["normalized"], st.clip, rtol=0, atol=1e-7)
        assert abs(info["self_max"] - float(st.self_max)) <= 1e-5 * float(st.self_max)
        assert np.max(np.abs(info["self_correlation"] - st.self_corr)) < 2e-6


def native_lufs(audio, sr):
    from oracle import native
    secs = len(audio) / sr
    return native.integrated_loudness(np.ascontiguousarray(audio, dtype=np.float32), sr, secs if secs < 0.5 else 0.4)


def test_silence_and_edges():
    sr = 8000
    t = np.arange(int(0.23 * sr)) / sr
    beep = np.sin(2 * np.pi * 1000 * t).astype(np.float32)
    clips = [{"name": "beep", "audio": beep, "strategy": None, "strategy_params": {}}]
    # all-silent stream: gain = inf, 0*inf -> NaN -> 0 ; no detections, lufs = -inf
    out = compare_with_oracle(clips, np.zeros(5 * sr, np.float32), sr, 2)
    assert out["accepted"] == 0
    # stream shorter than the clip, and a one-sample final chunk
    compare_with_oracle(clips, (0.1 * np.random.RandomState(0).randn(1000)).astype(np.float32), sr, 2)
    compare_with_oracle(clips, (0.1 * np.random.RandomState(1).randn(2 * 2 * sr + 1)).astype(np.float32), sr, 2)
