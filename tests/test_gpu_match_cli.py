"""End-to-end drop-in surface on the GPU: match_pattern() and the `match` CLI with WAV files written from the
committed fixtures, checked against what the unmodified reference produced (tests/golden/fixture_runs.json) and its
published goldens (reference tests/test_integration_matching.py:16-26)."""
import json
import os
import subprocess
import sys
import wave

import numpy as np
import pytest

from tests.golden_util import fixtures, load_json

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = {r["wav"]: r for r in load_json("fixture_runs.json") if r["sr"] == 8000}


def write_wav(path, pcm16, sr=8000):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.asarray(pcm16, dtype="<i2").tobytes())


def clip_pcm(name):
    a = fixtures()[f"clip8000:{name}"]                      # float32 = int16 / 32768 exactly (reference loader)
    pcm = np.round(a * 32768.0).astype(np.int16)
    assert np.array_equal(pcm.astype(np.float32) / 32768.0, a)
    return pcm


@pytest.fixture()
def files(tmp_path):
    out = {}
    for wav in ("cbs_news_audio_section.wav", "am1430_section_with_rainbow_intro.wav"):
        p = tmp_path / wav
        write_wav(p, fixtures()["wav:" + wav])
        out[wav] = str(p)
    clips = tmp_path / "clips"
    clips.mkdir()
    for name in ("cbs_news", "天空下的彩虹intro"):
        write_wav(clips / f"{name}.wav", clip_pcm(name))
    out["clips"] = str(clips)
    return out


def test_match_pattern_matches_reference_goldens(files):
    from audio_pattern_detector_b200.match import match_pattern
    patterns = sorted(os.path.join(files["clips"], f) for f in os.listdir(files["clips"]))
    seen = []
    times, total = match_pattern(files["cbs_news_audio_section.wav"], patterns,
                                 on_pattern_detected=lambda n, t: seen.append((n, t)))
    ref = RUNS["cbs_news_audio_section.wav"]
    assert times["cbs_news"] == ref["timestamps"]["cbs_news"] == [25.89875]        # published golden, exact
    assert times["天空下的彩虹intro"] == []
    assert total == ref["total_time"] and seen == [("cbs_news", 25.89875)]
    times, _ = match_pattern(files["am1430_section_with_rainbow_intro.wav"], patterns)
    assert times["天空下的彩虹intro"] == RUNS["am1430_section_with_rainbow_intro.wav"]["timestamps"]["天空下的彩虹intro"]
    assert abs(times["天空下的彩虹intro"][0] - 13.848) < 1e-3 and times["cbs_news"] == []
    with pytest.raises(ValueError, match="does not exist"):
        match_pattern("/nonexistent.wav", patterns)
    with pytest.raises(ValueError, match="No pattern clips passed"):
        match_pattern(files["cbs_news_audio_section.wav"], [])


def test_cli_jsonl_contract(files):
    cmd = [sys.executable, "-m", "audio_pattern_detector_b200.cli", "match", files["cbs_news_audio_section.wav"],
           "--pattern-folder", files["clips"]]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    events = [json.loads(line) for line in r.stdout.splitlines() if line.strip()]
    assert [e["type"] for e in events] == ["start", "pattern_detected", "end"]       # reference match.py:582,524-565,598
    det = events[1]
    assert det["clip_name"] == "cbs_news" and det["timestamp_ms"] == round(25.89875 * 1000)
    assert isinstance(det["timestamp_formatted"], str)
    assert events[2]["total_time_ms"] == round(RUNS["cbs_news_audio_section.wav"]["total_time"] * 1000)


def test_debug_score_dump(files, tmp_path):
    """--debug writes the reference's per-(chunk, clip) score dump (apd.py:574-580); values against the trace the
    unmodified reference produced on the same fixture."""
    from audio_pattern_detector_b200.match import match_pattern
    patterns = sorted(os.path.join(files["clips"], f) for f in os.listdir(files["clips"]))
    dbg = tmp_path / "dbg"
    match_pattern(files["cbs_news_audio_section.wav"], patterns, debug_mode=True, debug_dir=str(dbg))
    path = dbg / "debug" / "cross_correlation_cbs_news" / "0_00:00:00.txt"
    assert path.exists()
    dump = json.loads(path.read_text())
    ref = [u for u in RUNS["cbs_news_audio_section.wav"]["units"] if u["clip"] == "cbs_news" and u["peaks"]][0]
    cand = ref["cands"][0]
    assert dump["peaks"] == ref["peaks"] and dump["seconds"] == [ref["peaks"][0] / 8000]
    sim, parts, pearson = dump["similarities"][0]
    assert abs(parts["whole"] - cand["similarity_whole"]) <= 1e-4 * cand["similarity_whole"]
    assert abs(parts["middle"] - cand["similarity_middle"]) <= 1e-4 * cand["similarity_middle"]
    assert sim == min(parts["whole"], parts["middle"])
    assert abs(pearson["pearson_r"] - cand["pearson"][1]) < 1e-4 and pearson["pearson_r"] == pearson["pearson_w4_6"]
    assert (pearson["best_window_left"], pearson["best_window_right"]) == (4.0, 6.0)
    assert set(pearson) == {"pearson_r", "best_window_left", "best_window_right", "pearson_w0_5", "pearson_w4_6",
                            "pearson_w5_10"}
    assert not (dbg / "debug" / "cross_correlation_天空下的彩虹intro").exists()      # no peaks -> no file
