"""GPU-vs-oracle comparison used by the -m gpu parity tests (and smoke())."""
from __future__ import annotations

import math

import numpy as np

from oracle.detector import OracleDetector
from tests.golden_util import peaks_match, rel_close

SCORE_REL = 1e-4          # BASELINE.json north_star: scores within 1e-4 relative of the float64 CPU path


def make_detector(clips, sr, spc, height_min=None, max_batch_chunks=None):
    from audio_pattern_detector_b200.audio_clip import AudioClip
    from audio_pattern_detector_b200.audio_pattern_detector import AudioPatternDetector
    acs = [AudioClip(name=c["name"], audio=c["audio"], sample_rate=sr, strategy=c.get("strategy"),
                     strategy_params=c.get("strategy_params") or {}) for c in clips]
    return AudioPatternDetector(audio_clips=acs, seconds_per_chunk=spc, target_sample_rate=sr,
                                height_min=height_min, max_batch_chunks=max_batch_chunks)


def close(a, b, rel=SCORE_REL, abs_=1e-7):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, float) and math.isnan(a):
        return isinstance(b, float) and math.isnan(b)
    return rel_close(float(a), float(b), rel, abs_)


def compare_with_oracle(clips, audio, sr, spc, height_min=None, max_batch_chunks=None, max_tie_frac=0.1):
    """Runs the CUDA path and the oracle on the same input; asserts parity; returns a summary dict."""
    det = make_detector(clips, sr, spc, height_min, max_batch_chunks)
    res = det.scan_array(audio, collect_trace=True)
    ora = OracleDetector(clips, sr, spc, height_min=height_min, precision="f64", keep_corr=True)
    assert ora.spc == det.seconds_per_chunk
    units = {}
    otimes, oevents, ototal = ora.run(audio, on_unit=lambda i, st, tr: units.__setitem__((i, st.name), tr))
    assert res.total_time == ototal
    lengths = {s.name: s.length for s in ora.states}
    assert set(res.unit_trace) == set(units)

    gpu_cands = {}
    for c in res.candidates:
        gpu_cands.setdefault((c.chunk, c.clip), []).append(c)

    tie_units = set()
    n_cand = n_accept = 0
    worst = 0.0
    for key, tr in units.items():
        g = res.unit_trace[key]
        assert g["n_out"] == tr["n_out"], key
        # loudness: same recurrence re-associated, float64
        if math.isinf(tr["lufs"]):
            assert g["lufs"] == tr["lufs"], (key, g["lufs"], tr["lufs"])
        else:
            assert abs(g["lufs"] - tr["lufs"]) < 1e-6, (key, g["lufs"], tr["lufs"])
        assert close(g["absmax"], tr["absmax"], 1e-4, 1e-6), (key, g["absmax"], tr["absmax"])
        if tr["absmax"] > 0:
            worst = max(worst, abs(g["absmax"] - tr["absmax"]) / tr["absmax"])
        gp = [c.peak for c in gpu_cands.get(key, [])]
        ok, exact = peaks_match(gp, tr["peaks"], tr["corr"], lengths[key[1]])
        assert ok, (key, gp, tr["peaks"])
        if g["n_peaks"] >= 0:
            assert g["n_peaks"] == len(gp)
        if not exact:
            tie_units.add(key)
            continue
        for gc, oc in zip(gpu_cands.get(key, []), tr["candidates"]):
            n_cand += 1
            assert gc.peak == oc["peak"]
            if oc["kind"] == "skipped":
                assert gc.skipped and not gc.accept, key
                continue
            assert not gc.skipped, key
            assert gc.kind == oc["kind"], (key, gc.kind, oc["kind"])
            assert close(gc.height, oc["height"]), (key, gc.height, oc["height"])
            if oc["kind"] == "tone":
                for seg, got in zip(("match", "left", "right"), gc.tone):
                    want = oc[seg]
                    names = ("detected_frequency", "overall_band_purity", "active_frame_ratio",
                             "longest_active_run", "active_frame_mean_purity")
                    for nm, gv in zip(names, got):
                        assert close(gv, want[nm], SCORE_REL, 1e-6), (key, seg, nm, gv, want[nm])
            else:
                assert close(gc.similarity_whole, oc["similarity_whole"], SCORE_REL, 1e-7), (key, gc, oc)
                assert close(gc.similarity_middle, oc["similarity_middle"], SCORE_REL, 1e-7), (key, gc, oc)
                if oc["pearson"] is not None:
                    for gv, ov in zip(gc.pearson, oc["pearson"]):
                        assert abs(gv - ov) < 1e-4, (key, gc.pearson, oc["pearson"])
                else:
                    assert all(math.isnan(v) for v in gc.pearson), (key, gc.pearson)
            assert gc.accept == oc["accept"], (key, gc, oc)
            n_accept += int(gc.accept)
    assert len(tie_units) <= max(2, int(len(units) * max_tie_frac)), tie_units
    tie_clips = {c for _, c in tie_units}
    for name, ts in otimes.items():
        if name not in tie_clips:
            assert res.peak_times[name] == ts, (name, res.peak_times[name], ts)   # identical floats
    if not tie_units:
        assert [(t, n) for t, n in res.events] == [(t, n) for t, n, _, _ in oevents]
    return {"units": len(units), "candidates": n_cand, "accepted": n_accept, "tie_units": len(tie_units),
            "worst_absmax_rel": worst, "result": res, "oracle_times": otimes}
