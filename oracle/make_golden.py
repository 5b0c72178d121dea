"""Generate tests/golden/ by running the UNMODIFIED reference Python package.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

It imports the reference's own ``AudioPatternDetector`` through oracle/refshim.py
(stand-ins only for the three modules that cannot be installed offline) and
records, for every (chunk x clip) unit, what the reference computed: the
``find_peaks`` result, each candidate's correlation height, the verifier's
scores and its accept decision, and the final timestamps.  Inputs are stored
beside the outputs so the GPU box (which has no /root/reference) can replay
them:

  tests/golden/fixtures.npz          WAV fixtures as int16 + the clip arrays the
                                     reference's loaders produced (float32)
  tests/golden/fixture_runs.json     reference traces on the real fixtures
  tests/golden/synthetic_runs.json   reference traces on workloads.py streams
                                     (inputs are regenerated from seeds; a
                                     checksum guards against generator drift)
  tests/golden/tone_kat.json         marker-tone verifier known answers
"""
from __future__ import annotations

import glob
import hashlib
import io
import json
import os
import sys
import wave

import numpy as np

from . import refshim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
SAMPLES = os.path.join(refshim.REFERENCE_ROOT, "sample_audios")


class Tracer:
    """Wraps reference methods to record what they computed (no behaviour change)."""

    def __init__(self, apd_mod, native_mod):
        self.apd = apd_mod
        self.native = native_mod
        self.units: list[dict] = []
        self.cur: dict | None = None
        self._pearson: list[float] = []
        self._tone = None
        D = apd_mod.AudioPatternDetector
        self._orig = {k: getattr(D, k) for k in ("_correlation_method", "_verify_peak_candidate",
                                                  "_get_peak_times_normal", "_analyze_tone_candidate_context")}
        self._orig_fp = native_mod.find_peaks
        self._orig_pc = native_mod.pearson_correlation
        tr = self

        def corr_method(self_, clip_data, audio_section, index):
            tr.cur = {"chunk": int(index), "clip": clip_data["clip_name"], "section_len": int(len(audio_section)),
                      "peaks": [], "cands": []}
            try:
                return tr._orig["_correlation_method"](self_, clip_data, audio_section=audio_section, index=index)
            finally:
                tr.units.append(tr.cur)
                tr.cur = None

        def find_peaks(data, **kw):
            res = tr._orig_fp(data, **kw)
            if tr.cur is not None:
                tr.cur["peaks"] = [int(p) for p in res[0]]
            return res

        def pearson(x, y):
            r = tr._orig_pc(x, y)
            tr._pearson.append(float(r))
            return r

        def verify(self_, **kw):
            before = len(kw["peaks_final"])
            tr._pearson = []
            tr._tone = None
            rec = {"peak": int(kw["peak"]), "height": float(kw["correlation"][kw["peak"]])}
            tr._rec = rec
            tr._orig["_verify_peak_candidate"](self_, **kw)
            rec["accept"] = len(kw["peaks_final"]) > before
            if tr._tone is not None:
                rec["kind"] = "tone"
                rec["tone"] = tr._tone
            tr.cur["cands"].append(rec)

        def normal(self_, **kw):
            cc, sl = kw["correlation_clip"], kw["correlation_slice"]
            ps = len(cc) // 10
            parts = np.array([apd_mod._mean_squared_error(cc[i * ps:(i + 1) * ps], sl[i * ps:(i + 1) * ps])
                              for i in range(10)], dtype=np.float32)
            tr._rec["kind"] = "short" if kw.get("is_short_clip") else "normal"
            tr._rec["similarity_whole"] = float(np.mean(parts))
            tr._rec["similarity_middle"] = float(np.mean(parts[4:6]))
            tr._orig["_get_peak_times_normal"](self_, **kw)
            tr._rec["pearson"] = list(tr._pearson)

        def tone_ctx(self_, **kw):
            res = tr._orig["_analyze_tone_candidate_context"](self_, **kw)
            tr._tone = [{"detected_frequency": m.detected_frequency, "overall_band_purity": m.overall_band_purity,
                         "active_frame_ratio": m.active_frame_ratio, "longest_active_run": m.longest_active_run,
                         "active_frame_mean_purity": m.active_frame_mean_purity} for m in res]
            return res

        D._correlation_method = corr_method
        D._verify_peak_candidate = verify
        D._get_peak_times_normal = normal
        D._analyze_tone_candidate_context = tone_ctx
        native_mod.find_peaks = find_peaks
        native_mod.pearson_correlation = pearson

    def take(self) -> list[dict]:
        u, self.units = self.units, []
        return u


def _read_wav_int16(path: str) -> tuple[np.ndarray, int]:
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1, path
        return np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).copy(), w.getframerate()


def _clip_meta(c) -> dict:
    return {"name": c.name, "strategy": c.strategy, "strategy_params": c.strategy_params,
            "length": int(len(c.audio))}


def _run_reference(ref, tracer, clips, audio_f32, sr, spc, height_min=None) -> dict:
    from audio_pattern_detector.audio_clip import AudioStream
    from audio_pattern_detector.audio_pattern_detector import AudioPatternDetector
    det = AudioPatternDetector(audio_clips=clips, seconds_per_chunk=spc, target_sample_rate=sr,
                               height_min=height_min)
    events: list = []
    tracer.take()
    times, total = det.find_clip_in_audio(
        AudioStream(name="golden", audio_stream=io.BytesIO(audio_f32.tobytes()), sample_rate=sr),
        on_pattern_detected=lambda name, t: events.append([name, float(t)]))
    return {"seconds_per_chunk": det.seconds_per_chunk, "total_time": float(total),
            "timestamps": {k: [float(x) for x in v] for k, v in times.items()},
            "events": events, "units": tracer.take()}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# synthetic cases replayed by tests/ (inputs regenerated from these parameters)
SYNTHETIC_CASES = [
    {"id": "s8k_c10", "sr": 8000, "seconds": 95.3, "spc": 10, "n_patterns": 10, "min_s": 0.2, "max_s": 4.5,
     "pattern_seed": 11, "stream_seed": 5, "plants": 2},
    {"id": "s8k_c60", "sr": 8000, "seconds": 200.0, "spc": 60, "n_patterns": 16, "min_s": 0.3, "max_s": 10.0,
     "pattern_seed": 1, "stream_seed": 0, "plants": 2},
    {"id": "s16k_auto", "sr": 16000, "seconds": 70.0, "spc": None, "n_patterns": 8, "min_s": 0.3, "max_s": 6.0,
     "pattern_seed": 3, "stream_seed": 7, "plants": 2},
    {"id": "s8k_c4_lowheight", "sr": 8000, "seconds": 30.0, "spc": 4, "n_patterns": 6, "min_s": 0.15, "max_s": 1.9,
     "pattern_seed": 21, "stream_seed": 9, "plants": 3, "height_min": 0.1},
]


def synthetic_inputs(case: dict):
    from audio_pattern_detector_b200 import workloads as W
    pats = W.make_patterns(case["n_patterns"], case["sr"], case["pattern_seed"], case["min_s"], case["max_s"])
    spc = case["spc"]
    if spc is None:
        spc_eff = int(np.ceil(max(p["audio"].size for p in pats) / case["sr"])) * 2
    else:
        spc_eff = spc
    audio, plants = W.make_stream(case["seconds"], pats, case["sr"], case["stream_seed"], case["plants"], spc_eff)
    return pats, audio, plants


def beeps_in_silence(sr: int = 8000):
    """The reference's sliding-window test shape (tests/test_sliding_window.py:19-50): 1 kHz beeps in silence."""
    t = np.linspace(0, 0.23, int(sr * 0.23), endpoint=False)
    beep = np.sin(2 * np.pi * 1000.0 * t).astype(np.float32)
    audio = np.zeros(int(sr * 31.0), dtype=np.float32)
    starts = [1.0, 4.0, 5.9, 14.95, 29.0]
    for s in starts:
        i = int(s * sr)
        audio[i:i + beep.size] = beep
    return beep, audio, starts


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    ref = refshim.install()
    import audio_pattern_detector.audio_pattern_detector as apd_mod
    from audio_pattern_detector.audio_clip import AudioClip
    native_mod = sys.modules["audio_pattern_detector._native"]
    tracer = Tracer(apd_mod, native_mod)

    store: dict[str, np.ndarray] = {}
    fixture_runs: list[dict] = []

    def load_clips(paths, sr):
        clips = [AudioClip.from_audio_file(p, sample_rate=sr) for p in paths]
        for c in clips:
            store[f"clip{sr}:{c.name}"] = np.asarray(c.audio, dtype=np.float32)
        return clips

    clip_paths = sorted(glob.glob(os.path.join(SAMPLES, "clips", "*")))
    clips8 = load_clips(clip_paths, 8000)
    gen_clips = load_clips(sorted(glob.glob(os.path.join(SAMPLES, "test_generated", "clips", "*.wav"))), 8000)

    wavs8 = sorted(glob.glob(os.path.join(SAMPLES, "*.wav"))) + \
        sorted(glob.glob(os.path.join(SAMPLES, "regressions", "*", "*.wav"))) + \
        sorted(glob.glob(os.path.join(SAMPLES, "test_generated", "*.wav")))
    for path in wavs8:
        pcm, sr = _read_wav_int16(path)
        assert sr == 8000, path
        key = os.path.relpath(path, SAMPLES)
        store["wav:" + key] = pcm
        audio = pcm.astype(np.float32) / 32768.0          # reference match.py:407
        clipset = clips8 + (gen_clips if "test_generated" in key else [])
        run = _run_reference(ref, tracer, clipset, audio, 8000, 60)
        run.update({"wav": key, "sr": 8000, "clips": [_clip_meta(c) for c in clipset]})
        fixture_runs.append(run)
        print(key, {k: v for k, v in run["timestamps"].items() if v}, file=sys.stderr)

    # native 16 kHz runs (target_sample_rate=16000; .apd.toml clips re-synthesised at 16 k)
    clip_paths16 = sorted(glob.glob(os.path.join(SAMPLES, "test_16khz", "clips", "*.wav"))) + \
        [p for p in clip_paths if p.endswith(".apd.toml") and "base64" not in p]
    clips16 = load_clips(clip_paths16, 16000)
    for path in sorted(glob.glob(os.path.join(SAMPLES, "test_16khz", "*.wav"))):
        pcm, sr = _read_wav_int16(path)
        assert sr == 16000
        key = os.path.relpath(path, SAMPLES)
        store["wav:" + key] = pcm
        audio = pcm.astype(np.float32) / 32768.0
        for spc in (60, None):
            run = _run_reference(ref, tracer, clips16, audio, 16000, spc)
            run.update({"wav": key, "sr": 16000, "requested_spc": spc, "clips": [_clip_meta(c) for c in clips16]})
            fixture_runs.append(run)
            print(key, spc, {k: v for k, v in run["timestamps"].items() if v}, file=sys.stderr)

    np.savez_compressed(os.path.join(GOLDEN, "fixtures.npz"), **store)
    with open(os.path.join(GOLDEN, "fixture_runs.json"), "w") as f:
        json.dump(fixture_runs, f, ensure_ascii=False)

    # synthetic multi-chunk streams
    syn_runs = []
    for case in SYNTHETIC_CASES:
        pats, audio, plants = synthetic_inputs(case)
        clips = [AudioClip(name=p["name"], audio=p["audio"], sample_rate=case["sr"], strategy=p["strategy"],
                           strategy_params=p["strategy_params"]) for p in pats]
        run = _run_reference(ref, tracer, clips, audio, case["sr"], case["spc"], case.get("height_min"))
        run.update({"case": case, "audio_sha": sha(audio), "pattern_sha": sha(np.concatenate([p["audio"] for p in pats])),
                    "plants": [[n, int(s), float(g)] for n, s, g in plants]})
        syn_runs.append(run)
        print(case["id"], sum(len(v) for v in run["timestamps"].values()), "detections /", len(plants), "plants",
              file=sys.stderr)
    beep, audio, starts = beeps_in_silence()
    for spc in (3, 10, 60):
        clips = [AudioClip(name="test_beep", audio=beep, sample_rate=8000)]
        run = _run_reference(ref, tracer, clips, audio, 8000, spc)
        run.update({"case": {"id": f"beeps_c{spc}", "sr": 8000, "spc": spc, "kind": "beeps_in_silence"},
                    "audio_sha": sha(audio), "starts": starts})
        syn_runs.append(run)
        print(run["case"]["id"], run["timestamps"], file=sys.stderr)
    with open(os.path.join(GOLDEN, "synthetic_runs.json"), "w") as f:
        json.dump(syn_runs, f)

    # marker-tone verifier KAT (reference tests/test_marker_tone_verification.py:77-95 shapes)
    from audio_pattern_detector.detection_utils import analyze_pure_tone_candidate
    rthk = [c for c in clips8 if c.name == "rthk_beep"][0]
    f0 = float(rthk.strategy_params["dominant_frequency_hz"])
    n = len(rthk.audio)
    t = np.arange(n, dtype=np.float32) / 8000
    env = np.hanning(n).astype(np.float32)
    clean = (0.9 * np.sin(2 * np.pi * f0 * t) * env).astype(np.float32)
    stack = sum(a * np.sin(2 * np.pi * 260.0 * k * t) for k, a in zip(range(1, 6), (0.5, 0.35, 0.3, 0.28, 0.22)))
    stack = (stack.astype(np.float32) * env)
    stack = (stack / np.max(np.abs(stack))).astype(np.float32)
    inst = np.linspace(920.0, 1160.0, n, dtype=np.float32)
    sweep = (0.9 * np.sin(2 * np.pi * np.cumsum(inst) / 8000) * env).astype(np.float32)
    kat = {"f0": f0, "sr": 8000, "signals": {}}
    det = apd_mod.AudioPatternDetector(audio_clips=[rthk])
    for name, sig in (("clean", clean), ("harmonic_stack", stack), ("sweep", sweep)):
        m = analyze_pure_tone_candidate(sig, 8000, f0)
        ok = det._verify_marker_tone(clip_name="rthk_beep", audio_section=sig, peak=n - 1, clip_length=n,
                                     dominant_frequency=f0, sr=8000, section_ts="00:00:00")
        kat["signals"][name] = {"samples": [float(v) for v in sig], "metrics": {
            "detected_frequency": m.detected_frequency, "overall_band_purity": m.overall_band_purity,
            "active_frame_ratio": m.active_frame_ratio, "longest_active_run": m.longest_active_run,
            "active_frame_mean_purity": m.active_frame_mean_purity}, "accept": bool(ok)}
    with open(os.path.join(GOLDEN, "tone_kat.json"), "w") as f:
        json.dump(kat, f)
    print("golden written to", GOLDEN, file=sys.stderr)


if __name__ == "__main__":
    main()
