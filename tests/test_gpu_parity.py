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
    print(f"tie_units={out['tie_units']} of {out['units']} units")
    # no real fixture has rounding-level ties between competing peaks (checked against the float64 oracle on the
    # CPU): the tie allowance of golden_util.peaks_match must never be what makes a fixture pass
    assert out["tie_units"] == 0
    # and the unmodified reference's own result on this fixture (tests/golden/fixture_runs.json)
    assert out["result"].peak_times == run["timestamps"]
    assert [[n, t] for t, n in out["result"].events] == run["events"]


@pytest.mark.parametrize("run", SYN_RUNS, ids=lambda r: r["case"]["id"])
def test_synthetic_streams_match_oracle(run):
    clips, audio = synthetic_inputs(run)
    case = run["case"]
    out = compare_with_oracle(clips, audio, case["sr"], case["spc"], case.get("height_min"), max_batch_chunks=4)
    print(f"tie_units={out['tie_units']} of {out['units']} units")
    # only the exactly periodic tone-in-silence cases (beeps_*) have float32 ties between neighbouring maxima of the
    # same height; every other stream must match peak for peak
    if not case["id"].startswith("beeps_"):
        assert out["tie_units"] == 0
    if out["tie_units"] == 0:
        assert out["result"].peak_times == run["timestamps"]


def test_batching_does_not_change_results():
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    ref = None
    n_chunks = -(-len(audio) // (8000 * 10))
    for b in (1, 3, 16):
        det = make_detector(clips, 8000, 10, max_batch_chunks=b)
        det.work_counters(reset=True)
        res = det.scan_array(audio)
        got = (res.peak_times, res.events)
        if ref is None:
            ref = got
        assert got == ref
        # the phase-2 work counters bench.py reports: one record per candidate, one sub-batch per b chunks
        work = det.work_counters()
        assert work["candidate_records"] == len(res.records) == res.n_candidates
        assert -(-n_chunks // b) <= work["sub_batches"] <= n_chunks      # host input is scanned in segments
        assert 0 < work["selected_units"] <= n_chunks * len(clips)
        assert work["tone_items"] == int(((res.records["flags"] >> 2) == 2).sum())
        assert det.work_counters() == {k: 0 for k in work}


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


def test_slab_scan_equals_full_scan():
    """Sharded use: a rank scans a chunk range of a slab (with look-back halo) of a longer stream."""
    from audio_pattern_detector_b200 import sharding
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    det = make_detector(clips, 8000, 10, max_batch_chunks=2)
    C_ = det._chunk_samples
    n_chunks = (audio.size + C_ - 1) // C_
    full = det.scan_array(audio)
    shards = []
    for rank in range(3):
        c0, c1 = sharding.chunk_range_for_rank(n_chunks, 3, rank)
        lo, hi = sharding.slab_bounds(c0, c1, C_, det._max_halo, audio.size)
        res = det.scan_array(audio[lo:hi], chunk_range=(c0, c1), base_sample=lo, total_samples=audio.size)
        shards.append((res.peak_times, res.events))
    times, events = sharding.merge_shards(shards)
    assert times == full.peak_times and events == full.events
    assert sum(len(v) for v in times.values()) > 0


def test_host_tensor_input_equals_device_input():
    """scan_array from a (pinned) host tensor copies segment-wise on a copy stream, overlapped with the scan."""
    import torch
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    det = make_detector(clips, 8000, 10, max_batch_chunks=1)      # 8 chunks per segment -> several segments
    dev = torch.from_numpy(audio).cuda()
    want = det.scan_array(dev)
    host = torch.from_numpy(audio).pin_memory()
    got = det.scan_array(host)
    assert got.peak_times == want.peak_times and got.events == want.events
    assert det.scan_array(audio).peak_times == want.peak_times          # numpy input takes the same path


def test_planted_patterns_found_at_bench_scale():
    """Size-independent property on the bench workload shape (64 patterns of 0.3-10 s, 60 s chunks, device-
    generated stream too large for the oracle): every planted clip is reported at its planted sample
    (reference timestamps are one sample early, apd.py:439-456), batching does not change the result."""
    import torch
    from audio_pattern_detector_b200 import workloads as W
    sr, spc = 8000, 60
    pats = W.make_patterns(64, sr, seed=1)
    audio, plants = W.make_stream_device(1800.0, pats, sr, seed=3, plants_per_pattern=1, chunk_seconds=spc,
                                         device="cuda")
    det = make_detector(pats, sr, spc, max_batch_chunks=8)
    res = det.scan_array(audio)
    found = {(n, int(round(t * sr))) for n, ts in res.peak_times.items() for t in ts}
    missing = [(n, s0) for n, s0, _g in plants if (n, max(s0 - 1, 0)) not in found]
    assert not missing, missing
    # a clip inside a chunk's look-back can be reported by both chunks (the reference allows such duplicates)
    assert sum(len(v) for v in res.peak_times.values()) <= len(plants) + 8
    det2 = make_detector(pats, sr, spc, max_batch_chunks=3)
    res2 = det2.scan_array(audio.cpu().numpy())
    assert res2.peak_times == res.peak_times and res2.events == res.events


def test_16khz_auto_chunk_many_patterns():
    """BASELINE configs[3] shape at reduced length: 16 kHz, many patterns (0.3-10 s), auto chunk seconds
    (2 * ceil(longest clip) = 20 s): planted clips are found at their planted samples."""
    from audio_pattern_detector_b200 import workloads as W
    sr = 16000
    pats = W.make_patterns(40, sr, seed=7)
    audio, plants = W.make_stream_device(600.0, pats, sr, seed=11, plants_per_pattern=1, chunk_seconds=20, device="cuda")
    det = make_detector(pats, sr, None, max_batch_chunks=6)
    assert det.seconds_per_chunk == 20
    res = det.scan_array(audio)
    found = {(n, int(round(t * sr))) for n, ts in res.peak_times.items() for t in ts}
    missing = [(n, s0) for n, s0, _g in plants if (n, max(s0 - 1, 0)) not in found]
    assert not missing, missing
    assert sum(len(v) for v in res.peak_times.values()) <= len(plants) + 8


def test_three_set_pipeline_matches(monkeypatch):
    """APD_B200_SETS=3: the tone verification of a sub-batch overlaps the normal phase 2 of the next one (records and
    tone work items per set).  Same candidates, scores and order as the default two-set pipeline, with tone clips
    in the mix and one chunk per sub-batch so that every hand-over happens many times."""
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    assert any(c.get("strategy") == "marker_tone" for c in clips)
    want = make_detector(clips, 8000, 10, max_batch_chunks=1).scan_array(audio, collect_trace=True)
    monkeypatch.setenv("APD_B200_SETS", "3")
    got = make_detector(clips, 8000, 10, max_batch_chunks=1).scan_array(audio, collect_trace=True)
    assert got.peak_times == want.peak_times and got.events == want.events
    assert got.records.tobytes() == want.records.tobytes()
    assert got.unit_trace == want.unit_trace


@pytest.mark.parametrize("bad", [float("nan"), float("inf"), -float("inf")])
def test_non_finite_samples(bad):
    """A float stream with a NaN / Inf sample (a corrupt float32 WAV): the K-weighting filter state stays non-finite
    for the rest of the section, every block fails the gates, the gain becomes infinite and the section saturates --
    whatever the reference's arithmetic gives (lib.rs:128-227, apd.py:489-490), the device gives the same, and the
    chunks after the look-back has left the bad sample behind are unaffected."""
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    audio = audio.copy()
    audio[10 * 8000 + 4321] = bad                                        # inside chunk 1
    out = compare_with_oracle(clips, audio, 8000, 10, max_batch_chunks=4)
    assert out["accepted"] > 0


def test_tone_known_answer_signals_on_device():
    """The reference's verifier known answers (tests/test_marker_tone_verification.py:77-95: clean tone accepted,
    harmonic stack and 920 -> 1160 Hz sweep rejected), through the device tone verifier (apd_verify_tone) and the
    reference-named shim; metrics within 1e-4 of the values recorded from the unmodified reference
    (tests/golden/tone_kat.json)."""
    kat = load_json("tone_kat.json")
    run = [r for r in FIXTURE_RUNS if r["wav"] == "rthk_section_with_beep.wav" and r["sr"] == 8000][0]
    clips = [c for c in fixture_clips(run) if c["name"] == "rthk_beep"]
    assert clips and clips[0]["strategy"] == "marker_tone"
    det = make_detector(clips, 8000, 60)
    L = len(clips[0]["audio"])
    accepts = []
    names = ("detected_frequency", "overall_band_purity", "active_frame_ratio", "longest_active_run",
             "active_frame_mean_purity")
    for name, sig in kat["signals"].items():
        x = np.array(sig["samples"], dtype=np.float32)
        assert x.size == L
        ok, m = det.tone_candidate_metrics("rthk_beep", x, L - 1)
        for nm, got in zip(names, m[0]):
            want = sig["metrics"][nm]
            assert abs(got - want) <= 1e-4 * max(abs(want), 1e-2), (name, nm, got, want)
        assert ok == sig["accept"], name
        accepts.append(det._verify_marker_tone(clip_name="rthk_beep", audio_section=x, peak=L - 1, clip_length=L,
                                               dominant_frequency=kat["f0"], sr=8000, section_ts="00:00:00"))
    assert accepts == [True, False, False]


def test_phase2_normalised_maximum_is_exactly_one():
    """Phase 2 re-computes a selected unit's correlation with the arithmetic that produced its phase-1 maximum, so the
    normalised correlation (apd.py:492-494) peaks at exactly 1.0f whenever the unit's maximum is the divisor."""
    import ctypes as C
    import torch
    from audio_pattern_detector_b200 import _lib
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c60"][0]
    clips, audio = synthetic_inputs(run)
    det = make_detector(clips, 8000, 60, max_batch_chunks=4)
    res = det.scan_array(audio, collect_trace=True)
    dev = torch.from_numpy(audio).cuda()
    L_ = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n_chunks = (audio.size + det._chunk_samples - 1) // det._chunk_samples
    checked = 0
    _lib.check(L_.apd_stage_loudness(det._ctx, C.c_void_p(dev.data_ptr()), 0, audio.size, 0, min(4, n_chunks), st), "stage")
    _lib.check(L_.apd_stage_forward_fft(det._ctx, st), "stage")
    _lib.check(L_.apd_stage_correlate_max(det._ctx, st), "stage")
    buf = np.empty(det._chunk_samples + det._max_halo + max(det._clip_lengths), dtype=np.float32)
    for (ci, name), tr in res.unit_trace.items():
        if ci >= min(4, n_chunks) or not tr["absmax"] >= tr["max_choose"] or tr["absmax"] <= 0:
            continue
        idx = [c["name"] for c in clips].index(name)
        n = C.c_int32()
        _lib.check(L_.apd_stage_unit_correlation(det._ctx, ci, idx, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size,
                                                 C.byref(n), st), "unit_correlation")
        assert float(buf[:n.value].max()) == 1.0, (ci, name, float(buf[:n.value].max()))
        checked += 1
    assert checked > 0


def test_dense_candidates_overflow_is_retried_not_fatal():
    """A 0.15 s marker-tone clip inside minutes of steady tone keeps ~N_out / L peaks per unit: more records per
    sub-batch than the device lists hold on average.  The scan must split the range and finish (it used to
    raise mid-stream), with the same candidates as a chunk-by-chunk scan."""
    sr = 8000
    t = np.arange(int(0.15 * sr)) / sr
    beep = (0.8 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    clips = [{"name": "pip", "audio": beep, "strategy": "marker_tone",
              "strategy_params": {"dominant_frequency_hz": 1000.0}}]
    n = 900 * sr
    audio = (0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / sr)).astype(np.float32)
    big = make_detector(clips, sr, 10, max_batch_chunks=90).scan_array(audio)
    assert big.n_candidates > 4096
    one = make_detector(clips, sr, 10, max_batch_chunks=1).scan_array(audio)
    assert big.records.tobytes() == one.records.tobytes()
    assert big.peak_times == one.peak_times


def test_explicit_zero_height_is_honoured():
    """height_min=0.0 is a value, not "unset" (reference :520 substitutes 0.25 only for None): every local maximum
    becomes a candidate, as in the oracle."""
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c4_lowheight"][0]
    clips, audio = synthetic_inputs(run)
    out = compare_with_oracle(clips[:2], audio[:8 * 8000], 8000, 4, height_min=0.0, max_batch_chunks=2)
    low = compare_with_oracle(clips[:2], audio[:8 * 8000], 8000, 4, height_min=None, max_batch_chunks=2)
    assert out["candidates"] > low["candidates"]


def test_short_reads_are_refilled():
    """A raw source that returns fewer bytes than asked for: the chunk is refilled, nothing is dropped."""
    import io
    from audio_pattern_detector_b200.audio_clip import AudioStream

    class Dribble(io.RawIOBase):
        def __init__(self, data):
            self.b, self.k = io.BytesIO(data), 0

        def read(self, n=-1):
            self.k += 1
            return self.b.read(min(n, 4 * (3001 + 7 * (self.k % 11))))

    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    det = make_detector(clips, 8000, 10, max_batch_chunks=3)
    want = det.scan_array(audio)
    times, total = det.find_clip_in_audio(AudioStream(name="s", audio_stream=Dribble(audio.tobytes()), sample_rate=8000))
    assert times == want.peak_times and total == want.total_time


def test_sharded_archive_generator_is_world_size_independent():
    """bench.py --workload c5: every rank generates its slab of ONE logical stream; any cut gives the same samples."""
    import torch
    from audio_pattern_detector_b200 import workloads as W
    sr, spc = 8000, 60
    pats = W.make_patterns(8, sr, seed=1)
    secs = 1500.0                                   # 12 M samples: three noise blocks, plants across block borders
    full, plants = W.make_stream_slab_device(secs, pats, sr, 0, 4, spc, 0, int(secs * sr), device="cuda")
    assert len(plants) > 0
    n = full.numel()
    for lo, hi in ((0, n // 3), (n // 3 - 80000, 2 * n // 3), (4194304 - 17, 4194304 + 900001), (n - 123457, n)):
        part, plants2 = W.make_stream_slab_device(secs, pats, sr, 0, 4, spc, lo, hi, device="cuda")
        assert plants2 == plants
        assert torch.equal(part, full[lo:hi])


def test_tone_flank_gate_does_not_change_decisions():
    """Without a trace the flank transforms of tone candidates whose matched segment already fails are skipped
    (verify.cu: k_tone_gate) and the flanks' frame metrics, which the decision never reads, are not computed; the accept
    decisions, timestamps, the matched-segment metrics and the flank frequency / band purity of the survivors are
    those of the full computation."""
    run = [r for r in FIXTURE_RUNS if r["wav"].endswith("radio1_2026-04-06_12_to_13_28m49_opening.wav") and r["sr"] == 8000][0]
    clips, audio = fixture_clips(run), fixture_audio(run)
    full = make_detector(clips, 8000, 60).scan_array(audio, collect_trace=True)
    fast = make_detector(clips, 8000, 60).scan_array(audio)
    assert fast.peak_times == full.peak_times and fast.events == full.events
    a, b = full.records, fast.records
    assert a.shape == b.shape and (a["flags"] == b["flags"]).all() and (a["peak"] == b["peak"]).all()
    tone = (a["flags"] >> 2) == 2
    assert tone.any()
    assert np.array_equal(a["tone"][tone][:, 0, :], b["tone"][tone][:, 0, :])
    kept = tone & ((b["tone"][:, 1, :] != 0).any(axis=1) | (b["tone"][:, 2, :] != 0).any(axis=1))
    assert np.array_equal(a["tone"][kept][:, :, :2], b["tone"][kept][:, :, :2])
    assert not b["tone"][tone][:, 1:, 2:].any()
    accepted = tone & ((a["flags"] & 1) != 0)
    assert accepted.any() and not (accepted & ~kept).any()


def test_declared_zero_hz_is_a_tone_frequency_not_none():
    """`dominant_frequency_hz = 0.0` is a number: the reference (:217-221, :605-620) sends such a clip's candidates to the
    marker-tone verifier with f0 = 0 Hz (only `None` falls through to the normal verifier).  The C ABI says "none" with
    NaN; 0 Hz must reach the device tone verifier and agree with the oracle."""
    sr = 8000
    rng = np.random.default_rng(3)
    t = np.arange(int(0.6 * sr)) / sr
    beep = (0.7 * np.sin(2 * np.pi * 950.0 * t)).astype(np.float32)
    audio = (0.02 * rng.standard_normal(30 * sr)).astype(np.float32)
    for at in (3.2, 9.9, 17.35):
        audio[int(at * sr):int(at * sr) + beep.size] += beep
    for f0 in (0.0, 950.0):
        clips = [{"name": "beep", "audio": beep, "strategy": "marker_tone",
                  "strategy_params": {"dominant_frequency_hz": f0}}]
        out = compare_with_oracle(clips, audio, sr, 10, max_batch_chunks=2)
        assert out["candidates"] >= 3 and out["tie_units"] == 0
        assert all(c.kind == "tone" for c in out["result"].candidates)
        assert out["accepted"] == (3 if f0 else 0)


def test_22050_hz_detector_matches_oracle():
    """Any rate whose 100 ms gating hop is a whole number of samples is supported (here 2 205 = 21 cells of 105
    samples); the generic four-step shapes serve its section sizes."""
    sr = 22050
    rng = np.random.default_rng(11)
    t = np.arange(int(0.8 * sr)) / sr
    chirp = (0.5 * np.sin(2 * np.pi * (400.0 + 900.0 * t) * t)).astype(np.float32)
    jingle = (0.3 * rng.standard_normal(int(1.3 * sr))).astype(np.float32)
    audio = (0.03 * rng.standard_normal(25 * sr)).astype(np.float32)
    for at, clip in ((2.5, chirp), (9.7, jingle), (14.2, chirp), (19.6, jingle)):
        audio[int(at * sr):int(at * sr) + clip.size] += clip
    clips = [{"name": "chirp", "audio": chirp}, {"name": "jingle", "audio": jingle}]
    out = compare_with_oracle(clips, audio, sr, 10, max_batch_chunks=2)
    assert out["accepted"] == 4 and out["tie_units"] == 0


def test_11025_hz_general_rate_loudness_matches_oracle():
    """0.1 * 11 025 is not a whole number of samples: the gating blocks' truncated bounds (lib.rs:113-122) fall on no
    cell grid, and the loudness takes the serial general-rate kernel (k_kw_serial) - same LUFS, gains and detections as
    the oracle, short (< 0.5 s, single-block) clip included."""
    sr = 11025
    rng = np.random.default_rng(12)
    t = np.arange(int(0.9 * sr)) / sr
    chirp = (0.5 * np.sin(2 * np.pi * (300.0 + 800.0 * t) * t)).astype(np.float32)
    blip = (0.4 * rng.standard_normal(int(0.31 * sr))).astype(np.float32)
    audio = (0.03 * rng.standard_normal(int(34.7 * sr))).astype(np.float32)
    audio[5 * sr:6 * sr] = 0.0                                                    # a silent second inside a chunk
    for at, clip in ((2.5, chirp), (9.8, blip), (14.2, chirp), (19.9, blip), (27.3, chirp)):
        audio[int(at * sr):int(at * sr) + clip.size] += clip
    clips = [{"name": "chirp", "audio": chirp}, {"name": "blip", "audio": blip}]
    out = compare_with_oracle(clips, audio, sr, 10, max_batch_chunks=2)
    assert out["accepted"] == 5 and out["tie_units"] == 0
    silent = np.zeros(12 * sr, dtype=np.float32)
    quiet = compare_with_oracle(clips, silent, sr, 10)
    assert quiet["candidates"] == 0


def test_hot_shapes_384_and_448():
    """40 s chunks at 8 kHz put N_out at 330-390 k (N1 = 384 = 8 x 8 x 6) and 410-450 k (N1 = 448 = 8 x 8 x 7): the
    radix-6 / radix-7 last passes of the hot-shape kernels, forward and inverse, both phases."""
    sr = 8000
    rng = np.random.default_rng(21)
    t1 = np.arange(int(1.2 * sr)) / sr
    short = (0.5 * np.sin(2 * np.pi * (300.0 + 700.0 * t1) * t1)).astype(np.float32)          # sw 2 s: N_out 345 599
    long_ = (0.3 * rng.standard_normal(int(6.3 * sr))).astype(np.float32)                    # sw 7 s: N_out 426 399
    tone = (0.6 * np.sin(2 * np.pi * 1040.0 * np.arange(int(2.5 * sr)) / sr)).astype(np.float32)   # sw 3 s: 363 999
    audio = (0.03 * rng.standard_normal(130 * sr)).astype(np.float32)
    for at, clip in ((5.0, short), (39.7, long_), (61.3, tone), (79.9, short), (100.2, long_), (121.0, tone)):
        audio[int(at * sr):int(at * sr) + clip.size] += clip
    clips = [{"name": "short", "audio": short}, {"name": "long", "audio": long_},
             {"name": "tone", "audio": tone, "strategy": "marker_tone",
              "strategy_params": {"dominant_frequency_hz": 1040.0}}]
    det = make_detector(clips, sr, 40)
    assert [det.clip_info(i)["fft_points"] for i in range(3)] == [2 * 384 * 512, 2 * 448 * 512, 2 * 384 * 512]
    out = compare_with_oracle(clips, audio, sr, 40, max_batch_chunks=3)
    assert out["accepted"] == 6 and out["tie_units"] == 0
    assert out["worst_absmax_rel"] < 1e-5


def test_pattern_split_equals_full_scan():
    """Splitting the PATTERN list instead of the chunk range (the alternative partition of SURVEY.md section 8e, for
    short streams against many patterns): two detectors, each over all chunks and half of the clips, merged by
    sharding.merge_pattern_shards, give the single detector's peak times and callback order."""
    from audio_pattern_detector_b200 import sharding
    run = [r for r in SYN_RUNS if r["case"]["id"] == "s8k_c10"][0]
    clips, audio = synthetic_inputs(run)
    full = make_detector(clips, 8000, 10).scan_array(audio)
    assert len(clips) >= 2 and len(full.events) > 0
    k = len(clips) // 2
    shards = [sharding.detections_of(make_detector(part, 8000, 10).scan_array(audio), off)
              for part, off in ((clips[:k], 0), (clips[k:], k))]
    times, events = sharding.merge_pattern_shards(shards, [c["name"] for c in clips])
    assert times == full.peak_times and events == full.events
