"""Dev helper: the fused correlate + max stage timed alone (CUDA events) on the bench workload, nothing else run.
python tools/corr_alone.py [hours] [batch_chunks]   -> ms, roofline fraction of the measured HBM peak"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from audio_pattern_detector_b200 import workloads as W  # noqa: E402
from audio_pattern_detector_b200.audio_clip import AudioClip  # noqa: E402
from audio_pattern_detector_b200.audio_pattern_detector import AudioPatternDetector  # noqa: E402

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 48
SR, SPC, NP = 8000, 60, 64
pats = W.make_patterns(NP, SR, seed=1)
audio, _ = W.make_stream_device(hours * 3600.0, pats, SR, seed=0, plants_per_pattern=max(1, int(hours)), chunk_seconds=SPC,
                                device="cuda:0")
clips = [AudioClip(name=p["name"], audio=p["audio"], sample_rate=SR, strategy=p["strategy"],
                   strategy_params=p["strategy_params"]) for p in pats]
sys.stderr = open(os.devnull, "w")
det = AudioPatternDetector(audio_clips=clips, seconds_per_chunk=SPC, target_sample_rate=SR, device=0, max_batch_chunks=batch)
n = audio.numel()
n_chunks = (n + SPC * SR - 1) // (SPC * SR)
alg = sum(8 * det.unit_n_out(ci, p, n) for ci in range(n_chunks) for p in range(NP))
best = 1e30
for _ in range(3):
    ms, launches = det.time_correlate_stage(audio)
    best = min(best, ms)
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6650.0
print(f"alone_ms={best:.2f} frac={alg / 1e9 / (best / 1e3) / peak:.3f} launches={launches} us_per_unit={best * 1e3 / (n_chunks * NP):.3f}")
