"""Dev helper: cost of the serial general-rate loudness path (sample rates whose 100 ms hop is not a whole number of
samples) against the cell-parallel path at a neighbouring rate.  python tools/rate_path_bench.py [hours]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from audio_pattern_detector_b200 import workloads as W  # noqa: E402
from audio_pattern_detector_b200.audio_clip import AudioClip  # noqa: E402
from audio_pattern_detector_b200.audio_pattern_detector import AudioPatternDetector  # noqa: E402

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
for sr in (11020, 11025):
    pats = W.make_patterns(16, sr, seed=1)
    audio, _ = W.make_stream_device(hours * 3600.0, pats, sr, seed=0, plants_per_pattern=2, chunk_seconds=60, device="cuda:0")
    clips = [AudioClip(name=p["name"], audio=p["audio"], sample_rate=sr, strategy=p["strategy"],
                       strategy_params=p["strategy_params"]) for p in pats]
    det = AudioPatternDetector(audio_clips=clips, seconds_per_chunk=60, target_sample_rate=sr, device=0, max_batch_chunks=48)
    det.scan_array(audio)
    det.enable_profiling(True)
    det.stage_times_ms(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = det.scan_array(audio)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    st = {k: round(v, 1) for k, v in det.stage_times_ms().items()}
    print(f"sr={sr}: {hours} h x 16 patterns in {ms:.1f} ms, stages {st}, detections {sum(len(v) for v in res.peak_times.values())}")
