"""Dev helper for ncu launch lists: one warm-up scan, then ONE scan of the bench workload between cudaProfilerStart /
cudaProfilerStop (run under `ncu --profile-from-start off ...`).  python tools/scan_only.py [hours]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from audio_pattern_detector_b200 import workloads as W  # noqa: E402
from audio_pattern_detector_b200.audio_clip import AudioClip  # noqa: E402
from audio_pattern_detector_b200.audio_pattern_detector import AudioPatternDetector  # noqa: E402

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
SR, SPC, NP = 8000, 60, 64
pats = W.make_patterns(NP, SR, seed=1)
audio, _ = W.make_stream_device(hours * 3600.0, pats, SR, seed=0, plants_per_pattern=max(1, int(hours)), chunk_seconds=SPC,
                                device="cuda:0")
clips = [AudioClip(name=p["name"], audio=p["audio"], sample_rate=SR, strategy=p["strategy"],
                   strategy_params=p["strategy_params"]) for p in pats]
sys.stderr = open(os.devnull, "w")
det = AudioPatternDetector(audio_clips=clips, seconds_per_chunk=SPC, target_sample_rate=SR, device=0, max_batch_chunks=48)
det.scan_array(audio)
torch.cuda.synchronize()
torch.cuda.profiler.start()
res = det.scan_array(audio)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("detections", sum(len(v) for v in res.peak_times.values()))
