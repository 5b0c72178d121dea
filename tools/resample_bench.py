"""Dev helper: device FFT resampler throughput (row N2): python tools/resample_bench.py [in_rate] [chunks]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_pattern_detector_b200.resample import resample_into, workspace_bytes   # noqa: E402

in_rate = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 48
n, m = 60 * in_rate, 60 * 8000
x = torch.randn(chunks * n, device="cuda") * 0.2
y = torch.empty(chunks * m, device="cuda")
for _ in range(3):
    resample_into(x, n, y, m, chunks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    resample_into(x, n, y, m, chunks)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
hours = chunks * 60 / 3600
print(f"{in_rate} Hz -> 8000 Hz, {chunks} chunks of 60 s: {ms:.2f} ms per batch = {ms / chunks * 1e3:.1f} us per chunk, "
      f"{hours / (ms / 1e3):.0f} audio-hours/s, workspace {workspace_bytes(n, m, 1) / 1e6:.0f} MB per chunk, "
      f"algorithmic {(4 * n + 4 * m) * chunks / ms / 1e6:.0f} GB/s (read f32 in + write f32 out)")
