"""Dev helper: throughput of the file-based drop-in path (match._WavFileStreamWrapper -> find_clip_in_audio)
on a synthetic 16-bit WAV: python tools/wav_throughput.py [hours] [wav rate, default 8000]

With a WAV rate other than 8000 the stream is resampled per chunk read: on the device in the "pcm" mode (row N2),
on the host (numpy, what the reference does) in the "float" mode."""
import os
import sys
import tempfile
import time
import wave

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_pattern_detector_b200 import workloads as W                       # noqa: E402
from audio_pattern_detector_b200.audio_clip import AudioClip, AudioStream    # noqa: E402
from audio_pattern_detector_b200.audio_pattern_detector import AudioPatternDetector  # noqa: E402
from audio_pattern_detector_b200.match import _WavFileStreamWrapper          # noqa: E402

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
wav_rate = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
sr = 8000
pats = W.make_patterns(64, sr, seed=1)
audio, plants = W.make_stream_device(hours * 3600, pats, sr, seed=0, plants_per_pattern=max(1, int(hours)), device="cuda")
if wav_rate != sr:
    from audio_pattern_detector_b200.resample import resample
    C_ = 60 * sr
    full = audio.numel() // C_
    audio = resample(audio[:full * C_].view(full, C_), 60 * wav_rate).reshape(-1)     # chunk-wise up-sampling
pcm = (audio * 32768.0).round_().clamp_(-32768, 32767).to(torch.int16).cpu().numpy()
path = os.path.join(tempfile.gettempdir(), "apd_b200_stream.wav")
with wave.open(path, "wb") as w:
    w.setnchannels(1)
    w.setsampwidth(2)
    w.setframerate(wav_rate)
    w.writeframes(pcm.tobytes())
clips = [AudioClip(name=p["name"], audio=p["audio"], sample_rate=sr, strategy=p["strategy"],
                   strategy_params=p["strategy_params"]) for p in pats]
sys.stderr = open(os.devnull, "w")
det = AudioPatternDetector(audio_clips=clips, seconds_per_chunk=60, target_sample_rate=sr, max_batch_chunks=48)
for mode in ("pcm", "float") if wav_rate == sr or hours <= 2 else ("pcm",):
    for rep in range(2):
        wr = _WavFileStreamWrapper(path, sr)
        if mode == "float":
            wr.pcm_format = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        times, total = det.find_clip_in_audio(AudioStream(name="f", audio_stream=wr, sample_rate=sr))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        wr.close()
    n = sum(len(v) for v in times.values())
    print(f"{mode:5s} path: {hours / dt:7.2f} audio-hours/s ({dt * 1000:.0f} ms for {hours:g} h, {n} detections, {len(plants)} planted)")
os.remove(path)
