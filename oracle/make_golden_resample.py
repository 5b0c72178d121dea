"""Generate tests/golden/resample_runs.json by running the UNMODIFIED reference ``match_pattern`` on WAV files whose
sample rate differs from the detector's (row N2: the reference resamples every chunk read, match.py:395-423).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden_resample

Inputs are the fixtures already committed in tests/golden/fixtures.npz (the 16 kHz WAVs as int16, the 8 kHz clips as
the float32 arrays the reference's loaders produced); the reference's _native.resample is the stand-in of
oracle/refshim.py (the Rust crate cannot be built offline).  Recorded: timestamps per clip, callback order, total time.
"""
from __future__ import annotations

import glob
import json
import os

from . import refshim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
SAMPLES = os.path.join(refshim.REFERENCE_ROOT, "sample_audios")


def main() -> None:
    refshim.install()
    from audio_pattern_detector.match import match_pattern          # the reference's
    clips = sorted(glob.glob(os.path.join(SAMPLES, "clips", "*")))
    runs = []
    for wav in ("test_16khz/cbs_news_audio_section_16k.wav", "test_16khz/rthk_section_with_beep_16k.wav"):
        for spc in (60, 10, 8):
            events: list[list] = []
            times, total = match_pattern(os.path.join(SAMPLES, wav), clips, seconds_per_chunk=spc,
                                         target_sample_rate=8000,
                                         on_pattern_detected=lambda n, t: events.append([n, t]))
            runs.append({"wav": wav, "wav_rate": 16000, "sr": 8000, "spc": spc,
                         "clips": [os.path.basename(c) for c in clips],
                         "timestamps": times, "events": events, "total_time": total})
            print(wav, spc, {k: v for k, v in times.items() if v}, total)
    with open(os.path.join(GOLDEN, "resample_runs.json"), "w") as f:
        json.dump(runs, f, ensure_ascii=False, indent=1)


if __name__ == "__main__":
    main()
