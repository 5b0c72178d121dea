"""B200-native implementation of audio_pattern_detector's two-step detection hot path.

Public surface mirrors the reference package: ``AudioPatternDetector``, ``AudioClip``,
``AudioStream``, ``match_pattern`` and the ``audio-pattern-detector`` CLI (``python -m
audio_pattern_detector_b200.cli``).  Importing the package does not need a GPU; constructing a
detector does.
"""
from .audio_clip import AudioClip, AudioStream  # noqa: F401

__all__ = ["AudioClip", "AudioStream", "AudioPatternDetector", "match_pattern"]


def __getattr__(name: str):
    if name == "AudioPatternDetector":
        from .audio_pattern_detector import AudioPatternDetector
        return AudioPatternDetector
    if name == "match_pattern":
        from .match import match_pattern
        return match_pattern
    raise AttributeError(name)
