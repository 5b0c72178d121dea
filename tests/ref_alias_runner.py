"""Runs test files of the reference against THIS package: `audio_pattern_detector[.x]` is aliased to
`audio_pattern_detector_b200[.x]` before pytest collects them (tests/test_reference_host_suite.py).
usage: python tests/ref_alias_runner.py <pytest args...>"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_pattern_detector_b200 as pkg                                  # noqa: E402

sys.modules["audio_pattern_detector"] = pkg
for sub in ("audio_clip", "audio_utils", "pattern_config", "detection_utils", "match", "cli", "audio_pattern_detector"):
    sys.modules[f"audio_pattern_detector.{sub}"] = importlib.import_module(f"audio_pattern_detector_b200.{sub}")

import pytest                                                              # noqa: E402

sys.exit(pytest.main(sys.argv[1:]))
