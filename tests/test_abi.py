"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/apd_b200.h declares.

No compute call is made here (that needs a GPU: tests/test_gpu_parity.py)."""
import ctypes
import os
import re

from audio_pattern_detector_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "apd_b200.h")


def declared_symbols() -> list[str]:
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"APD_API\s+[\w\s\*]+?\b(apd_\w+)\s*\(", text)


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) == len(set(syms)) >= 16
    for must in ("apd_create", "apd_destroy", "apd_scan", "apd_last_error", "apd_stage_correlate_max"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/apd_b200.h but not exported by {path}"


def test_ctypes_prototypes_cover_the_header():
    assert sorted(_lib.PROTOTYPES) == sorted(declared_symbols())
    _lib.lib()          # binds restype/argtypes for all of them


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.Candidate) == 176          # 4 int32 + 4 float + 3 double + 15 double
    assert ctypes.sizeof(_lib.UnitTrace) == 16
    assert ctypes.sizeof(_lib.ClipDesc) == 8 + 4 + 4 + 7 * 8


def test_errors_without_a_device_are_loud():
    """No CPU fallback: creating a context on a box without CUDA fails with a status code and message."""
    import torch
    if torch.cuda.is_available():
        return
    import numpy as np
    L = _lib.lib()
    a = np.zeros(100, dtype=np.float32)
    d = (_lib.ClipDesc * 1)()
    d[0].samples = a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    d[0].length = 100
    ctx = ctypes.c_void_p()
    rc = L.apd_create(ctypes.byref(ctx), 0, 8000, 16000, 0.0, 1, d, 1)
    assert rc != 0 and L.apd_last_error()


def test_build_is_stamped_with_a_content_hash_of_its_sources():
    """A stale in-tree .so is never reused silently: build.py records the hash of every source, header and flag it
    compiled and rebuilds when they differ."""
    import os
    from audio_pattern_detector_b200 import build as b
    headers = [os.path.join(b.CSRC, f) for f in os.listdir(b.CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(b.HERE), "include", "apd_b200.h"))
    fp = b._fingerprint([os.path.join(b.CSRC, s) for s in b.SOURCES] + headers)
    b.build()
    with open(b.STAMP) as fh:
        assert fh.read().strip() == fp
