"""Builds libapd_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the
GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["fft4.cu", "corr_inv.cu", "loudness.cu", "peaks.cu", "verify.cu", "pcm.cu", "resample.cu", "api.cu"]
LIB = os.path.join(HERE, "libapd_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


STAMP = os.path.join(HERE, "build", "sources.sha256")


def _newer(a: str, b: str) -> bool:
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def _fingerprint(files: list[str]) -> str:
    """Content hash of every source, header and compile flag: an in-tree .so built from other sources is never reused
    silently, whatever the file times say (a checkout or a copy can reset them)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(files):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "apd_b200.h"))
    fp = _fingerprint([os.path.join(CSRC, src) for src in SOURCES] + headers)
    try:
        with open(STAMP) as fh:
            stale = fh.read().strip() != fp
    except OSError:
        stale = True                      # no record of what the objects were built from
    force = force or stale
    objs = []
    relink = force
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            relink = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            sys.stderr.write(out.decode())
    if relink or not os.path.exists(LIB):
        # -cudart shared: the library binds to the libcudart.so.12 of the image instead of carrying a private static copy
        subprocess.run([nvcc, "-shared", "-cudart", "shared", "-o", LIB, *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"], check=True)
        with open(STAMP, "w") as fh:
            fh.write(fp + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
