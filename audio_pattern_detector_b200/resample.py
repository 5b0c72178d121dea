"""Device FFT resampler: host binding of ``apd_resample``.

Drop-in for ``audio_pattern_detector._native.resample(data, num_samples)`` (reference
native-helper/src/python.rs:106-116 -> native-helper/src/lib.rs:235-275) on CUDA tensors: the float64 FFT
resampling the reference applies to every chunk it reads from a WAV file whose rate differs from the detector's
(match.py:421-423, audio_utils.py:154-171).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

from . import _lib

_WS_LIMIT_BYTES = 2 << 30          # float64 workspace per call; larger batches are cut into sub-batches
_workspace: dict[int, Any] = {}    # device index -> uint8 CUDA tensor, grown on demand


def workspace_bytes(n_in: int, n_out: int, batch: int = 1) -> int:
    need = C.c_int64(0)
    rc = _lib.lib().apd_resample_workspace_bytes(n_in, n_out, batch, C.byref(need))
    if rc != _lib.APD_OK:
        raise ValueError(f"resample: unsupported lengths {n_in} -> {n_out} (batch {batch})")
    return int(need.value)


def resample_into(src: Any, n_in: int, dst: Any, n_out: int, batch: int = 1, stream: Any = None) -> None:
    """``batch`` signals of ``n_in`` float32 samples stored back to back in the CUDA tensor ``src`` -> ``n_out``
    samples each, back to back in ``dst``.  Enqueued on ``stream`` (default: the current stream)."""
    import torch
    if not (src.is_cuda and dst.is_cuda and src.dtype == torch.float32 and dst.dtype == torch.float32):
        raise TypeError("resample_into needs float32 CUDA tensors")
    if src.numel() < batch * n_in or dst.numel() < batch * n_out:
        raise ValueError("resample_into: tensor shorter than batch * length")
    dev = src.device.index
    stream = stream if stream is not None else torch.cuda.current_stream(dev)
    per = max(1, workspace_bytes(n_in, n_out, 1))
    sub = max(1, min(batch, _WS_LIMIT_BYTES // per, 65535))
    need = workspace_bytes(n_in, n_out, sub)
    ws = _workspace.get(dev)
    if ws is None or ws.numel() < need:
        _workspace.pop(dev, None)                  # release the smaller buffer before allocating the larger one
        ws = None
        ws = _workspace[dev] = torch.empty(max(need, 16), dtype=torch.uint8, device=src.device)
    L = _lib.lib()
    for b0 in range(0, batch, sub):
        k = min(sub, batch - b0)
        rc = L.apd_resample(C.c_void_p(src.data_ptr() + 4 * b0 * n_in), n_in, n_in,
                            C.c_void_p(dst.data_ptr() + 4 * b0 * n_out), n_out, n_out, k,
                            C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(stream.cuda_stream))
        if rc != _lib.APD_OK:
            raise _lib.ApdError(f"apd_resample({n_in} -> {n_out}, batch {k}): {_lib.ERR_NAMES.get(rc, rc)}")


def resample(data: Any, num_samples: int) -> Any:
    """``_native.resample`` on the device: float32 CUDA tensor ``[n]`` or ``[batch, n]`` -> ``[..., num_samples]``."""
    import torch
    x = data.to(dtype=torch.float32).contiguous()
    if x.dim() not in (1, 2):
        raise TypeError("data must be a 1-D signal or a 2-D batch of signals")
    batch = 1 if x.dim() == 1 else x.shape[0]
    n_in = x.shape[-1]
    out = torch.empty((*x.shape[:-1], int(num_samples)), dtype=torch.float32, device=x.device)
    if out.numel():
        with torch.cuda.device(x.device):
            resample_into(x, n_in, out, int(num_samples), batch)
    return out
