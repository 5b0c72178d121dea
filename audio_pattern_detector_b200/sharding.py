"""Sharding the chunk loop across the GPUs of one box.

A (chunk x pattern) unit needs only samples ``[i*C - sw*sr, (i+1)*C)`` and the pattern's
precomputed data (reference audio_pattern_detector.py:406-412; ``previous_chunk`` at :330 is just
that look-back), so the path shards with no data-path collective: every rank owns a contiguous
range of chunks, holds the samples of that range plus a look-back halo, replicates the pattern
spectra, and scans independently.  Only the accepted detections (a few KB) travel: they are
gathered to rank 0 *through the host* (a gloo group; NCCL is not used on this path) and
concatenated in rank order, which is chunk order, so callbacks keep the reference's order
(chunk by chunk, sorted by timestamp inside a chunk, :324-327).
"""
from __future__ import annotations

from collections.abc import Callable
from typing import Any, Optional

ShardResult = tuple[dict[str, list[float]], list[tuple[float, str]]]      # (peak_times, events)


def chunk_range_for_rank(n_chunks: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first ``n_chunks % world`` ranks get one extra chunk."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    q, r = divmod(max(n_chunks, 0), world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def slab_bounds(chunk_begin: int, chunk_end: int, chunk_samples: int, max_halo: int,
                total_samples: int) -> tuple[int, int]:
    """Stream samples ``[first, last)`` a rank must hold to scan chunks [chunk_begin, chunk_end)."""
    first = max(0, chunk_begin * chunk_samples - (max_halo if chunk_begin > 0 else 0))
    last = min(total_samples, chunk_end * chunk_samples)
    return first, max(first, last)


def merge_shards(shards: list[ShardResult]) -> ShardResult:
    """Concatenate per-rank results given in rank (= chunk) order."""
    peak_times: dict[str, list[float]] = {}
    events: list[tuple[float, str]] = []
    for times, ev in shards:
        for name, ts in times.items():
            peak_times.setdefault(name, []).extend(ts)
        events.extend(ev)
    return peak_times, events


def sharded_scan(scan_range: Callable[[int, int], ShardResult], n_chunks: int, group: Any = None,
                 dst: int = 0) -> Optional[ShardResult]:
    """Run ``scan_range(chunk_begin, chunk_end)`` on this rank's share of ``n_chunks`` and gather the
    detections on rank ``dst`` (returns None on the other ranks).

    ``group`` must be a host (gloo) process group, or None for the default group when that is gloo;
    without an initialised process group this is a plain single-rank scan."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return scan_range(0, n_chunks)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    c0, c1 = chunk_range_for_rank(n_chunks, world, rank)
    local = scan_range(c0, c1) if c1 > c0 else ({}, [])
    gathered: Optional[list[Any]] = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    assert gathered is not None
    return merge_shards(gathered)
