"""Sharding the chunk loop across the GPUs of one box.

A (chunk x pattern) unit needs only samples ``[i*C - sw*sr, (i+1)*C)`` and the pattern's
precomputed data (reference audio_pattern_detector.py:406-412; ``previous_chunk`` at :330 is just
that look-back), so the path shards with no data-path collective: every rank owns a contiguous
range of chunks, holds the samples of that range plus a look-back halo, replicates the pattern
spectra, and scans independently.  Only the accepted detections (a few KB) travel: they are
gathered to rank 0 *through the host* (a gloo group; NCCL is not used on this path) and
concatenated in rank order, which is chunk order, so callbacks keep the reference's order
(chunk by chunk, sorted by timestamp inside a chunk, :324-327).

When a stream has fewer chunks than there are GPUs worth feeding (a short file against hundreds of patterns), the
other axis of the unit grid is split instead: every rank scans ALL chunks against a contiguous slice of the pattern
list (``sharded_scan_by_pattern``).  The merge then has to interleave the ranks' detections chunk by chunk, so each
detection travels with its chunk index and its clip's position in the full list.
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


Detection = tuple[int, float, int, str]          # (chunk index, timestamp, clip index in the full list, clip name)


def pattern_range_for_rank(n_patterns: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of the pattern list (same rule as ``chunk_range_for_rank``)."""
    return chunk_range_for_rank(n_patterns, world, rank)


def merge_pattern_shards(shards: list[list[Detection]], clip_names: list[str]) -> ShardResult:
    """Detections of disjoint pattern slices -> the single-process result: per chunk the detections sorted by
    timestamp, equal timestamps in clip-list order (the stable sort of reference :324-327 over the clip loop of
    :306-313); ``peak_times`` has every clip name, as reference :256 initialises it."""
    merged = sorted((d for shard in shards for d in shard), key=lambda d: (d[0], d[1], d[2]))
    peak_times: dict[str, list[float]] = {name: [] for name in clip_names}
    per_clip = sorted(merged, key=lambda d: (d[2], d[0]))           # stable: a clip's own order inside a chunk is kept
    for _, t, _, name in per_clip:
        peak_times[name].append(t)
    return peak_times, [(t, name) for _, t, _, name in merged]


def detections_of(result: Any, clip_offset: int = 0) -> list[Detection]:
    """Accepted candidates of an ``AudioPatternDetector.scan_array`` result as ``Detection`` tuples;
    ``clip_offset`` is the position of the detector's first clip in the full pattern list."""
    out: list[Detection] = []
    for rec, t in zip(result.records, result.timestamps):
        if int(rec["flags"]) & 1:
            out.append((int(rec["chunk"]), float(t), clip_offset + int(rec["clip"]), result.clip_names[int(rec["clip"])]))
    return out


def sharded_scan_by_pattern(scan_patterns: Callable[[int, int], list[Detection]], clip_names: list[str],
                            group: Any = None, dst: int = 0) -> Optional[ShardResult]:
    """Run ``scan_patterns(pattern_begin, pattern_end)`` - all chunks against that slice of the pattern list - on this
    rank's share of the patterns and gather + interleave the detections on rank ``dst`` (None on the others)."""
    import torch.distributed as dist
    n = len(clip_names)
    if not (dist.is_available() and dist.is_initialized()):
        return merge_pattern_shards([scan_patterns(0, n)], clip_names)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    p0, p1 = pattern_range_for_rank(n, world, rank)
    local = scan_patterns(p0, p1) if p1 > p0 else []
    gathered: Optional[list[Any]] = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    assert gathered is not None
    return merge_pattern_shards(gathered, clip_names)


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
