"""Host-side sharding logic of the N > 1 path on CPU: world_size-2 gloo processes, each scanning its
chunk range with the oracle as the stand-in scanner, gathered and merged on rank 0."""
import os
import socket

import numpy as np
import pytest

from audio_pattern_detector_b200 import sharding
from audio_pattern_detector_b200 import workloads as W


def test_chunk_ranges_partition_exactly():
    for n in (0, 1, 5, 7, 1440, 60000):
        for world in (1, 2, 3, 4, 8):
            rs = [sharding.chunk_range_for_rank(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.chunk_range_for_rank(4, 2, 2)


def test_slab_bounds_include_halo():
    assert sharding.slab_bounds(0, 3, 100, 30, 1000) == (0, 300)
    assert sharding.slab_bounds(3, 6, 100, 30, 1000) == (270, 600)
    assert sharding.slab_bounds(9, 10, 100, 30, 950) == (870, 950)


def test_merge_keeps_rank_order():
    a = ({"x": [1.0], "y": []}, [(1.0, "x")])
    b = ({"x": [7.0], "y": [6.5]}, [(6.5, "y"), (7.0, "x")])
    times, ev = sharding.merge_shards([a, b])
    assert times == {"x": [1.0, 7.0], "y": [6.5]} and ev == [(1.0, "x"), (6.5, "y"), (7.0, "x")]


def _inputs():
    sr, spc = 8000, 4
    pats = W.make_patterns(3, sr, seed=3, min_s=0.2, max_s=1.5, tone_every=3, chirp_every=2)
    audio, _ = W.make_stream(22.5, pats, sr, seed=5, plants_per_pattern=2, chunk_seconds=spc)
    return sr, spc, pats, audio


def _oracle_scanner(pats, audio, sr, spc):
    from oracle.detector import OracleDetector
    det = OracleDetector(pats, sr, spc, precision="f32")

    def scan(c0, c1):
        times, events, _ = det.run(audio, chunk_range=(c0, c1))
        return times, [(t, n) for t, n, _, _ in events]
    return scan, (audio.size + det.chunk_samples - 1) // det.chunk_samples


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sr, spc, pats, audio = _inputs()
        scan, n_chunks = _oracle_scanner(pats, audio, sr, spc)
        out = sharding.sharded_scan(scan, n_chunks)
        if rank == 0:
            q.put(out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_scan_equals_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sr, spc, pats, audio = _inputs()
    scan, n_chunks = _oracle_scanner(pats, audio, sr, spc)
    want = scan(0, n_chunks)
    assert n_chunks >= 4 and sum(len(v) for v in want[0].values()) > 0
    assert got[0] == want[0] and got[1] == want[1]


# ---- the other axis: every rank scans all chunks against a slice of the pattern list ----
def _oracle_pattern_scanner(pats, audio, sr, spc):
    from oracle.detector import OracleDetector

    def scan(p0, p1):
        det = OracleDetector(pats[p0:p1], sr, spc, precision="f32")
        _, events, _ = det.run(audio)
        index = {p["name"]: p0 + k for k, p in enumerate(pats[p0:p1])}
        return [(chunk, t, index[name], name) for t, name, chunk, _ in events]
    return scan


def test_pattern_ranges_and_merge_order():
    assert [sharding.pattern_range_for_rank(5, 3, r) for r in range(3)] == [(0, 2), (2, 4), (4, 5)]
    names = ["a", "b", "c"]
    shard0 = [(0, 2.0, 0, "a"), (1, 5.0, 0, "a"), (1, 5.5, 1, "b")]
    shard1 = [(0, 1.0, 2, "c"), (1, 5.0, 2, "c"), (1, 4.0, 2, "c")]
    times, events = sharding.merge_pattern_shards([shard0, shard1], names)
    # chunk by chunk, by timestamp, equal timestamps in clip-list order
    assert events == [(1.0, "c"), (2.0, "a"), (4.0, "c"), (5.0, "a"), (5.0, "c"), (5.5, "b")]
    assert times == {"a": [2.0, 5.0], "b": [5.5], "c": [1.0, 4.0, 5.0]}
    assert sharding.merge_pattern_shards([], names) == ({"a": [], "b": [], "c": []}, [])


def _pattern_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sr, spc, pats, audio = _inputs()
        out = sharding.sharded_scan_by_pattern(_oracle_pattern_scanner(pats, audio, sr, spc), [p["name"] for p in pats])
        if rank == 0:
            q.put(out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_pattern_split_equals_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pattern_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sr, spc, pats, audio = _inputs()
    scan, n_chunks = _oracle_scanner(pats, audio, sr, spc)
    want = scan(0, n_chunks)
    assert sum(len(v) for v in want[0].values()) > 0
    assert got[0] == want[0] and got[1] == want[1]
