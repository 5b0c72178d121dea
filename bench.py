#!/usr/bin/env python
"""Headline benchmark: audio-hours scanned per second (8 kHz, 64 patterns, 60 s chunks) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # B200 path
    python bench.py --impl reference --steps K --warmup W    # CPU arm (oracle port, all host cores)

A "step" is one pass of the detection hot path over a whole synthetic stream
(BASELINE.json configs[2]: 24 h of 8 kHz radio x 64 patterns of 0.3-10 s, 60 s chunks =
92,160 (chunk x pattern) units) per GPU.  Streams are generated per rank on the device
(weak scaling: every GPU scans its own 24 h slab, no collective on the data path; torch.distributed
is used only for the barrier and the max-over-ranks of the step time).

JSON keys are described in DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-hours scanned/sec (8 kHz, 64 patterns, 60 s chunks)"
UNIT = "audio-hours/s"
SR = 8000
SPC = 60
N_PATTERNS = 64
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel pair from the committed ncu --set full
# capture (profiles/ncu_r1_end_corr_summary.txt, end of round 1; profiles/ncu_r1_corr_summary.txt earlier: 3.238 GB),
# scaled from the captured launches to one step: one 512-unit launch pair of the 640 x 512 shape moves 3.265 GB
# (k_corr_rows 0.443 r + 1.284 w, k_corr_cols2 1.534 r + 0.004 w) = 6.38 MB/unit; the 512 x 512 shape scales with M
TRAFFIC_BYTES_PER_STEP = int(49 * 1440 * 3.265e9 / 512 + 15 * 1440 * 3.265e9 / 512 * 0.8)


def measured_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index = index
        self.rows: list[list[str]] = []
        self.stop_flag = threading.Event()

    def run(self) -> None:
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(index: int) -> None:
    """One process per GPU: run (and therefore first-touch / pin host memory) on the CPUs of the NUMA node the GPU's
    PCIe root hangs off, so that the H2D copies of the e2e arm do not cross sockets.  Best effort, Linux only."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus: set[int] = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:  # noqa: BLE001 - affinity is an optimisation, never a requirement
        pass


def cpu_oracle_rate(patterns, audio_np: np.ndarray, n_chunks: int, procs: int) -> tuple[float, float, int]:
    """Oracle port (oracle/detector.py) over chunks [1, 1+n_chunks) of audio_np, sharded over `procs`
    processes (each chunk keeps its look-back halo).  Returns (audio-hours/s, seconds, units)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    first = 1
    bounds = np.linspace(first, first + n_chunks, procs + 1).astype(int)
    jobs = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    global _CPU_JOB
    _CPU_JOB = (patterns, audio_np)
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    hours = n_chunks * SPC / 3600.0
    return hours / dt, dt, sum(res)


_CPU_JOB = None


def _cpu_worker(rng):
    from oracle.detector import OracleDetector
    patterns, audio = _CPU_JOB
    det = OracleDetector(patterns, SR, SPC, precision="f32")
    n = [0]
    det.run(audio, on_unit=lambda i, st, tr: n.__setitem__(0, n[0] + 1), chunk_range=rng)
    return n[0]


def run_reference(args) -> None:
    """CPU arm: the oracle port of the reference path (the reference's Rust/wheel natives cannot be
    installed offline), all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from audio_pattern_detector_b200 import workloads as W
    from oracle import native
    native.build()
    cores = os.cpu_count() or 1
    patterns = W.make_patterns(N_PATTERNS, SR, seed=1)
    per_step = max(cores, 4) * args.cpu_chunks_per_core
    audio, _ = W.make_stream((per_step + 2) * SPC, patterns, SR, seed=0, plants_per_pattern=1, chunk_seconds=SPC)
    rates, secs = [], []
    for s in range(args.warmup + args.steps):
        r, dt, units = cpu_oracle_rate(patterns, audio, per_step, cores)
        if s >= args.warmup:
            rates.append(r)
            secs.append(dt)
    value = float(np.mean(rates))
    sample = f"{per_step} chunks x {N_PATTERNS} patterns ({per_step * N_PATTERNS} units) of the same workload per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1000,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "8 kHz synthetic radio x 64 patterns (0.3-10 s), 60 s chunks (configs[2])",
                   "chunk_seconds": SPC, "patterns": N_PATTERNS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hours", type=float, default=24.0, help="audio hours per GPU per step")
    ap.add_argument("--batch-chunks", type=int, default=48)
    ap.add_argument("--cpu-chunks-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly one JSON line: whatever native libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        bind_to_gpu_numa_node(local)      # before any pinned allocation: host buffers on the GPU's own socket
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from audio_pattern_detector_b200 import workloads as W
    from audio_pattern_detector_b200.audio_clip import AudioClip
    from audio_pattern_detector_b200.audio_pattern_detector import AudioPatternDetector

    patterns = W.make_patterns(N_PATTERNS, SR, seed=1)
    seconds = args.hours * 3600.0
    audio, plants = W.make_stream_device(seconds, patterns, SR, seed=rank, plants_per_pattern=max(1, int(args.hours)),
                                         chunk_seconds=SPC, device=dev)
    n = audio.numel()
    n_chunks = (n + SPC * SR - 1) // (SPC * SR)
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.copy_(audio)
    torch.cuda.synchronize()

    clips = [AudioClip(name=p["name"], audio=p["audio"], sample_rate=SR, strategy=p["strategy"],
                       strategy_params=p["strategy_params"]) for p in patterns]
    import logging
    logging.getLogger("audio_pattern_detector_b200").setLevel(logging.ERROR)
    sys.stderr, saved = open(os.devnull, "w"), sys.stderr
    det = AudioPatternDetector(audio_clips=clips, seconds_per_chunk=SPC, target_sample_rate=SR,
                               device=local, max_batch_chunks=args.batch_chunks)
    sys.stderr = saved
    alg_bytes = 0
    for ci in (0, 1):
        for p in range(N_PATTERNS):
            nb = 8 * det.unit_n_out(ci, p, n)
            alg_bytes += nb * (1 if ci == 0 else n_chunks - 2)
    for p in range(N_PATTERNS):
        alg_bytes += 8 * det.unit_n_out(n_chunks - 1, p, n)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, out

    # inputs (2.8 GB/GPU) exceed the 126 MB L2, so every step re-reads them from HBM
    resident = lambda: det.scan_array(audio)            # noqa: E731
    through_host = lambda: det.scan_array(host)         # noqa: E731  (pinned host buffer -> H2D inside the call)

    for _ in range(args.warmup):
        res = resident()
    det.enable_profiling(True)
    det.stage_times_ms(reset=True)
    l0 = det.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, res = timed(resident, args.steps)
    sampler.stop_flag.set()
    sampler.join()
    launches = (det.launch_count() - l0) // args.steps
    stages = {k: v / args.steps for k, v in det.stage_times_ms(reset=True).items()}
    det.enable_profiling(False)
    through_host()                                       # untimed: first-use allocations of the host-input path
    ms_e2e, res_h = timed(through_host, max(1, min(args.steps, 2)))
    assert res_h.peak_times == res.peak_times

    # row N1 (streaming ingestion): the same scan from 16-bit PCM in pinned host memory -- half the PCIe bytes, widened
    # to float32 on the device (apd_pcm_to_float); the stream is the synthetic one quantised to int16
    pcm = (audio * 32768.0).round_().clamp_(-32768, 32767).to(torch.int16).cpu().pin_memory()
    through_pcm = lambda: det.scan_array(pcm)           # noqa: E731
    through_pcm()
    ms_pcm, res_p = timed(through_pcm, max(1, min(args.steps, 2)))
    del pcm

    # the dominant stage timed alone (nothing else on the GPU), for the roofline of the kernel itself; the
    # in-step figure above shares the SMs with the overlapped phase-2 and loudness streams
    iso_ms, iso_launches = det.time_correlate_stage(audio)

    hours_total = args.hours * world
    value = hours_total / (ms_step / 1000.0)
    e2e_value = hours_total / (ms_e2e / 1000.0)
    n_det = sum(len(v) for v in res.peak_times.values())
    peak, peak_src = measured_peak_gbs()
    achieved_in_step = alg_bytes / 1e9 / (stages["correlate_max"] / 1000.0)
    achieved = alg_bytes / 1e9 / (iso_ms / 1000.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.hours:g} h synthetic 8 kHz radio per GPU x 64 patterns (0.3-10 s), "
                                   "60 s chunks (BASELINE configs[2])",
                       "units_per_step_per_gpu": int(n_chunks * N_PATTERNS), "batch_chunks": args.batch_chunks,
                       "l2": "inputs (2.8 GB/GPU) larger than L2; no flush needed",
                       "detections_per_step": n_det, "planted": len(plants)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 4),
                    "d2h_bytes_per_step": int(res.n_candidates * 176 + n_chunks // args.batch_chunks * 64),
                    "ms_per_step": ms_e2e},
            "e2e_pcm16": {"value": hours_total / (ms_pcm / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(n * 2),
                          "ms_per_step": ms_pcm, "detections_per_step": sum(len(v) for v in res_p.peak_times.values()),
                          "note": "same scan from int16 PCM in pinned host memory, widened on the device (row N1)"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "stage_ms": stages,
            "roofline": {"bound": "hbm", "kernel": "fused spectral multiply + inverse FFT + |.| + max "
                                                   "(k_corr_rows + k_corr_cols2)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_BYTES_PER_STEP, "algorithmic_bytes_per_step": int(alg_bytes),
                         "peak_source": peak_src,
                         "timing": "stage timed alone over the whole workload (CUDA events, apd_stage_correlate_max "
                                   "per sub-batch after untimed loudness + forward stages)",
                         "stage_ms_alone": iso_ms, "launches_alone": int(iso_launches),
                         "in_step": {"achieved": achieved_in_step, "frac": achieved_in_step / peak,
                                     "stage_ms": stages["correlate_max"],
                                     "note": "same stage inside the timed step, sharing the SMs with the "
                                             "overlapped phase-2 and loudness streams"}},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            nck = max(cores, 4) * args.cpu_chunks_per_core
            sample_audio = host[: (nck + 2) * SPC * SR].numpy().copy()
            rate, dt, units = cpu_oracle_rate(patterns, sample_audio, nck, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"chunks 1..{nck} of the same stream x 64 patterns ({units} units), "
                                              f"{dt:.1f} s wall; oracle port of the reference's scipy-style CPU path"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
