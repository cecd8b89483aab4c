#!/usr/bin/env python
"""bench.py -- mash.Sketch throughput on B200 (BASELINE.json metric).

A "step" is one pass of the Sketch hot path over one batch of synthetic reads:
BASELINE configs[1] = 10 M x 150 bp reads, k=21, sketchSize=1000 per GPU (reads are
independent, so N GPUs sketch N x 10 M reads with no data-path collective: weak scaling;
N=8 is configs[3], 80 M reads).

  value     whole-job Gbases/s with the reads resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the reference-facing C ABI call with HOST buffers
            (pg_mash_sketch_uniform: H2D of the reads, kernels, D2H of the sketches)
  roofline  algorithmic bytes (L + 4*min(L-k, s) per read) / kernel time vs the
            measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  C restatement of the Go algorithm (oracle/, "port") on the host cores

`--impl reference` times that CPU restatement alone (the Go reference cannot run here:
no Go toolchain) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN, KMER, SKETCH = 150, 21, 1000
NK = READ_LEN - KMER  # 129 informative words per read
BYTES_PER_READ = READ_LEN + 4 * min(NK, SKETCH)  # 666
METRIC = "mash.Sketch Gbases/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t_load = self.t0 = self.t1 = None  # load start (warm-up), timed region start / end

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # samples inside the timed region; a region shorter than a few sampling periods falls
        # back to the whole loaded window (warm-up + timed), and says so
        rows = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= self.t1]
        window = "timed region"
        if len(rows) < 3 and self.t_load is not None:
            rows = [r for (t, r) in self.rows if self.t_load <= t <= self.t1 + 0.05]
            window = "warm-up + timed region (timed region shorter than 3 sampling periods)"
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def _cpu_sample(threads: int, budget_s: float, cap_reads: int = 2_000_000):
    """Choose a sample of the cfg2 workload that the C restatement sketches in ~budget_s seconds
    on `threads` host threads, and generate it ONCE (numpy generation is slower than hashing)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi
    from poly_b200 import synth

    probe = min(cap_reads, 10_000 * threads)
    reads = synth.independent_reads(probe, READ_LEN)
    off = synth.uniform_offsets(probe, READ_LEN)
    oracle_ffi.sketch_batch_timing(reads[: 1000 * READ_LEN], off[:1001], KMER, SKETCH, 0, threads)  # thread start-up
    t0 = time.perf_counter()
    oracle_ffi.sketch_batch_timing(reads, off, KMER, SKETCH, 0, threads)
    rate = probe / (time.perf_counter() - t0)
    n = int(min(max(rate * budget_s, probe), cap_reads))
    if n > probe:
        reads = synth.independent_reads(n, READ_LEN)
        off = synth.uniform_offsets(n, READ_LEN)
    return reads, off, n


def _cpu_time(reads, off, n, threads: int):
    """One timed pass of the C restatement of mash.go:59-104 (faithful variant: one zeroed
    4*s-byte sketch per read as mash.New does, full re-sort on qualifying insert; static
    parallel-for over reads).  The only place bench.py executes oracle/."""
    import oracle_ffi

    t0 = time.perf_counter()
    rc, _ = oracle_ffi.sketch_batch_timing(reads, off, KMER, SKETCH, 0, threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    return dt


def _cpu_desc(n, threads, dt, value):
    return {"value": value, "unit": "Gbases/s", "cores": threads, "kind": "port",
            "sample": f"first {n} reads of the 10M x 150bp k=21 s=1000 workload per step; C restatement of the Go algorithm "
                      f"(mash.go:59-104, not Go: no Go toolchain here), static parallel-for over reads on {threads} threads, {dt:.2f} s per step"}


def cpu_reference_leg(threads: int, budget_s: float):
    reads, off, n = _cpu_sample(threads, budget_s)
    dt = min(_cpu_time(reads, off, n, threads) for _ in range(2))
    return _cpu_desc(n, threads, dt, n * READ_LEN / dt / 1e9), n, dt


def run_reference(args, rank: int, world: int):
    """`--impl reference`: the reference's own CPU implementation of the path.  The Go code
    cannot run here, so this is its C restatement on all host threads; each step is one pass
    over a bounded sample of the cfg2 workload (same metric / unit / config as our arm)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = args.steps + args.warmup
    reads, off, n = _cpu_sample(threads, max(0.3, min(4.0, 100.0 / max(steps, 1))))
    times = [_cpu_time(reads, off, n, threads) for _ in range(steps)][args.warmup:]
    tot_t = sum(times)
    value = n * len(times) * READ_LEN / tot_t / 1e9
    sample = _cpu_desc(n, threads, tot_t / len(times), value)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "configs[1]: 10M x 150bp short reads, k=21, sketchSize=1000 (bounded sample per step)",
                   "reads_per_step": n, "read_len": READ_LEN, "k": KMER, "sketch_size": SKETCH},
        "cpu_baseline": sample,
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    from poly_b200 import _lib, synth

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    _lib.check(L.pg_init(local))
    dev = torch.device("cuda", local)
    n = args.reads
    first_read = rank * n  # rank r sketches reads [r*n, (r+1)*n): no data-path collective
    stream = torch.cuda.current_stream().cuda_stream

    d_in = torch.empty(n * READ_LEN, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n * NK, dtype=torch.int32, device=dev)
    _lib.check(L.pg_synth_reads_dev(d_in.data_ptr(), first_read, n, READ_LEN, synth.SEED_READS, 0, 0, stream))
    torch.cuda.synchronize()

    def step():
        _lib.check(L.pg_mash_sketch_uniform_dev(d_in.data_ptr(), n, READ_LEN, KMER, SKETCH, 0, d_out.data_ptr(), NK, None, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)  # let nvidia-smi start streaming
    clocks.t_load = time.time()
    for _ in range(max(args.warmup, 3)):
        step()
    # keep the device loaded for >= 100 ms before the timed region so that the clock samples
    # describe clocks under load (untimed; same kernel)
    barrier()
    t_pre = time.time()
    while time.time() - t_pre < 0.1:
        step()
        torch.cuda.synchronize()
    barrier()
    launches0 = L.pg_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks.t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = L.pg_launch_count() - launches0
    kernel_name = (L.pg_last_kernel() or b"").decode()
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * n * READ_LEN / (ms_per_step * 1e-3) / 1e9  # whole-job Gbases/s

    # ---- end to end through the host-pointer C ABI (pinned host buffers) -------------
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n * READ_LEN, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(n * NK, dtype=torch.int32, pin_memory=True)
        h_in.copy_(d_in)
        torch.cuda.synchronize()
        ksteps = args.e2e_steps or min(args.steps, 5)

        def e2e_step():
            _lib.check(L.pg_mash_sketch_uniform(h_in.data_ptr(), n, READ_LEN, KMER, SKETCH, 0, h_out.data_ptr(), NK, None))

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()  # synchronous: returns when the sketches are in host memory
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * READ_LEN * ksteps / dt / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": n * READ_LEN,
               "d2h_bytes_per_step": n * NK * 4, "steps": ksteps, "ms_per_step": 1e3 * dt / ksteps,
               "api": "pg_mash_sketch_uniform (host buffers, pinned)"}
        # the host path and the device path must agree bit for bit
        same = bool(torch.equal(h_out[: 4096 * NK].to(dev), d_out[: 4096 * NK]))
        e2e["matches_device_path"] = same

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    achieved = n * BYTES_PER_READ / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_read"] * n  # per launch, scaled from the ncu capture's reads
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "algorithmic_bytes_per_launch": n * BYTES_PER_READ,
                "peak_source": peak_src, "avg_launch_ms": ms_per_step}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cpu_baseline, ns, _ = cpu_reference_leg(os.cpu_count() or 1, 8.0)
        # parity spot check of the GPU result against the oracle on the head of the batch
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_ffi
        m = 2048
        rc, want = oracle_ffi.sketch_batch(synth.independent_reads(m, READ_LEN), synth.uniform_offsets(m, READ_LEN), KMER, SKETCH, 1, 1)
        got = d_out[: m * NK].cpu().numpy().view(np.uint32).reshape(m, NK)
        cpu_baseline["gpu_matches_oracle_on_first_2048_reads"] = bool(np.array_equal(got, want[:, :NK]))

    out = {
        "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 10M x 150bp short reads, k=21, sketchSize=1000 per GPU"
                   + (" (x8 = configs[3], 80M reads)" if world == 8 else ""),
                   "reads_per_gpu": n, "read_len": READ_LEN, "k": KMER, "sketch_size": SKETCH,
                   "l2": "inputs (1.5 GB) + outputs (5.2 GB) per step are larger than L2 (126 MB); no explicit flush",
                   "parallelism": f"reads sharded over {world} GPU(s), no data-path collective"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
