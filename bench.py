#!/usr/bin/env python
"""bench.py -- canonical k-mers counted per second on the C2 workload
(BASELINE.json configs[1]: k=31, 10 M synthetic 150-bp reads from a random
5 Mbp genome, ~300x coverage; one B200, small hot table).

A step = build the count table for the whole read set from scratch
(clear the table, consume every read).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference ...                      CPU restatement of the reference path

Prints ONE JSON line (rank 0).  `value` = device-timed, reads resident in HBM;
`e2e` = same job through the host-buffer C-ABI call (pinned host -> H2D inside
the timed region); `roofline` = the consume kernel against measured HBM copy
bandwidth; `cpu_baseline` = the oracle port on a bounded sample of the same
reads on this box's host cores.
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

METRIC = "canonical k-mers counted/sec"
UNIT = "kmers/s"
SEED = 0xC20001


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--ksize", type=int, default=31)
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--table-hint", type=int, default=0, help="expected distinct k-mers (default: genome length)")
    return ap.parse_args()


def workload_name(a, world):
    return (f"C2: k={a.ksize}, {a.reads * world / 1e6:g} M x {a.read_len}-bp reads "
            f"({a.reads * world * a.read_len / 1e9:g} Gbp) from a random {a.genome / 1e6:g} Mbp genome")


def measured_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic_per_launch(kmers_per_launch: float):
    """dram__bytes_read+write of the consume kernel from the committed ncu capture,
    scaled to this run's k-mers per launch (same workload shape); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)
        return t["dram_bytes_per_launch"] * kmers_per_launch / t["kmers_per_launch"]
    except Exception:
        return None


def alg_bytes_per_kmer(read_len: int, k: int) -> float:
    # every base read once (1 B) + 16-B slot read + 8-B count write
    return read_len / (read_len - k + 1) + 24.0


def run_reference(a, rank, world):
    """CPU arm: the oracle port (C restatement of src/lib.rs:545-607 + sourmash/murmur3)
    with all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    import oracle
    from oracle.synth import synth_reads, uniform_offsets

    cores = os.cpu_count() or 1
    n = min(a.reads, max(a.cpu_sample_reads, 50_000 * cores))
    bases = synth_reads(n, a.read_len, a.genome, SEED)
    offs = uniform_offsets(n, a.read_len)
    times, total = [], 0
    for i in range(a.warmup + a.steps):
        t = oracle.OracleTable(a.ksize)
        t0 = time.perf_counter()
        total, _, _ = t.consume_batch(bases, offs, True, nthreads=cores)
        dt = time.perf_counter() - t0
        if i >= a.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = total / (ms / 1e3)
    sample = f"first {n} reads of the workload ({total} k-mers) per step, table rebuilt each step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a, 1), "ksize": a.ksize, "read_len": a.read_len, "genome_len": a.genome},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "the Rust reference cannot be built here (no cargo/rustc); this is oracle/, "
                                 "its C restatement, reads sharded over threads and merged like KmerCountTable.add"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(a, capi, d_bases) -> dict:
    import oracle
    from oracle.synth import uniform_offsets

    n = min(a.reads, a.cpu_sample_reads)
    bases = np.empty(n * a.read_len, dtype=np.uint8)
    capi.d2h(bases, d_bases)
    offs = uniform_offsets(n, a.read_len)
    t = oracle.OracleTable(a.ksize)
    t0 = time.perf_counter()
    total, _, _ = t.consume_batch(bases, offs, True, nthreads=1)
    dt = time.perf_counter() - t0
    return {"value": total / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n} reads of the same workload ({total} k-mers), single thread like the reference's consume",
            "seconds": dt}


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference(a, rank, world)

    from oxli_b200 import _capi as capi

    if world > 1:
        from oxli_b200.sharded import run_sharded_bench
        return run_sharded_bench(a, rank, world, local)

    n, L, k = a.reads, a.read_len, a.ksize
    total_bases = n * L
    kmers_per_step = n * (L - k + 1)
    dev = local
    d_bases = capi.device_alloc(total_bases + 64, dev)
    d_offs = capi.device_alloc((n + 1) * 8, dev)
    capi.synth_reads_device(d_bases, n, L, a.genome, SEED, device=dev)
    offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    capi.h2d(d_offs, offs, dev)
    table = capi.Table(k, device=dev, capacity_hint=a.table_hint or a.genome)

    def step_resident():
        table.clear()
        st, total, _, _ = table.consume_batch_device(d_bases, d_offs, n, total_bases, True)
        assert st == 0 and total == kmers_per_step, (st, total, kmers_per_step)
        return table.last_consume_kernel_ms()

    for _ in range(a.warmup):
        step_resident()
    table.sync()
    launches0 = int(capi.lib.oxg_launch_count())
    kernel_ms, kernel_launches = 0.0, 0
    with ClockSampler(dev) as clocks:
        table.timer_start()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ms, nl = step_resident()
            kernel_ms += ms; kernel_launches += nl
        dev_ms = table.timer_stop()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = int(capi.lib.oxg_launch_count()) - launches0
    ms_per_step = dev_ms / a.steps
    value = kmers_per_step / (ms_per_step / 1e3)
    distinct = len(table)
    cap = table.capacity

    # roofline of the dominant kernel (consume_kernel<31, count>): algorithmic bytes / its own duration
    peak, peak_src = measured_peak_gbs()
    balg = alg_bytes_per_kmer(L, k)
    avg_launch_ms = kernel_ms / max(kernel_launches, 1)
    kmers_per_launch = kmers_per_step * a.steps / max(kernel_launches, 1)
    achieved = balg * kmers_per_launch / (avg_launch_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": measured_traffic_per_launch(kmers_per_launch) if (k, L) == (31, 150) else None,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_traffic.json)",
                "alg_bytes_per_launch": balg * kmers_per_launch,
                "kernel": f"consume_kernel<{k},count>", "alg_bytes_per_kmer": balg,
                "kmers_per_launch": kmers_per_launch, "avg_launch_ms": avg_launch_ms,
                "kernel_share_of_step": kernel_ms / dev_ms, "peak_source": peak_src}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a, 1), "ksize": k, "read_len": L, "reads": n, "genome_len": a.genome,
                   "distinct_kmers": distinct, "table_slots": cap, "table_bytes": cap * 16,
                   "l2_policy": "inputs (1.5 GB of reads per step) far exceed the 126 MB L2; no flush needed",
                   "gbases_per_s": total_bases / (ms_per_step / 1e3) / 1e9},
        "roofline": roofline, "gpu_launches": launches, "wall_ms_per_step": wall_ms / a.steps,
        "clocks": clocks.summary(),
    }

    if not a.no_e2e:
        # same job through the host-buffer entry point: pinned host memory -> H2D inside the timed region
        h_bases = capi.pinned_empty(total_bases)
        capi.d2h(h_bases, d_bases, dev)
        h_offs_raw = capi.pinned_empty((n + 1) * 8)
        h_offs = h_offs_raw.view(np.uint64)
        h_offs[:] = offs

        def step_e2e():
            table.clear()
            st, total, _, _ = table.consume_batch(h_bases, h_offs, True)
            assert st == 0 and total == kmers_per_step
            return len(table)  # device->host read of the result

        for _ in range(max(1, a.warmup // 2)):
            step_e2e()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            got = step_e2e()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / a.steps
        assert got == distinct
        n_chunks = -(-total_bases // (64 << 20))
        out["e2e"] = {"value": kmers_per_step / (e2e_ms / 1e3), "unit": UNIT,
                      "h2d_bytes_per_step": int(total_bases + (n + 1) * 8 + n_chunks * 8),
                      "d2h_bytes_per_step": int((n_chunks + 2) * 128),
                      "ms_per_step": e2e_ms, "timing": "host wall clock around the blocking C-ABI call"}
        capi.pinned_free(h_bases); capi.pinned_free(h_offs_raw)

    if not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a, capi, d_bases)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
