#!/usr/bin/env python
"""Partitioned pipeline vs fused kernel on a device-resident workload (default: C2 shape).
Prints one line per configuration: ms per step, G k-mers/s, pass A / pass B split, and checks
that every configuration builds the same table (digest)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oxli_b200 import _capi as capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--read-len", type=int, default=150)
ap.add_argument("--genome", type=int, default=5_000_000)
ap.add_argument("--ksize", type=int, default=31)
ap.add_argument("--sub-ppm", type=int, default=0)
ap.add_argument("--n-ppm", type=int, default=0)
ap.add_argument("--hint", type=int, default=-1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--configs", default="fused,part,part:512,part:1024,part:2048,part:4096")
ap.add_argument("--digest", action="store_true")
ap.add_argument("--fresh", action="store_true", help="a new table every step (growth from nothing is part of the step)")
a = ap.parse_args()

n, L, k = a.reads, a.read_len, a.ksize
d_bases = capi.device_alloc(n * L + 64)
d_offs = capi.device_alloc((n + 1) * 8)
capi.synth_reads_device(d_bases, n, L, a.genome, 0xC20001, sub_ppm=a.sub_ppm, n_ppm=a.n_ppm)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
hint = a.genome if a.hint < 0 else a.hint
ref = None
for cfg in a.configs.split(","):
    name, *rest = cfg.split(":")
    parts = int(rest[0]) if rest else 0
    groups = int(rest[1]) if len(rest) > 1 else 0
    capi.set_pipeline(name, parts, groups)
    t = capi.Table(k, capacity_hint=hint)
    best, split = 1e30, (0, 0)
    for s in range(a.steps + 1):
        if a.fresh:
            t.close()
            t = capi.Table(k, capacity_hint=hint)
        else:
            t.clear()
        t.timer_start()
        st, total, _, _ = t.consume_batch_device(d_bases, d_offs, n, n * L, True)
        ms = t.timer_stop()
        if s and ms < best:
            best, split = ms, t.last_consume_pass_ms()
    kms, nl = t.last_consume_kernel_ms()
    line = (f"{cfg:14s} {best:8.2f} ms/step  {total / best / 1e6:7.2f} G k-mers/s  passA {split[0]:7.2f}  passB {split[1]:7.2f}"
            f"  kernels {kms:7.2f} ms in {nl} launches  len {len(t)}  cap {t.capacity}")
    if a.digest:
        d = (total, tuple(sorted(t.digest().items())))
        ref = ref or d
        line += "  digest " + ("same" if d == ref else "DIFFERENT")
    print(line, flush=True)
    t.close()
capi.set_pipeline("auto")
