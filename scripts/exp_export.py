#!/usr/bin/env python
"""Export of a large table: slot order, sorted by hash, sorted by (count, hash) -- the device
radix sort against numpy on the host (equality checked), seconds and GB/s of (hash, count) pairs."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oxli_b200 import _capi as capi  # noqa: E402

n, L, k = int(os.environ.get("READS", 3_000_000)), 150, 21
d_b = capi.device_alloc(n * L + 64); d_o = capi.device_alloc((n + 1) * 8)
capi.synth_reads_device(d_b, n, L, 100_000_000, 0xC30001, sub_ppm=10_000, n_ppm=1_000)
capi.h2d(d_o, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
t = capi.Table(k)
t.consume_batch_device(d_b, d_o, n, n * L, True)
m = len(t)
print(f"table: {m / 1e6:.1f} M keys, {t.capacity * 16 / 2**30:.1f} GiB of slots")
res = {}
for mode, name in ((0, "slot order"), (1, "sorted by hash"), (2, "sorted by (count, hash)")):
    t.export(mode) if mode == 0 else None  # warm the scratch allocations once
    t0 = time.perf_counter()
    kk, vv = t.export(mode)
    dt = time.perf_counter() - t0
    res[mode] = (kk, vv)
    print(f"export {name:24s} {dt * 1e3:9.1f} ms   {m * 16 / dt / 1e9:6.2f} GB/s of pairs to host memory", flush=True)
k0, v0 = res[0]
t0 = time.perf_counter(); order = np.argsort(k0, kind="stable"); dt1 = time.perf_counter() - t0
assert np.array_equal(res[1][0], k0[order]) and np.array_equal(res[1][1], v0[order])
t0 = time.perf_counter(); order = np.lexsort((k0, v0)); dt2 = time.perf_counter() - t0
assert np.array_equal(res[2][0], k0[order]) and np.array_equal(res[2][1], v0[order])
print(f"equal to the host-sorted result; numpy argsort {dt1 * 1e3:.0f} ms, lexsort {dt2 * 1e3:.0f} ms on this box")
