#!/usr/bin/env python
"""Low-complexity input at scale: every read is a homopolymer / short tandem repeat, so all windows
of the batch are one of a handful of k-mers.  Fused kernel vs partitioned pipeline."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oxli_b200 import _capi as capi  # noqa: E402

n, L, k = int(os.environ.get("READS", 2_000_000)), 150, 31
for name, unit in (("poly-A", b"A"), ("(AT)n", b"AT"), ("(ACG)n", b"ACG"), ("mixed: 1% poly-A in random", None)):
    if unit is not None:
        read = (unit * L)[:L]
        bases = np.frombuffer(read * n, dtype=np.uint8)
    else:
        rng = np.random.default_rng(1)
        bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * L)].copy()
        rows = bases.reshape(n, L)
        rows[rng.random(n) < 0.01] = ord("A")
    offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    d_b = capi.device_alloc(n * L + 64); d_o = capi.device_alloc((n + 1) * 8)
    capi.h2d(d_b, bases); capi.h2d(d_o, offs)
    res = {}
    for pipe in ("fused", "part"):
        capi.set_pipeline(pipe)
        t = capi.Table(k)
        best = 1e30
        for _ in range(3):
            t.clear()
            t.timer_start()
            st, total, _, _ = t.consume_batch_device(d_b, d_o, n, n * L, True)
            best = min(best, t.timer_stop())
        res[pipe] = (best, total, len(t), t.stats()["max"])
        t.close()
    capi.set_pipeline("auto")
    capi.device_free(d_b); capi.device_free(d_o)
    f, p = res["fused"], res["part"]
    assert f[1:] == p[1:], (f, p)
    print(f"{name:28s} {p[1] / 1e6:7.1f} M k-mers, {p[2]} distinct: fused {f[0]:8.2f} ms ({f[1] / f[0] / 1e6:6.2f} G/s)   "
          f"partitioned {p[0]:8.2f} ms ({p[1] / p[0] / 1e6:6.2f} G/s)", flush=True)
