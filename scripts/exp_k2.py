import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oxli_b200 import _capi as capi
n, L, k, G = 4_000_000, 150, 31, 5_000_000
tb = n * L
d_bases = capi.device_alloc(tb + 64); d_offs = capi.device_alloc((n + 1) * 8)
capi.synth_reads_device(d_bases, n, L, G, 0xC20001)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
nw = tb - k + 1
d_h = capi.device_alloc(nw * 8)
for hint in (5_000_000, 10_000_000, 20_000_000):
    t = capi.Table(k, capacity_hint=hint)
    t.hash_batch_device(d_bases, d_offs, n, tb, d_h)
    for it in range(3):
        t.timer_start(); c = t.count_hashes_device(d_h, nw, True); ms = t.timer_stop()
        print(f"hint={hint} slots={t.capacity} load={len(t)/t.capacity:.2f} it={it}: {ms:.2f} ms {c/ms/1e6:.1f} G/s kernel={t.last_consume_kernel_ms()}", flush=True)
    t.close()
