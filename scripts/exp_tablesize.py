"""K2 (count a device hash list) throughput vs table footprint: where does the
random-access update rate stop being DRAM-bound?"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oxli_b200 import _capi as capi
n, L, k = 4_000_000, 150, 31
tb = n * L
d_bases = capi.device_alloc(tb + 64); d_offs = capi.device_alloc((n + 1) * 8)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
nw = tb - k + 1
d_h = capi.device_alloc(nw * 8)
for G in (100_000, 300_000, 600_000, 1_200_000, 2_500_000, 5_000_000, 10_000_000, 40_000_000):
    capi.synth_reads_device(d_bases, n, L, G, 0xC20001)
    t = capi.Table(k, capacity_hint=G)
    t.hash_batch_device(d_bases, d_offs, n, tb, d_h)
    best = 1e9
    for it in range(3):
        t.clear()
        t.timer_start(); c = t.count_hashes_device(d_h, nw, True); ms = t.timer_stop()
        best = min(best, ms)
    # steady state: all keys present
    t.timer_start(); c = t.count_hashes_device(d_h, nw, True); ms2 = t.timer_stop()
    print(f"G={G:>9} distinct={len(t):>9} slots={t.capacity:>10} table={t.capacity*16/2**20:7.1f} MiB  K2 {c/best/1e6:6.1f} G/s (fresh) {c/ms2/1e6:6.1f} G/s (steady)", flush=True)
    t.close()
