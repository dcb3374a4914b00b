import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
from oxli_b200 import _capi as capi
n, L, k, G = 10_000_000, 150, 31, 5_000_000
tb = n * L
d_bases = capi.device_alloc(tb + 64); d_offs = capi.device_alloc((n + 1) * 8)
capi.synth_reads_device(d_bases, n, L, G, 0xC20001)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
t = capi.Table(k, capacity_hint=G)
# populate the table with real keys first (dbg flags only affect the consume kernel; build via hash list)
nw = tb - k + 1
d_h = capi.device_alloc(nw * 8)
t.hash_batch_device(d_bases, d_offs, n, tb, d_h)
t.count_hashes_device(d_h, nw, True)
capi.device_free(d_h)
for it in range(3):
    t.timer_start(); st, tot, _, _ = t.consume_batch_device(d_bases, d_offs, n, tb, True); ms = t.timer_stop()
print("DBG=%s  %.2f ms  (%.1f G k-mers/s) size=%d" % (os.environ.get("OXLI_B200_DBG", "0"), ms, tot / ms / 1e6, len(t)))
