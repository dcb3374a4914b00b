import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from oxli_b200 import _capi as capi
n, L, k, G = 20_000_000, 150, 21, 100_000_000
tb = n * L
d_bases = capi.device_alloc(tb + 64); d_offs = capi.device_alloc((n + 1) * 8)
capi.synth_reads_device(d_bases, n, L, G, 0xC30001, sub_ppm=10000, n_ppm=1000)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
for hint in (0, 600_000_000):
    t = capi.Table(k, capacity_hint=hint)
    l0 = int(capi.lib.oxg_launch_count())
    t0 = time.perf_counter(); st, total, er, ep = t.consume_batch_device(d_bases, d_offs, n, tb, True); dt = time.perf_counter() - t0
    ms, nl = t.last_consume_kernel_ms()
    print(f"hint={hint}: wall {dt*1e3:.0f} ms, consume kernels {ms:.0f} ms in {nl} launches, all launches {int(capi.lib.oxg_launch_count())-l0}, slots {t.capacity}, {total/dt/1e9:.2f} G/s", flush=True)
    t.close()
