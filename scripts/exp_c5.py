"""C5-shaped run (BASELINE.json configs[4]): histo and intersection / union / jaccard between two large
tables (k=21) built from two disjoint-but-overlapping sets of 10-kbp long reads of one 400 Mbp genome (5x each), with timings and the
size-independent identities between the results."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from oxli_b200 import _capi as capi
n, L, k = (int(sys.argv[1]) if len(sys.argv) > 1 else 200_000), 10_000, 21
G = 400_000_000
tb = n * L
d_bases = capi.device_alloc(tb + 64); d_offs = capi.device_alloc((n + 1) * 8)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
tabs = []
for name, first in (("A", 0), ("B", n // 2)):   # reads [0,n) and [n/2, 3n/2) of the same read stream
    capi.synth_reads_device(d_bases, n, L, G, 0xC50001, first_read=first)
    t = capi.Table(k, capacity_hint=int(n * L * 0.99))
    t0 = time.perf_counter(); st, total, _, _ = t.consume_batch_device(d_bases, d_offs, n, tb, True); dt = time.perf_counter() - t0
    print(f"table {name}: {total/1e9:.2f} G k-mers from {n} x {L}-bp reads in {dt*1e3:.0f} ms ({total/dt/1e9:.1f} G k-mers/s), distinct {len(t)/1e6:.0f} M, {t.capacity*16/2**30:.0f} GiB")
    tabs.append(t)
a, b = tabs
t0 = time.perf_counter(); ha = a.histo(); dt = time.perf_counter() - t0
sa = a.stats()
print(f"histo(A): {len(ha)} bins in {dt*1e3:.1f} ms ({a.capacity*16/dt/1e9:.0f} GB/s over the slot array); head {ha[:3]}")
assert sum(c for _, c in ha) == sa["len"] and sum(f * c for f, c in ha) == sa["sum"]
t0 = time.perf_counter(); inter, uni = a.setop_sizes(b); dt = time.perf_counter() - t0
print(f"|A&B| = {inter}, |A|B| = {uni} in {dt*1e3:.1f} ms ({len(a)/dt/1e9:.1f} G probes/s)")
j = a.jaccard(b)
assert j == inter / uni and uni == len(a) + len(b) - inter
assert b.setop_sizes(a) == (inter, uni) and a.jaccard(a) == 1.0
print(f"jaccard = {j:.6f}  (two 5x samplings of the same genome: nearly every k-mer is in both)")
