"""C3-shaped run at reduced size (BASELINE.json configs[2]: k=21, 150-bp reads, 1 % substitutions,
0.1 % N, 100 Mbp genome): singleton-heavy table that grows on the device while it is fed.
Checks size-independent properties and prints throughput for skip mode, error mode and a second pass."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from oxli_b200 import _capi as capi
n, L, k, G = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000, 150, 21, 100_000_000
tb = n * L
d_bases = capi.device_alloc(tb + 64); d_offs = capi.device_alloc((n + 1) * 8)
capi.synth_reads_device(d_bases, n, L, G, 0xC30001, sub_ppm=10000, n_ppm=1000)
capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
t = capi.Table(k)  # no capacity hint: starts at 1024 slots and grows
t0 = time.perf_counter(); st, total, er, ep = t.consume_batch_device(d_bases, d_offs, n, tb, True); dt = time.perf_counter() - t0
s = t.stats()
print(f"pass 1 (skip mode, growing from 1024 slots): {total} k-mers in {dt*1e3:.0f} ms = {total/dt/1e9:.2f} G k-mers/s; "
      f"distinct {s['len']}, slots {t.capacity} ({t.capacity*16/2**30:.1f} GiB), load {s['len']/t.capacity:.2f}")
assert st == 0 and s["sum"] == total and s["min"] >= 1
valid_frac = total / (n * (L - k + 1))
print(f"  valid windows {100*valid_frac:.2f} % of all windows (N rate 0.1 % -> expect ~{100*(1-0.001)**k:.2f} %)")
h = t.histo(); assert sum(c for _, c in h) == s["len"] and sum(f * c for f, c in h) == total
print("  histo head:", h[:5])
t0 = time.perf_counter(); st, total2, _, _ = t.consume_batch_device(d_bases, d_offs, n, tb, True); dt = time.perf_counter() - t0
s2 = t.stats()
print(f"pass 2 (all keys present): {total2/dt/1e9:.2f} G k-mers/s")
assert total2 == total and s2["len"] == s["len"] and s2["sum"] == 2 * total and s2["min"] >= 2   # linearity
e = capi.Table(k, capacity_hint=1000)
t0 = time.perf_counter(); st, tot_e, er, ep = e.consume_batch_device(d_bases, d_offs, n, tb, False); dt = time.perf_counter() - t0
print(f"error mode: status {st}, stopped at read {er} window {ep} after {tot_e} k-mers ({dt*1e3:.1f} ms incl. the pre-scan of {tb/1e9:.1f} Gbases)")
assert st == capi.ERR_BAD_KMER and er >= 0 and len(e) <= tot_e
# the prefix counted in error mode equals skip mode on the same prefix
p = capi.Table(k)
p.consume_batch_device(d_bases, d_offs, er, er * L, True)
got = e.setop_sizes(p)
print(f"  error-mode table vs skip-mode table of reads [0,{er}): |A&B|={got[0]} |A|B|={got[1]} (prefix of the failing read adds {len(e)-got[0]} keys)")
assert got[0] == len(p)
print("jaccard(t, t) =", t.jaccard(t))
