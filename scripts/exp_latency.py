"""Per-call latency of the drop-in API (the reference's one-call-per-record usage)."""
import sys, time
sys.path.insert(0, '/root/repo')
import oxli
seq = "".join(l.strip() for l in open('/root/repo/tests/golden/example.fa') if not l.startswith('>'))
t = oxli.KmerCountTable(31)
t.consume(seq[:1000])
def rate(name, fn, n):
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    dt = time.perf_counter() - t0
    print(f"{name:34s} {dt / n * 1e6:8.1f} us/call")
rate("consume(150 bp read)", lambda i: t.consume(seq[i * 7:i * 7 + 150]), 3000)
rate("count(kmer)", lambda i: t.count(seq[i:i + 31]), 3000)
rate("get(kmer)", lambda i: t.get(seq[i:i + 31]), 3000)
rate("hash_kmer(kmer)", lambda i: t.hash_kmer(seq[i:i + 31]), 3000)
rate("get_hash(h)", lambda i: t.get_hash(i + 1), 3000)
reads = [seq[i * 7:i * 7 + 150] for i in range(20000)]
t0 = time.perf_counter(); n = t.consume_many(reads); dt = time.perf_counter() - t0
print(f"consume_many(20000 reads), first   {dt * 1e3:8.2f} ms total ({dt / 20000 * 1e6:.2f} us/read)")
t0 = time.perf_counter(); n = t.consume_many(reads); dt = time.perf_counter() - t0
print(f"consume_many(20000 reads), again   {dt * 1e3:8.2f} ms total ({dt / 20000 * 1e6:.2f} us/read)")
d = oxli.KmerCountTable(31, deferred=True)
rate("consume(150 bp read), deferred=True", lambda i: d.consume(seq[i * 7:i * 7 + 150]), 20000)
t0 = time.perf_counter(); d.flush(); print(f"flush {1e3 * (time.perf_counter() - t0):.2f} ms")
print("len", len(t), "max", t.max, len(d))
