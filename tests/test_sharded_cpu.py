"""world_size-2 (and 4) test of the sharded table's host logic on CPU (gloo): owner function,
the split of reads among ranks, zero-filled histograms, error mapping -- oxli_b200/sharded.py
with an oracle-backed engine standing in for `_capi.Shard`.  The engine models the product's
exchange (hash locally, hand every hash to its owner, count there; reductions by all-gather)
with gloo collectives, so the test also pins what the GPU path has to reproduce."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    """Same surface as oxli_b200._capi.Shard, CPU oracle inside, gloo between ranks."""

    def __init__(self, ksize, rank, world):
        from oracle import OracleTable

        self.table, self.ksize, self.rank, self.n_ranks = OracleTable(ksize), ksize, rank, world

    def export_handle(self):
        return bytes([self.rank]) * 64

    def connect(self, handles):
        assert [h[0] for h in handles] == list(range(self.n_ranks))

    def consume_batch(self, bases, offsets, skip_bad=True):
        import torch.distributed as dist

        import oracle
        from oxli_b200.sharded import owner_of

        hs, counted, err = [], 0, (-1, 0)
        for r in range(len(offsets) - 1):
            h = oracle.hash_windows(bases[int(offsets[r]):int(offsets[r + 1])], self.ksize)
            if not skip_bad and np.any(h == 0):
                first = int(np.flatnonzero(h == 0)[0])  # (hash 0 marks a bad window here)
                hs.append(h[:first]); counted += first
                err = (r, first)
                break
            h = h[h != 0]
            hs.append(h); counted += len(h)
        h = np.concatenate(hs) if hs else np.zeros(0, dtype=np.uint64)
        own = owner_of(h, self.n_ranks)
        outgoing = [h[own == r] for r in range(self.n_ranks)]
        gathered = [None] * self.n_ranks
        dist.all_gather_object(gathered, outgoing)
        absorbed = 0
        for src in range(self.n_ranks):
            mine = gathered[src][self.rank]
            for x in mine:
                self.table.count_hash(int(x))
            absorbed += len(mine)
        return (3 if err[0] >= 0 else 0), counted, absorbed, err[0], err[1]

    def _gather(self, obj):
        import torch.distributed as dist

        out = [None] * self.n_ranks
        dist.all_gather_object(out, obj)
        return out

    def stats(self):
        parts = self._gather((len(self.table), self.table.sum_counts, self.table.min, self.table.max))
        n = sum(p[0] for p in parts)
        live = [p for p in parts if p[0]]
        return {"len": n, "sum": sum(p[1] for p in parts) % (1 << 64), "min": min(p[2] for p in live) if live else 0,
                "max": max(p[3] for p in live) if live else 0}

    def histo(self):
        merged = {}
        for part in self._gather(self.table.histo(zero=False)):
            for f, c in part:
                merged[f] = merged.get(f, 0) + c
        return sorted(merged.items())

    def setop_sizes(self, other):
        parts = self._gather(self.table.setop_sizes(other.table))
        return sum(p[0] for p in parts), sum(p[1] for p in parts)

    def jaccard(self, other):
        i, u = self.setop_sizes(other)
        return 1.0 if u == 0 else float(np.float64(i) / np.float64(u))


def _worker(rank, world, port, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import OracleTable
    from oracle.synth import ragged_batch
    from oxli_b200.sharded import BadKmerError, ShardedTable, owner_of, split_reads

    dist.init_process_group("gloo", rank=rank, world_size=world)

    def exchange(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    def make():
        return ShardedTable(k, rank, world, exchange=exchange, engine=OracleShardEngine(k, rank, world))

    try:
        rng = np.random.default_rng(1234)  # same data everywhere; each rank takes its slice of the reads
        bases, offs = ragged_batch(rng, 400, 180, p_bad=0.01)
        bases2, offs2 = ragged_batch(rng, 300, 180, p_bad=0.0)
        lo, hi = split_reads(len(offs) - 1, rank, world)
        a = make()
        counted = a.consume_batch(bases, offs[lo:hi + 1])
        b = make()
        n2 = len(offs2) - 1
        lo2, hi2 = split_reads(n2, rank, world)
        b.consume_batch(bases2, offs2[lo2:hi2 + 1])
        # a second batch on top of `a`, all of it from rank 0: the others take part with no reads
        a.consume_batch(bases2, offs2[: n2 // 2 + 1] if rank == 0 else offs2[:1])

        ta, tb = OracleTable(k), OracleTable(k)  # unsharded truth
        want_counted = ta.consume_batch(bases, offs)[0]
        ta.consume_batch(bases2, offs2[: n2 // 2 + 1])
        tb.consume_batch(bases2, offs2)

        # every key sits on its owner and nowhere else; shards reassemble the truth
        keys, vals = a.engine.table.items_sorted()
        assert np.all(owner_of(keys, world) == rank)
        parts = exchange((keys, vals))
        allk = np.concatenate([p[0] for p in parts]); allv = np.concatenate([p[1] for p in parts])
        order = np.argsort(allk)
        tk, tv = ta.items_sorted()
        assert np.array_equal(allk[order], tk) and np.array_equal(allv[order], tv)
        assert sum(exchange(counted)) == want_counted

        assert a.stats() == {"len": len(ta), "sum": ta.sum_counts, "min": ta.min, "max": ta.max}
        assert len(a) == len(ta)
        assert a.histo(zero=False) == ta.histo(zero=False) and a.histo() == ta.histo(zero=True)
        assert a.setop_sizes(b) == ta.setop_sizes(tb)
        assert a.jaccard(b) == ta.jaccard(tb)
        e1, e2 = make(), make()
        assert e1.jaccard(e2) == 1.0 and e1.stats() == {"len": 0, "sum": 0, "min": 0, "max": 0}
        assert e1.histo() == [(0, 0)] and e1.histo(zero=False) == []

        # error mode is per rank: rank 0 stops at its first bad window, the others are unaffected
        c = make()
        seq = np.frombuffer(b"ACGTACGTTTGACCA" * 4 + b"N" + b"ACGTAGGCTAGCTAG" * 4, dtype=np.uint8)
        clean = np.frombuffer(b"GATTACAGATTACACCGGTTAACCGGTAGCAT" * 3, dtype=np.uint8)
        mine, moffs = (seq, np.array([0, len(seq)], dtype=np.uint64)) if rank == 0 else (clean, np.array([0, len(clean)], dtype=np.uint64))
        try:
            got = c.consume_batch(mine, moffs, skip_bad_kmers=False)
            assert rank != 0 and got == len(clean) - k + 1
        except BadKmerError as e:
            assert rank == 0 and e.read == 0 and e.position == 60 - k + 1 and str(e) == f"bad k-mer encountered at position {60 - k + 1}"
        tc = OracleTable(k)
        tc.consume_batch(seq, np.array([0, len(seq)], dtype=np.uint64), skip_bad_kmers=False)
        for _ in range(world - 1):
            tc.consume_batch(clean, np.array([0, len(clean)], dtype=np.uint64))
        assert c.stats() == {"len": len(tc), "sum": tc.sum_counts, "min": tc.min, "max": tc.max}
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,k", [(2, 21), (4, 31)])
def test_sharded_exchange_and_reductions(tmp_path, world, k):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), k, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_owner_function_and_split():
    from oxli_b200.sharded import owner_of, split_reads

    h = np.array([0, 2**61, 2**63, 2**64 - 1], dtype=np.uint64)
    assert list(owner_of(h, 2)) == [0, 0, 1, 1]
    assert list(owner_of(h, 8)) == [0, 1, 4, 7]
    assert list(owner_of(h, 1)) == [0, 0, 0, 0] and owner_of(2**64 - 1, 4) == 3
    for world in (1, 2, 4, 8):
        cuts = [split_reads(1001, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == 1001 and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))


def test_product_package_is_torch_free():
    import re

    pkg = os.path.join(ROOT, "oxli_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(import|from)\s+torch\b", src, re.M), name
