"""world_size-2 (and 4) test of the sharded driver's host logic on CPU (gloo):
owner function, length/hash exchange, and the cross-shard reductions.  The
per-rank compute is an oracle-backed stand-in for CudaShardEngine -- the
exchange and reduction code under test is the product's (oxli_b200/sharded.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    def __init__(self, ksize, rank, world):
        from oracle import OracleTable

        self.table, self.ksize, self.rank, self.world = OracleTable(ksize), ksize, rank, world

    def route_host(self, bases, offsets):
        import torch

        import oracle
        from oxli_b200.sharded import RouteResult, owner_of

        hs = [oracle.hash_windows(bases[int(offsets[r]):int(offsets[r + 1])], self.ksize) for r in range(len(offsets) - 1)]
        h = np.concatenate(hs) if hs else np.zeros(0, dtype=np.uint64)
        h = h[h != 0]
        own = owner_of(h, self.world)
        for x in h[own == self.rank]:
            self.table.count_hash(int(x))
        out = [torch.from_numpy(h[own == r].view(np.int64).copy()) if r != self.rank else torch.zeros(0, dtype=torch.int64)
               for r in range(self.world)]
        return RouteResult(int((own == self.rank).sum()), out)

    def new_buffer(self, n):
        import torch

        return torch.empty(n, dtype=torch.int64)

    def count(self, t):
        for x in t.numpy().view(np.uint64):
            self.table.count_hash(int(x))
        return t.numel()

    def stats(self):
        return {"len": len(self.table), "sum": self.table.sum_counts, "min": self.table.min, "max": self.table.max}

    def histo(self):
        return self.table.histo(zero=False)

    def setop_sizes(self, other):
        return self.table.setop_sizes(other.table)


def _worker(rank, world, port, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import OracleTable
    from oracle.synth import ragged_batch
    from oxli_b200.sharded import ShardedCounter, owner_of

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1234)  # same data everywhere; each rank takes its slice of the reads
        bases, offs = ragged_batch(rng, 400, 180, p_bad=0.01)
        bases2, offs2 = ragged_batch(rng, 300, 180, p_bad=0.0)
        n = len(offs) - 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        a = ShardedCounter(OracleShardEngine(k, rank, world))
        absorbed = a.consume_routed(a.engine.route_host(bases, offs[lo:hi + 1]))
        b = ShardedCounter(OracleShardEngine(k, rank, world))
        n2 = len(offs2) - 1
        b.consume_routed(b.engine.route_host(bases2, offs2[n2 * rank // world: n2 * (rank + 1) // world + 1]))
        # a second batch on top of `a`: half of bases2
        a.consume_routed(a.engine.route_host(bases2, offs2[: n2 // 2 + 1] if rank == 0 else offs2[:1]))

        # unsharded truth
        ta, tb = OracleTable(k), OracleTable(k)
        ta.consume_batch(bases, offs); ta.consume_batch(bases2, offs2[: n2 // 2 + 1])
        tb.consume_batch(bases2, offs2)

        # every key sits on its owner and nowhere else; shards reassemble the truth
        keys, vals = a.engine.table.items_sorted()
        assert np.all(owner_of(keys, world) == rank)
        parts = [None] * world
        dist.all_gather_object(parts, (keys, vals))
        allk = np.concatenate([p[0] for p in parts]); allv = np.concatenate([p[1] for p in parts])
        order = np.argsort(allk)
        tk, tv = ta.items_sorted()
        assert np.array_equal(allk[order], tk) and np.array_equal(allv[order], tv)
        tot = [None] * world
        dist.all_gather_object(tot, absorbed)
        o1 = OracleTable(k)
        assert sum(tot) == o1.consume_batch(bases, offs)[0]

        s = a.stats()
        assert s == {"len": len(ta), "sum": ta.sum_counts, "min": ta.min, "max": ta.max}
        assert len(a) == len(ta)
        assert a.histo(zero=False) == ta.histo(zero=False) and a.histo() == ta.histo(zero=True)
        assert a.setop_sizes(b) == ta.setop_sizes(tb)
        assert a.jaccard(b) == ta.jaccard(tb)
        e1 = ShardedCounter(OracleShardEngine(k, rank, world)); e2 = ShardedCounter(OracleShardEngine(k, rank, world))
        assert e1.jaccard(e2) == 1.0 and e1.stats() == {"len": 0, "sum": 0, "min": 0, "max": 0}
        assert e1.histo() == [(0, 0)] and e1.histo(zero=False) == []
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


def test_owner_function():
    from oxli_b200.sharded import owner_of

    h = np.array([0, 2**61, 2**63, 2**64 - 1], dtype=np.uint64)
    assert list(owner_of(h, 2)) == [0, 0, 1, 1]
    assert list(owner_of(h, 8)) == [0, 1, 4, 7]
    assert list(owner_of(h, 1)) == [0, 0, 0, 0] and owner_of(2**64 - 1, 4) == 3
