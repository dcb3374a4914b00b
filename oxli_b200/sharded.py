"""Hash-sharded k-mer count table across the GPUs of one node: the Python face of the
`oxg_shard_*` entry points (include/oxli_b200.h).

The table is sharded by the high bits of the hash -- owner(h) = h >> (64 - log2 N) -- as
BASELINE.json's north_star prescribes; every rank hashes its own reads, leaves the hashes in
fragments per (owner, table partition) in its own HBM, and each owner's aggregation kernel
pulls its fragments from all ranks over NVLink (peer-mapped memory) and counts them.  All of
that, including the rank-to-rank flags and the reductions (len / sum / min / max, histo,
|A & B|, jaccard: reference src/lib.rs:464-539, 610-638, 708-722), lives in the C library.

Nothing here imports torch.  A launcher (bench.py, the tests, a user's torchrun script)
provides ONE thing: `exchange(blob: bytes) -> list[bytes]`, an all-gather of a 64-byte CUDA IPC
handle among the ranks at construction time -- any transport will do (torch.distributed,
MPI, a pipe).  Ranks that live in one process use `ShardedTable.local(...)` instead.

The per-rank engine is injectable, so the host-side logic in this file (owner function, the
split of a read set among ranks, zero-filled histograms, error mapping) is testable on CPU
with world_size 2 over gloo; the product engine is `_capi.Shard`.
"""
from __future__ import annotations

import numpy as np


def owner_of(h: np.ndarray | int, world: int):
    """Shard that owns hash h: the top log2(world) bits."""
    if world == 1:
        return h * 0 if isinstance(h, np.ndarray) else 0
    shift = 64 - (world.bit_length() - 1)
    if isinstance(h, np.ndarray):
        return (np.asarray(h, dtype=np.uint64) >> np.uint64(shift)).astype(np.int64)
    return int(h) >> shift


def split_reads(n_reads: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of a read set that rank `rank` ingests (any split works: counts add)."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


class BadKmerError(ValueError):
    """consume met a non-ACGT window in error mode (reference: src/lib.rs:593-596)."""

    def __init__(self, position: int, read: int):
        super().__init__(f"bad k-mer encountered at position {position}")
        self.position, self.read = position, read


class ShardedTable:
    """Rank-local handle on a hash-sharded count table.  Methods marked (collective) must be
    called by every rank."""

    def __init__(self, ksize: int, rank: int, world: int, device: int = 0, exchange=None,
                 capacity_hint: int = 0, round_windows: int = 0, engine=None):
        assert world >= 1 and world & (world - 1) == 0, "number of shards must be a power of two"
        self.ksize, self.rank, self.world = ksize, rank, world
        if engine is None:
            from . import _capi as capi  # fails loudly without the CUDA library: there is no CPU engine in the product

            engine = capi.Shard(ksize, rank, world, device=device, capacity_hint=capacity_hint, round_windows=round_windows)
        self.engine = engine
        if world > 1 and exchange is not None:
            handles = exchange(engine.export_handle())
            assert len(handles) == world
            engine.connect(list(handles))

    @classmethod
    def local(cls, ksize: int, world: int, devices: list[int], capacity_hint: int = 0, round_windows: int = 0):
        """All shards in this process (one driver over a node's GPUs, or N > 1 on one GPU in tests).
        Collective calls must then be issued from one thread per shard."""
        from . import _capi as capi

        shards = [cls(ksize, r, world, device=devices[r % len(devices)], capacity_hint=capacity_hint,
                      round_windows=round_windows) for r in range(world)]
        capi.Shard.connect_local([s.engine for s in shards])
        return shards

    def close(self):
        self.engine.close()

    # -- ingest (collective) ----------------------------------------------------------
    def consume_batch(self, bases: np.ndarray, offsets: np.ndarray, skip_bad_kmers: bool = True) -> int:
        """This rank's reads (CSR batch in host memory).  Returns the number of k-mers counted
        from them, i.e. the sum of what the reference's consume would return per read."""
        st, counted, absorbed, er, ep = self.engine.consume_batch(bases, offsets, skip_bad_kmers)
        self.last_absorbed = absorbed
        if st != 0:
            raise BadKmerError(ep, er)
        return counted

    def consume_batch_device(self, d_bases: int, d_offsets: int, n_reads: int, total_bases: int,
                             skip_bad_kmers: bool = True) -> int:
        st, counted, absorbed, er, ep = self.engine.consume_batch_device(d_bases, d_offsets, n_reads, total_bases, skip_bad_kmers)
        self.last_absorbed = absorbed
        if st != 0:
            raise BadKmerError(ep, er)
        return counted

    # -- reductions (collective) ------------------------------------------------------
    def stats(self) -> dict:
        return self.engine.stats()

    def __len__(self) -> int:
        return self.stats()["len"]

    def histo(self, zero: bool = True) -> list[tuple[int, int]]:
        sparse = self.engine.histo()
        if not zero:
            return sparse
        merged = dict(sparse)
        top = sparse[-1][0] if sparse else 0
        return [(f, merged.get(f, 0)) for f in range(top + 1)]  # src/lib.rs:475-480

    def setop_sizes(self, other: "ShardedTable") -> tuple[int, int]:
        return self.engine.setop_sizes(other.engine)

    def jaccard(self, other: "ShardedTable") -> float:
        return self.engine.jaccard(other.engine)

    def digest(self) -> dict:
        """{n, sum, xor, sum_hc, foreign} of the whole table; foreign must be 0."""
        return self.engine.digest()

    # -- this rank's shard only ---------------------------------------------------------
    def local_items_sorted(self):
        return self.engine.table.export(1)
