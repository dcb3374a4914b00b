"""N > 1 parity of the hash-sharded table against the unsharded oracle, bit-exact.

The shards of a table normally sit on one GPU each.  They do not have to: several shards on
ONE device run exactly the same code (pass A fragments per (owner, partition), flag handshake,
pass B pulling every rank's fragments through mapped pointers), so these tests prove the N > 1
path on a single-GPU box too -- threads in one process (oxg_shard_connect_local) and separate
processes connected through CUDA IPC handles (oxg_shard_connect, the arrangement bench.py
uses under torchrun).  With more GPUs visible the same tests spread the shards over them."""
import os
import sys
import threading

import numpy as np
import pytest

from oracle import OracleTable
from synth import ragged_batch, synth_reads, uniform_offsets

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from oxli_b200 import _capi

    assert _capi.lib.oxg_device_count() > 0, "GPU tests need a CUDA device"
    return _capi


def run_ranks(fns):
    """one thread per rank; re-raise the first failure"""
    errs = [None] * len(fns)

    def wrap(i):
        try:
            fns[i]()
        except BaseException as e:  # noqa: BLE001
            errs[i] = e

    ts = [threading.Thread(target=wrap, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ts), "a rank hung"
    for e in errs:
        if e is not None:
            raise e


def devices_for(capi, world):
    n = capi.lib.oxg_device_count()
    return [r % n for r in range(world)]


def gather_shards(shards):
    ks, vs = zip(*[s.local_items_sorted() for s in shards])
    allk, allv = np.concatenate(ks), np.concatenate(vs)
    order = np.argsort(allk, kind="stable")
    return allk[order], allv[order]


@pytest.mark.parametrize("world,k,round_windows", [(2, 21, 0), (2, 31, 1 << 16), (4, 31, 1 << 17), (8, 21, 1 << 16)])
def test_shards_in_one_process_vs_oracle(capi, world, k, round_windows):
    from oxli_b200.sharded import ShardedTable, owner_of, split_reads

    n, L, G = 24_000, 150, 150_000
    bases = synth_reads(n, L, G, seed=77, sub_ppm=5000, n_ppm=800)
    offs = uniform_offsets(n, L)
    shards = ShardedTable.local(k, world, devices_for(capi, world), round_windows=round_windows)
    counted = [0] * world

    def rank_fn(r):
        def go():
            lo, hi = split_reads(n, r, world)
            sub = offs[lo:hi + 1]
            counted[r] = shards[r].consume_batch(bases, sub)          # host buffers, several rounds when the round is small
            counted[r] += shards[r].consume_batch(bases, sub)         # twice: every count doubles
        return go

    run_ranks([rank_fn(r) for r in range(world)])
    truth = OracleTable(k)
    want = 2 * truth.consume_batch(bases, offs, True, nthreads=4)[0]
    assert sum(counted) == want
    assert sum(s.last_absorbed for s in shards) * 2 == want
    for r, s in enumerate(shards):
        keys, _ = s.local_items_sorted()
        assert np.all(owner_of(keys, world) == r), "a key sits on a shard that does not own it"
    gk, gv = gather_shards(shards)
    tk, tv = truth.items_sorted()
    assert np.array_equal(gk, tk) and np.array_equal(gv, 2 * tv)

    # reductions: every rank gets the whole table's answer
    out = [None] * world

    def red_fn(r):
        def go():
            out[r] = (shards[r].stats(), shards[r].histo(zero=False), shards[r].digest())
        return go

    run_ranks([red_fn(r) for r in range(world)])
    with np.errstate(over="ignore"):
        want_digest = {"n": len(tk), "sum": int((2 * tv).sum(dtype=np.uint64)), "xor": int(np.bitwise_xor.reduce(tk)),
                       "sum_hc": int((tk * (2 * tv)).sum(dtype=np.uint64)), "foreign": 0}
    for r in range(world):
        st, hi_, dg = out[r]
        assert st == {"len": len(truth), "sum": 2 * truth.sum_counts, "min": 2 * truth.min, "max": 2 * truth.max}
        assert hi_ == [(2 * f, c) for f, c in truth.histo(zero=False)]
        assert dg == want_digest
    for s in shards:
        s.close()


def test_uneven_ranks_error_mode_and_set_comparison(capi):
    from oxli_b200.sharded import BadKmerError, ShardedTable

    world, k = 2, 31
    rng = np.random.default_rng(5)
    bases, offs = ragged_batch(rng, 3000, 260, p_bad=0.0, p_empty=0.05)
    bad = bases.copy()
    r_bad = next(r for r in range(1700, 3000) if int(offs[r + 1]) - int(offs[r]) > 80)
    p_bad = int(offs[r_bad]) + 40
    bad[p_bad] = ord("N")
    a = ShardedTable.local(k, world, devices_for(capi, world), round_windows=1 << 16)
    b = ShardedTable.local(k, world, devices_for(capi, world), round_windows=1 << 16)
    res = {}

    def rank0():
        # all reads on rank 0, error mode: stops in read r_bad
        try:
            a[0].consume_batch(bad, offs, skip_bad_kmers=False)
            res["err"] = None
        except BadKmerError as e:
            res["err"] = (e.read, e.position)
        b[0].consume_batch(bases, offs[:1001])
        res["j0"] = a[0].jaccard(b[0]); res["s0"] = a[0].setop_sizes(b[0])

    def rank1():
        # no reads at all on this rank: it still takes part in every round
        assert a[1].consume_batch(bases[:0], offs[:1], skip_bad_kmers=False) == 0
        b[1].consume_batch(bases, offs[1000:2001])
        res["j1"] = a[1].jaccard(b[1]); res["s1"] = a[1].setop_sizes(b[1])

    run_ranks([rank0, rank1])
    ta, tb = OracleTable(k), OracleTable(k)
    want = ta.consume_batch(bad, offs, skip_bad_kmers=False)
    tb.consume_batch(bases, offs[:2001])
    assert res["err"] == (want[1], want[2]) and want[1] == r_bad
    gk, gv = gather_shards(a)
    tk, tv = ta.items_sorted()
    assert np.array_equal(gk, tk) and np.array_equal(gv, tv)
    assert res["s0"] == res["s1"] == ta.setop_sizes(tb)
    assert res["j0"] == res["j1"] == ta.jaccard(tb)
    for s in a + b:
        s.close()


def test_device_resident_batch_low_complexity_and_growth(capi):
    from oxli_b200.sharded import ShardedTable

    world, k = 4, 21
    devs = devices_for(capi, world)
    n, L = 30_000, 150
    shards = ShardedTable.local(k, world, devs, round_windows=1 << 18)  # nothing hinted: the shards grow
    per_rank = []
    for r in range(world):
        if r == 1:
            b = np.frombuffer(b"A" * L * n, dtype=np.uint8)  # one k-mer floods one partition of one owner
        else:
            b = synth_reads(n, L, 40_000_000, seed=100 + r)   # nearly all distinct
        per_rank.append(b)
    offs = uniform_offsets(n, L)
    counted = [0] * world
    # device buffers are made before the ranks start: cudaMalloc / cudaFree synchronise the device,
    # and a rank that allocates while another one (same GPU here) waits for it in a kernel would
    # deadlock -- the library itself allocates nothing inside a consume call for the same reason
    bufs = []
    for r in range(world):
        d_b = capi.device_alloc(n * L + 64, devs[r]); d_o = capi.device_alloc((n + 1) * 8, devs[r])
        capi.h2d(d_b, per_rank[r], devs[r]); capi.h2d(d_o, offs, devs[r])
        bufs.append((d_b, d_o))

    def rank_fn(r):
        def go():
            counted[r] = shards[r].consume_batch_device(bufs[r][0], bufs[r][1], n, n * L)
        return go

    run_ranks([rank_fn(r) for r in range(world)])
    for r, (d_b, d_o) in enumerate(bufs):
        capi.device_free(d_b, devs[r]); capi.device_free(d_o, devs[r])
    truth = OracleTable(k)
    want = sum(truth.consume_batch(b, offs, True, nthreads=4)[0] for b in per_rank)
    assert sum(counted) == want
    gk, gv = gather_shards(shards)
    tk, tv = truth.items_sorted()
    assert np.array_equal(gk, tk) and np.array_equal(gv, tv)
    for s in shards:
        s.close()


def _ipc_worker(rank, world, k, conns, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oxli_b200 import _capi as capi
    from oxli_b200.sharded import ShardedTable, owner_of, split_reads
    from synth import synth_reads, uniform_offsets

    def exchange(blob):
        # star all-gather over pipes: rank 0 collects and hands the list back
        if rank == 0:
            got = [blob] + [c.recv() for c in conns]
            for c in conns:
                c.send(got)
            return got
        conns.send(blob)
        return conns.recv()

    dev = rank % capi.lib.oxg_device_count()
    n, L, G = 20_000, 150, 120_000
    bases = synth_reads(n, L, G, seed=9, n_ppm=500)
    offs = uniform_offsets(n, L)
    t = ShardedTable(k, rank, world, device=dev, exchange=exchange, round_windows=1 << 17)
    lo, hi = split_reads(n, rank, world)
    counted = t.consume_batch(bases, offs[lo:hi + 1])
    keys, vals = t.local_items_sorted()
    assert np.all(owner_of(keys, world) == rank)
    st, dg = t.stats(), t.digest()
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"), keys=keys, vals=vals, counted=counted,
             stats=np.array([st["len"], st["sum"], st["min"], st["max"]], dtype=np.uint64),
             digest=np.array([dg["n"], dg["sum"], dg["xor"], dg["sum_hc"], dg["foreign"]], dtype=np.uint64))
    exchange(b"done")  # nobody unmaps an exchange area a peer may still be polling
    t.close()


@pytest.mark.parametrize("world", [2, 4])
def test_shards_in_separate_processes_over_cuda_ipc(capi, tmp_path, world):
    import multiprocessing as mp

    k = 31
    ctx = mp.get_context("spawn")
    pipes = [ctx.Pipe() for _ in range(world - 1)]
    procs = [ctx.Process(target=_ipc_worker, args=(0, world, k, [p[0] for p in pipes], str(tmp_path)))]
    procs += [ctx.Process(target=_ipc_worker, args=(r, world, k, pipes[r - 1][1], str(tmp_path))) for r in range(1, world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    n, L, G = 20_000, 150, 120_000
    bases = synth_reads(n, L, G, seed=9, n_ppm=500)
    truth = OracleTable(k)
    want = truth.consume_batch(bases, uniform_offsets(n, L), True, nthreads=4)[0]
    parts = [np.load(tmp_path / f"shard{r}.npz") for r in range(world)]
    assert sum(int(p["counted"]) for p in parts) == want
    allk = np.concatenate([p["keys"] for p in parts]); allv = np.concatenate([p["vals"] for p in parts])
    order = np.argsort(allk)
    tk, tv = truth.items_sorted()
    assert np.array_equal(allk[order], tk) and np.array_equal(allv[order], tv)
    for p in parts:
        assert list(p["stats"]) == [len(truth), truth.sum_counts, truth.min, truth.max]
        assert int(p["digest"][0]) == len(tk) and int(p["digest"][4]) == 0
