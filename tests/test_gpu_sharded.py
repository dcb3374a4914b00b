"""Two-GPU parity of the sharded CUDA path (NCCL exchange) against the oracle.
Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, k, out_dir, exchange):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    from oracle import OracleTable
    from oracle.synth import synth_reads, uniform_offsets
    from oxli_b200 import _capi as capi
    from oxli_b200.sharded import CudaShardEngine, ShardedCounter, owner_of

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        n, L, G = 60_000, 150, 200_000
        d_bases = capi.device_alloc(n * L + 64, rank)
        d_offs = capi.device_alloc((n + 1) * 8, rank)
        capi.synth_reads_device(d_bases, n, L, G, seed=77, first_read=rank * n, sub_ppm=5000, n_ppm=800, device=rank)
        capi.h2d(d_offs, uniform_offsets(n, L), rank)
        eng = CudaShardEngine(k, rank, world, rank, capacity_hint=G // world, exchange=exchange)
        sc = ShardedCounter(eng)
        absorbed = sc.consume_device(d_bases, d_offs, n, n * L)
        absorbed += sc.consume_device(d_bases, d_offs, n, n * L)  # twice: counts double

        truth = OracleTable(k)
        want_total = 0
        for r in range(world):
            b = synth_reads(n, L, G, seed=77, first_read=r * n, sub_ppm=5000, n_ppm=800)
            for _ in range(2):
                want_total += truth.consume_batch(b, uniform_offsets(n, L), True, nthreads=4)[0]
        keys, vals = eng.items_sorted()
        assert np.all(owner_of(keys, world) == rank)
        parts, tots = [None] * world, [None] * world
        dist.all_gather_object(parts, (keys, vals))
        dist.all_gather_object(tots, absorbed)
        allk = np.concatenate([p[0] for p in parts]); allv = np.concatenate([p[1] for p in parts])
        order = np.argsort(allk)
        tk, tv = truth.items_sorted()
        assert sum(tots) == want_total
        assert np.array_equal(allk[order], tk) and np.array_equal(allv[order], tv)
        assert sc.stats() == {"len": len(truth), "sum": truth.sum_counts, "min": truth.min, "max": truth.max}
        assert sc.histo(zero=False) == truth.histo(zero=False)
        other = ShardedCounter(CudaShardEngine(k, rank, world, rank, exchange=exchange))
        other.consume_device(d_bases, d_offs, n // 2, (n // 2) * L)
        t2 = OracleTable(k)
        for r in range(world):
            b = synth_reads(n // 2, L, G, seed=77, first_read=r * n, sub_ppm=5000, n_ppm=800)
            t2.consume_batch(b, uniform_offsets(n // 2, L), True, nthreads=4)
        assert sc.setop_sizes(other) == truth.setop_sizes(t2)
        assert sc.jaccard(other) == truth.jaccard(t2)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("k,exchange", [(21, "p2p"), (31, "p2p"), (31, "nccl")])
def test_two_gpu_sharded_counting(tmp_path, k, exchange):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), k, str(tmp_path), exchange), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


@pytest.mark.parametrize("k,n_ranks,me", [(21, 2, 1), (31, 4, 2), (31, 8, 0)])
def test_route_kernel_on_one_gpu(k, n_ranks, me):
    """One rank's fused hash + count-own + route launch with local buffers standing in for the
    peers' receive regions: the shard keeps exactly the hashes it owns, every other hash lands in
    its owner's list, and a second launch absorbs delivered lists while it routes."""
    import ctypes as C

    from oracle import OracleTable
    from oracle.synth import ragged_batch
    from oxli_b200 import _capi as capi
    from oxli_b200.sharded import owner_of

    rng = np.random.default_rng(7 * k + n_ranks)
    bases, offs = ragged_batch(rng, 6000, 300, p_bad=0.004)
    truth = OracleTable(k)
    want_total = truth.consume_batch(bases, offs, True, nthreads=4)[0]
    tk, tv = truth.items_sorted()
    own = owner_of(tk, n_ranks)

    d_bases = capi.device_alloc(bases.nbytes + 64)
    d_offs = capi.device_alloc(offs.nbytes)
    capi.h2d(d_bases, bases); capi.h2d(d_offs, offs)
    cap = want_total + 1024
    outs = [capi.device_alloc(cap * 8) for _ in range(n_ranks)]
    d_cnt = capi.device_alloc(n_ranks * 8)
    ptrs = (C.c_void_p * n_ranks)(*outs)

    def route(t, absorb):
        hc = (C.c_uint64 * n_ranks)(); loc = C.c_uint64(); ab = C.c_uint64()
        a_ptrs = (C.c_void_p * max(len(absorb), 1))(*[p for p, _ in absorb])
        a_n = (C.c_uint64 * max(len(absorb), 1))(*[m for _, m in absorb])
        capi.check(capi.lib.oxg_route_batch_device(t.handle, d_bases, d_offs, len(offs) - 1, 0, bases.nbytes, n_ranks, me,
                                                   ptrs, cap, d_cnt, hc, C.byref(loc), len(absorb), a_ptrs, a_n, C.byref(ab)))
        return [int(x) for x in hc], loc.value, ab.value

    try:
        t = capi.Table(k)
        hc, local, absorbed = route(t, [])
        assert absorbed == 0 and hc[me] == 0
        assert local == int(tv[own == me].sum()) and local + sum(hc) == want_total
        gk, gv = t.export(sort_mode=1)
        assert np.array_equal(gk, tk[own == me]) and np.array_equal(gv, tv[own == me])
        lists = {}
        for r in range(n_ranks):
            if r == me:
                continue
            got = np.empty(hc[r], dtype=np.uint64)
            if hc[r]:
                capi.d2h(got, outs[r])
            lists[r] = got
            assert np.array_equal(np.sort(got), np.repeat(tk[own == r], tv[own == r].astype(np.int64)))
        # second table: route the same reads and absorb the lists as if peers had delivered them
        # (copies: the launch overwrites outs[] while it reads the absorb segments)
        segs = []
        for r, got in lists.items():
            if len(got):
                d = capi.device_alloc(got.nbytes); capi.h2d(d, got); segs.append((d, len(got)))
        t2 = capi.Table(k)
        hc2, local2, absorbed2 = route(t2, segs)
        assert (hc2, local2) == (hc, local) and absorbed2 == sum(hc)
        g2k, g2v = t2.export(sort_mode=1)
        assert np.array_equal(g2k, tk) and np.array_equal(g2v, tv)
        for d, _ in segs:
            capi.device_free(d)
    finally:
        for o in outs:
            capi.device_free(o)
        capi.device_free(d_cnt); capi.device_free(d_bases); capi.device_free(d_offs)
