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
