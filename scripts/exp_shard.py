#!/usr/bin/env python
"""Sharded C3 step under torchrun, several settings in one launch (device-resident reads):
round size, work items per partition, hint.  Prints ms per step (CUDA events / wall, max over ranks)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oxli_b200 import _capi as capi  # noqa: E402
from oxli_b200.sharded import ShardedTable  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n, L, k = int(os.environ.get("READS", 12_500_000)), 150, 21
G = n * world
d_b = capi.device_alloc(n * L + 64, local); d_o = capi.device_alloc((n + 1) * 8, local)
capi.synth_reads_device(d_b, n, L, G, 0xC30001, first_read=rank * n, sub_ppm=10_000, n_ppm=1_000, device=local)
capi.h2d(d_o, np.arange(n + 1, dtype=np.uint64) * np.uint64(L), local)


def exchange(blob):
    out = [None] * world
    dist.all_gather_object(out, blob)
    return out


def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name, rw, groups, hint in (("default", 0, 0, 0), ("round=256Mi", 256 << 20, 0, 0), ("round=512Mi", 512 << 20, 0, 0),
                               ("round=256Mi hinted 280M/shard", 256 << 20, 0, 280_000_000)):
    capi.set_pipeline("auto", 0, groups)
    t = ShardedTable(k, rank, world, device=local, exchange=exchange, round_windows=rw, capacity_hint=hint)
    ms = []
    for step in range(4):
        t.engine.table.clear()
        dist.barrier()
        t0 = time.perf_counter()
        got = t.consume_batch_device(d_b, d_o, n, n * L, True)
        wall = 1e3 * (time.perf_counter() - t0)
        ms.append((mx(t.engine.last_ms()[0]), mx(wall)))
    if rank == 0:
        print(f"{name:24s} " + "  ".join(f"{a:7.1f}/{b:7.1f}" for a, b in ms) + f"   ms per step (events/wall), rounds {t.engine.last_ms()[1]}, cap {t.engine.table.capacity}", flush=True)
    t.close()
    dist.barrier()
dist.destroy_process_group()
