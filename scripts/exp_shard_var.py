#!/usr/bin/env python
"""Step-to-step spread of the sharded C3 step under torchrun (device-resident reads).
Runs STEPS steps three times: nothing else running, one `nvidia-smi -lms 100` per rank (what
bench.py's clock sampler used to do on every rank), and one sampler on rank 0 only.
Prints per-step ms (CUDA events, max over ranks / wall, max over ranks)."""
import os
import subprocess
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
steps = int(os.environ.get("STEPS", 10))
G = n * world
d_b = capi.device_alloc(n * L + 64, local); d_o = capi.device_alloc((n + 1) * 8, local)
capi.synth_reads_device(d_b, n, L, G, 0xC30001, first_read=rank * n, sub_ppm=10_000, n_ppm=1_000, device=local)
capi.h2d(d_o, np.arange(n + 1, dtype=np.uint64) * np.uint64(L), local)
Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap"


def exchange(blob):
    out = [None] * world
    dist.all_gather_object(out, blob)
    return out


def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if rank == 0:
    quota = open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "?"
    print(f"host: cpu_count {os.cpu_count()}, affinity {len(os.sched_getaffinity(0))}, cgroup cpu.max {quota}, load {os.getloadavg()}", flush=True)
t = ShardedTable(k, rank, world, device=local, exchange=exchange, round_windows=int(os.environ.get('ROUND_MW', 0)) << 20)
for _ in range(3):
    t.engine.table.clear()
    t.consume_batch_device(d_b, d_o, n, n * L, True)
for name in os.environ.get("SEGMENTS", "quiet,smi on every rank,quiet again,cpu burners,smi 1000 ms,quiet at last").split(","):
    proc = None
    if name.startswith("smi"):
        proc = subprocess.Popen(["nvidia-smi", f"--id={local}", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms",
                                 "1000" if "1000" in name else "100"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    elif name.startswith("cpu"):
        proc = subprocess.Popen([sys.executable, "-c", "import threading\nwhile True: pass"])
    if name != "quiet":
        time.sleep(1.0)  # (let the companion get going)
    print(f"--- {name}", file=sys.stderr, flush=True)
    dist.barrier(); torch.cuda.synchronize()
    ms = []
    for step in range(steps):
        t.engine.table.clear()
        if name.startswith("barrier"):
            dist.barrier()
        t0 = time.perf_counter()
        t.consume_batch_device(d_b, d_o, n, n * L, True)
        wall = 1e3 * (time.perf_counter() - t0)
        ms.append((t.engine.last_ms()[0], wall))
    if proc:
        proc.terminate(); proc.wait()
    ms = [(mx(a), mx(b)) for a, b in ms]
    if rank == 0:
        ev = [a for a, _ in ms]
        print(f"{name:26s} mean {np.mean(ev):7.1f}  min {min(ev):7.1f}  max {max(ev):7.1f} | " + " ".join(f"{a:.0f}/{b:.0f}" for a, b in ms), flush=True)
t.close()
dist.barrier()
dist.destroy_process_group()
