"""Raw pinned-host -> device copy bandwidth of the box (what bounds bench.py's e2e leg)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 64 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for rep in range(3):
        for o in range(0, n, chunk): d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"H2D {n / dt / 1e9:.1f} GB/s in chunks of {chunk >> 20} MiB")
