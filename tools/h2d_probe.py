"""Host->device copy bandwidth per GPU for subsets of the box's GPUs copying at the same time (one process, one stream per
GPU, pinned 1 GiB buffers): names the host-side bound of the multi-GPU `e2e` line.  python tools/h2d_probe.py"""
import json
import time

import torch

n = torch.cuda.device_count()
size = 1 << 30
host = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{g}") for g in range(n)]
streams = [torch.cuda.Stream(device=g) for g in range(n)]


def run(subset, reps=4):
    for g in subset:      # warm-up
        with torch.cuda.stream(streams[g]):
            dev[g].copy_(host[g], non_blocking=True)
    for g in subset:
        streams[g].synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for g in subset:
            with torch.cuda.stream(streams[g]):
                dev[g].copy_(host[g], non_blocking=True)
    for g in subset:
        streams[g].synchronize()
    dt = time.perf_counter() - t0
    return reps * size * len(subset) / dt / 1e9


subsets = [[0]]
if n >= 2:
    subsets += [[0, 1]]
if n >= 4:
    subsets += [[0, 2], [0, 1, 2, 3]]
if n >= 8:
    subsets += [[0, 4], [0, 2, 4, 6], list(range(8))]
for s in subsets:
    total = run(s)
    print(json.dumps({"gpus": s, "total_GBps": round(total, 1), "per_gpu_GBps": round(total / len(s), 1)}), flush=True)
