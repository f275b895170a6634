"""Per-pass time of the persistent fixed-point kernel vs shard size (tol = 0 -> exactly `maxiter` passes):
python tools/fp_pass_time.py   -> one line per size: us per pass.  RLVI_FP_CACHE_SLOTS as in fixed_point.cu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlvi_b200 import ops

dev = torch.device("cuda", 0)
for log2n in (16, 18, 20, 22, 23, 24, 26):
    n = 1 << log2n
    g = torch.Generator(device=dev).manual_seed(1)
    e = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 0.5
    ew = e.clone()
    out = torch.empty_like(e)
    res = None
    times = {}
    for iters in (10, 30):
        for _ in range(2):
            ew.copy_(e)
            _, res = ops.fixed_point(None, e_work=ew, tol=0.0, maxiter=iters, out=out, result=res)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            ops.fixed_point(None, e_work=ew, tol=0.0, maxiter=iters, out=out, result=res)
        b.record()
        torch.cuda.synchronize()
        times[iters] = a.elapsed_time(b) / reps
    per_pass = (times[30] - times[10]) / 20 * 1e3
    print(f"n=2^{log2n}: {per_pass:.2f} us per pass (10 passes {times[10]:.3f} ms, 30 passes {times[30]:.3f} ms); "
          f"streaming floor at 6538 GB/s: {n * 8 / 6538e9 * 1e6:.2f} us", flush=True)
