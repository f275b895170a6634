"""Per-round timeline of the persistent fixed-point kernel (RLVI_FP_TRACE=1): python tools/fp_trace.py [log2n] [passes]."""
import os
import sys

os.environ["RLVI_FP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlvi_b200 import ops

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
n = 1 << log2n
g = torch.Generator(device=dev).manual_seed(1)
e = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 0.5
out = torch.empty_like(e)
for rep in range(2):
    ew = e.clone()
    torch.cuda.synchronize()
    print(f"--- call {rep}", file=sys.stderr, flush=True)
    ops.fixed_point(None, e_work=ew, tol=0.0, maxiter=passes, out=out)
    torch.cuda.synchronize()
