"""Times rlvi_weighted_moments_f64: python tools/gram64_time.py [log2n] [reps] [d]  (d = 64: gram64_kernel, else gram_tma_kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlvi_b200 import ops

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
n, d = 1 << log2n, (int(sys.argv[3]) if len(sys.argv) > 3 else 64)
X = torch.randn((n, d), device=dev, dtype=torch.float64)
w = torch.rand(n, device=dev, dtype=torch.float64)
out = None
for _ in range(3):
    out = ops.weighted_moments(X, w, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ops.weighted_moments(X, w, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"gram f64 n=2^{log2n} d={d} {ms:.3f} ms  {n * (d + 1) * 8 / ms / 1e6:.0f} GB/s  {n * d * (d + 1) / ms / 1e9:.2f} TFLOP/s useful", flush=True)
