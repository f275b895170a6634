"""Times the Gaussian NLL pass (rlvi_loss_f64, kind GAUSSIAN) at d = 64: python tools/gauss_time.py [log2n] [reps]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rlvi_b200 import ops

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
n, d = 1 << log2n, 64
X = torch.randn((n, d), device=dev, dtype=torch.float64)
w = torch.rand(n, device=dev, dtype=torch.float64)
rng = np.random.default_rng(0)
U = np.triu(rng.normal(size=(d, d))) * 0.1 + np.eye(d)
params = torch.from_numpy(np.concatenate([[1.5], rng.normal(size=d) * 0.1, U.reshape(-1)])).to(dev)
for _ in range(3):
    ops.loss(ops.LOSS_GAUSSIAN, X, params, weights=w, want_losses=False, want_e=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ops.loss(ops.LOSS_GAUSSIAN, X, params, weights=w, want_losses=False, want_e=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"gaussian nll n=2^{log2n} d={d} {ms:.3f} ms  {n * (d * 8 + 16) / ms / 1e6:.0f} GB/s  {n * 4608 / ms / 1e9:.2f} TFLOP/s", flush=True)
