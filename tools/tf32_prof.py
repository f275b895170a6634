"""Small driver for ncu: a few launches of the TF32 Gram and the FP32 loss at config-3 width (python tools/tf32_prof.py [log2n] [d])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rlvi_b200 import ops

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
n = 1 << log2n
X = torch.randn((n, d), device=dev, dtype=torch.float32)
w = torch.rand(n, device=dev, dtype=torch.float64)
th = torch.randn(d, device=dev, dtype=torch.float64)
out = None
for prec in (ops.TF32X3, ops.TF32X1, ops.TF32X3, ops.TF32X1):
    out = ops.weighted_moments(X, w, power=2, precision=prec, out=out)
ops.loss(ops.LOSS_PCA, X, th)
ops.loss(ops.LOSS_PCA, X, th)
torch.cuda.synchronize()
print("ok")
