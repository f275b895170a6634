"""Host-side glue shared by the drop-in modules: move the caller's arrays to the GPU and back.

The reference's calling convention is NumPy in / NumPy out (standard-learning, online-learning) or CUDA
tensors mutated in place (deep-learning).  The drop-ins accept both; NumPy input is copied to
`default_device()` and results are copied back, CUDA tensors are used as they are.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib


def default_device() -> torch.device:
    """cuda:LOCAL_RANK under torchrun, else the current CUDA device.  Raises if there is no GPU or the
    extension is not built: rlvi_b200 has no CPU path."""
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("rlvi_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if "LOCAL_RANK" in os.environ:
        return torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    return torch.device("cuda", torch.cuda.current_device())


def as_device(a, like=None, dtype=torch.float64):
    """Returns (contiguous CUDA tensor of `dtype`, was_numpy)."""
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            dev = like.device if like is not None else default_device()
            return a.to(device=dev, dtype=dtype).contiguous(), False
        t = a if a.dtype == dtype else a.to(dtype)
        return t.contiguous(), False
    dev = like.device if like is not None else default_device()
    np_dtype = {torch.float64: np.float64, torch.float32: np.float32, torch.int64: np.int64}[dtype]
    arr = np.ascontiguousarray(a, dtype=np_dtype)
    return torch.from_numpy(arr).to(dev), True


def as_device_x(a, like=None):
    """`as_device` for a sample matrix: float32 input stays float32 (the FP32-stored mode: rlvi_loss_f32 /
    rlvi_weighted_moments_f32 on the TF32 tensor cores), everything else becomes float64 as in the reference."""
    is32 = (isinstance(a, torch.Tensor) and a.dtype == torch.float32) or \
           (isinstance(a, np.ndarray) and a.dtype == np.float32)
    return as_device(a, like=like, dtype=torch.float32 if is32 else torch.float64)


def to_caller(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t
