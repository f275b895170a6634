"""Drop-in for the competitor weight rule of `standard-learning/rrm.py` (Robust Risk Minimization, Osama et al. 2020)
on the same kernels as the RLVI path:

    update_weights(losses, eps)                                  rrm.py:12-33  (= online-learning/main.py:61-81)
    mean(sample, eps, maxiter=100, tol=1e-3)                     rrm.py:36-52
    linear_regression(X, y, eps, maxiter=100, tol=1e-3)          rrm.py:55-75

The weight rule is one reduction per objective evaluation, sum_i max(exp(-l_i / alpha), 1e-16), inside SciPy's unbounded
Brent search over xi = log(alpha) -- the same shape as the constrained E-step of rlvi.py:23-43 -- so only that sum is a
device call (rlvi_rrm_sum_f64); the M-steps reuse rlvi_weighted_moments_f64 / rlvi_loss_f64.  NumPy in -> NumPy out,
CUDA tensors in -> CUDA tensors out.  No CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy import optimize as _opt

from . import ops
from ._host import as_device, to_caller
from .rlvi import _gpu_sym_solve

__all__ = ["update_weights", "mean", "linear_regression"]

_CUTOFF = 1e-16                                                  # rrm.py:16 numeric_cutoff


def _update_weights_device(l, eps, w_out=None):
    n = l.numel()
    t = -math.log((1 - eps) * n)                                 # rrm.py:15
    acc = torch.empty(1, dtype=torch.float64, device=l.device)

    def sum_phi(inv_alpha, norm=0.0, out=None):
        ops.rrm_sum(l, inv_alpha, _CUTOFF, norm=norm, w_out=out, out=acc)
        return acc.item()

    def objective(xi):                                           # rrm.py:17-21
        with np.errstate(over="ignore"):
            return np.exp(xi) * (np.log(sum_phi(float(np.exp(-xi)))) + t)

    opt_alpha = np.exp(_opt.minimize_scalar(objective)["x"])     # rrm.py:23-25
    inv_alpha = 1.0 / opt_alpha
    s = sum_phi(inv_alpha)                                       # rrm.py:27-30
    beta_over_alpha = np.log(s) - 1
    if w_out is None:
        w_out = torch.empty_like(l)
    # rrm.py:32 multiplies by exp(-x/alpha) with the literal division; x * (1/alpha) differs by one rounding of the argument
    sum_phi(inv_alpha, norm=float(np.exp(-beta_over_alpha - 1)), out=w_out)
    return w_out


def update_weights(losses, eps):
    """rrm.py:12-33 -- RRM sample weights: exp(-l / alpha*) / sum_phi(alpha*), alpha* from Brent on the dual."""
    l, was_np = as_device(losses)
    return to_caller(_update_weights_device(l, eps), was_np)


def mean(sample, eps, maxiter=100, tol=1e-3):
    """rrm.py:36-52."""
    X, was_np = as_device(sample)
    n, d = X.shape
    w = torch.full((n,), 1.0 / n, dtype=torch.float64, device=X.device)
    losses = torch.empty(n, dtype=torch.float64, device=X.device)
    mom = None

    def mstep():
        nonlocal mom
        mom = ops.weighted_moments(X, w, want_gram=False, out=mom)
        m = ops.split_moments(mom, d)
        theta = m["S1"] / m["S0"]
        ops.loss(ops.LOSS_SQDIST, X, theta, losses_out=losses)
        return theta

    theta = mstep()
    for _ in range(maxiter):
        _update_weights_device(losses, eps, w_out=w)
        prev = theta
        theta = mstep()
        if float(torch.linalg.norm(theta - prev) / torch.linalg.norm(prev)) <= tol:
            break
    return to_caller(theta, was_np)


def linear_regression(X, y, eps, maxiter=100, tol=1e-3):
    """rrm.py:55-75."""
    Xd, was_np = as_device(X)
    yd, _ = as_device(y, like=Xd)
    n, d = Xd.shape
    w = torch.full((n,), 1.0 / n, dtype=torch.float64, device=Xd.device)
    losses = torch.empty(n, dtype=torch.float64, device=Xd.device)
    mom = None

    def mstep():
        nonlocal mom
        mom = ops.weighted_moments(Xd, w, y=yd, out=mom)
        m = ops.split_moments(mom, d)
        theta = _gpu_sym_solve(m["G"], m["Sy"], n)
        ops.loss(ops.LOSS_SQRES, Xd, theta, y=yd, losses_out=losses)
        return theta

    theta = mstep()
    for _ in range(maxiter):
        _update_weights_device(losses, eps, w_out=w)
        prev = theta
        theta = mstep()
        if float(torch.linalg.norm(theta - prev) / torch.linalg.norm(prev)) <= tol:
            break
    return to_caller(theta, was_np)
