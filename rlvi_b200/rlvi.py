"""Drop-in for the reference's `standard-learning/rlvi.py`: same function names, positional order and
keyword defaults, running on librlvi_b200.so.

    update_weights(losses, tol=1e-3, maxiter=100)                       rlvi.py:8-20
    update_weights_constrained(losses, n_eff, tol=1e-3, maxiter=100)    rlvi.py:23-43
    mean(sample, maxiter=100, tol=1e-3)                                 rlvi.py:46-65
    linear_regression(X, y, maxiter=100, tol=1e-3)                      rlvi.py:68-89
    logistic_regression(X, y, maxiter=100, tol=1e-2)                    rlvi.py:92-108
    pca(sample, maxiter=100, tol=1e-2, theta_init=None)                 rlvi.py:111-125
    covariance(sample, eps, maxiter=100, tol=1e-2)                      rlvi.py:128-144

Array arguments may be NumPy arrays (copied to the GPU, NumPy returned -- the reference's calling
convention) or CUDA FP64 `torch.Tensor`s (used in place, tensors returned, nothing leaves the device
except the d-sized quantities the outer stop tests need).  All N-sized work -- per-sample losses, the
epsilon fixed point, the pi-weighted statistics -- runs in the CUDA kernels; the host (or torch on the
device) only does the O(d^3) algebra: the d x d solve / eigendecomposition and the stop tests.
There is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy import optimize as _opt

from . import ops
from . import utils as _utils
from ._host import as_device, as_device_x, to_caller

__all__ = ["update_weights", "update_weights_constrained", "mean", "linear_regression", "logistic_regression",
           "pca", "covariance"]


# --------------------------------------------------------------------------------------------------
# E-step
# --------------------------------------------------------------------------------------------------
def update_weights(losses, tol=1e-3, maxiter=100):
    """rlvi.py:8-20 -- the epsilon fixed point; one persistent kernel, loop exit decided on the device."""
    l, was_np = as_device(losses)
    pi, _ = ops.fixed_point(l, variant=ops.FP_STANDARD, tol=tol, maxiter=maxiter)
    return to_caller(pi, was_np)


def update_weights_constrained(losses, n_eff, tol=1e-3, maxiter=100):
    """rlvi.py:23-43 -- the fixed point, then (if sum pi < n_eff) the KKT shift found by SciPy's
    unbounded Brent exactly as in the reference; only the objective's O(N) sum is a device call
    (SURVEY.md H4)."""
    l, was_np = as_device(losses)
    n = l.numel()
    e = torch.empty_like(l)
    pi, res = ops.fixed_point(l, e_work=e, variant=ops.FP_STANDARD, tol=tol, maxiter=maxiter)
    sum_pi = ops.read_result(res)["sum_pi"]                      # rlvi.py:32  np.sum(weights) < n_eff
    if sum_pi < n_eff:
        c = (n - n_eff) / n_eff
        acc = torch.empty(1, dtype=torch.float64, device=l.device)

        def shift_sum(s, pi_out=None):
            # exp(-l + s) = e * exp(s): the fixed point already left e = exp(-l) behind, so an evaluation is one
            # exp-free pass; outside exp's comfortable range fall back to the literal expression
            if -600.0 < s < 600.0:
                ops.shift_sum_e(e, math.exp(s), c, pi_out=pi_out, out=acc)
            else:
                ops.shift_sum(l, s, c, pi_out=pi_out, out=acc)
            return acc.item()

        def shift_obj(s):                                        # rlvi.py:34-39
            return np.square(shift_sum(s) - n_eff)

        shift = _opt.minimize_scalar(shift_obj)["x"]             # rlvi.py:41
        # rlvi.py:42 -- the returned weights use the literal exp(-l + s): where e = exp(-l) has underflowed (l > ~708,
        # far outliers of the Gaussian NLL) e * exp(s) would be 0 or imprecise although exp(-l + s) is representable
        ops.shift_sum(l, shift, c, pi_out=pi, out=acc)
    return to_caller(pi, was_np)


# --------------------------------------------------------------------------------------------------
# helpers shared by the outer loops
# --------------------------------------------------------------------------------------------------
def _rel_change(new, old):
    """||new - old|| / ||old|| as a DEVICE scalar (d-sized work)."""
    return torch.linalg.norm(new - old) / torch.linalg.norm(old)


def _em_loop(theta, em_step, maxiter, tol, block):
    """The outer loops of rlvi.py:52-64,75-87,98-107,115-123: `theta <- em_step(theta)` until the relative change is
    <= tol.  The stop test is evaluated on the device; the host reads `block` of them back at once and returns the
    FIRST iterate that met it (small problems are sync-bound: a few iterations are issued ahead and the surplus is
    discarded; large ones use block = 1), so the result is the reference's."""
    it = 0
    while it < maxiter:
        k = min(block, maxiter - it)
        thetas, rels = [], []
        for _ in range(k):
            prev = theta
            theta = em_step(prev)
            thetas.append(theta)
            rels.append(_rel_change(theta, prev))
        host = torch.stack(rels).cpu()                           # one read-back per block
        for j in range(k):
            if float(host[j]) <= tol:
                return thetas[j]
        it += k
    return theta


def _estep_scaled(r2, wsum, e_work, pi):
    """losses = 0.5 * r2 / sigma2, sigma2 = (pi . r2) / sum(pi) (rlvi.py:50-51,58-59,73-74,82-83) folded
    into the fixed-point kernel as a device scalar: no host round trip between loss pass and E-step."""
    scale = 0.5 * wsum[1:2] / wsum[0:1]                          # 0.5 / sigma2, stays on the device
    ops.fixed_point(r2, scale=scale, e_work=e_work, out=pi, variant=ops.FP_STANDARD)
    return pi


# --------------------------------------------------------------------------------------------------
# outer EM loops
# --------------------------------------------------------------------------------------------------
def mean(sample, maxiter=100, tol=1e-3):
    """rlvi.py:46-65."""
    X, was_np = as_device(sample)
    n, d = X.shape
    pi = torch.ones(n, dtype=torch.float64, device=X.device)
    e_work = torch.empty_like(pi)
    r2 = torch.empty_like(pi)
    mom = None

    def mstep():
        nonlocal mom
        mom = ops.weighted_moments(X, pi, want_gram=False, out=mom)
        m = ops.split_moments(mom, d)
        theta = m["S1"] / m["S0"]                                # rlvi.py:48,56
        _, _, wsum = ops.loss(ops.LOSS_SQDIST, X, theta, weights=pi, losses_out=r2)   # rlvi.py:49-50,57-58
        return theta, wsum

    theta, wsum = mstep()

    def em_step(_prev):
        nonlocal wsum
        _estep_scaled(r2, wsum, e_work, pi)                      # rlvi.py:54
        th, wsum = mstep()
        return th

    theta = _em_loop(theta, em_step, maxiter, tol, _utils._spec_block(n, d))    # rlvi.py:62-64
    return to_caller(theta, was_np)


def _gpu_sym_solve(G, b, n):
    """theta = argmin ||sqrt(Pi) (X theta - y)||.  The reference calls scipy lstsq (LAPACK gelsd) on the
    N x d scaled design (rlvi.py:71,80); this is the same minimiser from the d x d statistics
    G = X^T Pi X, b = X^T Pi y: Cholesky when G is numerically positive definite, else an eigen-decomposition
    pseudo-inverse so that rank-deficient designs give the minimum-norm solution as gelsd does.  A summed Gram
    matrix carries rounding noise of about eps * (d + sqrt(n)) * lambda_max in its null directions, so that
    (not gelsd's eps * sigma_max on the design itself) is the rank threshold: pivots / eigenvalues below it
    are noise, not signal."""
    d = G.shape[0]
    rtol = torch.finfo(G.dtype).eps * (d + math.sqrt(n))
    L, info = torch.linalg.cholesky_ex(G)
    # pivot_i / G_ii = the share of feature i not explained by features < i: scale-free, so badly scaled
    # but independent features still take the Cholesky path
    resid = torch.diagonal(L) ** 2 / torch.diagonal(G)
    if int(info) == 0 and bool(resid.min() > rtol):          # the usual case
        return torch.cholesky_solve(b.unsqueeze(1), L)[:, 0]
    evals, evecs = torch.linalg.eigh(G)      # rank-deficient design: minimum-norm solution
    cutoff = rtol * evals.abs().max()
    inv = torch.where(evals > cutoff, 1.0 / evals, torch.zeros_like(evals))
    return evecs @ (inv * (evecs.T @ b))


def linear_regression(X, y, maxiter=100, tol=1e-3):
    """rlvi.py:68-89."""
    Xd, was_np = as_device(X)
    yd, _ = as_device(y, like=Xd)
    n, d = Xd.shape
    pi = torch.ones(n, dtype=torch.float64, device=Xd.device)
    e_work = torch.empty_like(pi)
    r2 = torch.empty_like(pi)
    mom = None

    def mstep():
        nonlocal mom
        mom = ops.weighted_moments(Xd, pi, y=yd, out=mom)
        m = ops.split_moments(mom, d)
        theta = _gpu_sym_solve(m["G"], m["Sy"], n)                  # rlvi.py:70-71,79-80
        _, _, wsum = ops.loss(ops.LOSS_SQRES, Xd, theta, y=yd, weights=pi, losses_out=r2)   # rlvi.py:72-73
        return theta, wsum

    theta, wsum = mstep()

    def em_step(_prev):
        nonlocal wsum
        _estep_scaled(r2, wsum, e_work, pi)                      # rlvi.py:77
        th, wsum = mstep()
        return th

    theta = _em_loop(theta, em_step, maxiter, tol, _utils._spec_block(n, d))    # rlvi.py:85-87
    return to_caller(theta, was_np)


def logistic_regression(X, y, maxiter=100, tol=1e-2, mstep="sklearn"):
    """rlvi.py:92-108.  `mstep="sklearn"` (reference default, line 96/103: L2-regularised fit with C = 100
    and the label-independent softplus loss of utils.sklearn_log_reg) or `mstep="mm"` (the alternative the
    reference keeps commented at lines 95/102: utils.mm_log_reg with the true cross-entropy)."""
    Xd, was_np = as_device(X)
    yd, _ = as_device(y, like=Xd)
    if mstep == "sklearn":        # warm start from the previous EM iterate: same (unique) minimiser, fewer Newton steps
        def fit(X_, y_, w_, th0=None):
            return _utils.sklearn_log_reg(X_, y_, w_, theta0=th0)
    else:
        def fit(X_, y_, w_, th0=None):
            return _utils.mm_log_reg(X_, y_, w_)
    n = Xd.shape[0]
    pi = torch.ones(n, dtype=torch.float64, device=Xd.device)
    theta, losses = fit(Xd, yd, pi)
    e_work = torch.empty_like(pi)
    for _ in range(maxiter):
        ops.fixed_point(losses, e_work=e_work, out=pi, variant=ops.FP_STANDARD)   # rlvi.py:100
        prev = theta
        theta, losses = fit(Xd, yd, pi, prev)
        if float(_rel_change(theta, prev)) <= tol:               # rlvi.py:105-107
            break
    return to_caller(theta, was_np)


def pca(sample, maxiter=100, tol=1e-2, theta_init=None):
    """rlvi.py:111-125.  A float32 `sample` runs in the FP32-stored mode (utils.pca)."""
    X, was_np = as_device_x(sample)
    n = X.shape[0]
    pi = torch.ones(n, dtype=torch.float64, device=X.device)
    t0 = None if theta_init is None else as_device(theta_init, like=X)[0]
    theta, losses = _utils.pca(X, pi, t0)
    e_work = torch.empty_like(pi)

    def em_step(_prev):
        nonlocal losses
        ops.fixed_point(losses, e_work=e_work, out=pi, variant=ops.FP_STANDARD)   # rlvi.py:117
        th, losses = _utils.pca(X, pi)
        return th

    theta = _em_loop(theta, em_step, maxiter, tol, _utils._spec_block(n, X.shape[1]))   # rlvi.py:121-123
    return to_caller(theta, was_np)


def covariance(sample, eps, maxiter=100, tol=1e-2):
    """rlvi.py:128-144."""
    X, was_np = as_device(sample)
    n = X.shape[0]
    n_eff = n * (1 - eps)                                        # rlvi.py:130
    pi = torch.ones(n, dtype=torch.float64, device=X.device)
    cov, losses = _utils.covariance(X, pi)
    for _ in range(maxiter):
        pi = update_weights_constrained(losses, n_eff)           # rlvi.py:135
        prev = cov
        cov, losses = _utils.covariance(X, pi)
        if float(torch.linalg.norm(cov - prev) / torch.linalg.norm(prev)) <= tol:   # rlvi.py:140 (Frobenius)
            break
    return to_caller(cov, was_np)
