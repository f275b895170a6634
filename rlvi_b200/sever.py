"""Drop-in for the competitor filter of `standard-learning/sever.py` (SEVER, Diakonikolas et al. 2019) on the same kernels as
the RLVI path ("SEVER's per-sample gradient scores"):

    linear_regression(X, y, eps, numiter=4)                  sever.py:11-42
    pca(samples, eps, numiter=4, theta_init=None)            sever.py:82-113

One filter step = base learner on the active set, per-sample gradients g_i = c_i x_i, their centred top right singular
vector v, scores tau_i = ((g_i - mean g) . v)^2, drop the int(eps / 2 * n_active) largest.  The reference materialises the
n x d gradient matrix and runs an SVD of it; here the active set is a 0/1 weight vector and a step is four passes over X:
the base learner's statistics (rlvi_weighted_moments_f64), the coefficients c_i (rlvi_sever_pass_f64 op 0), the d x d
scatter of the gradients (one more statistics pass: sum_active g g^T and sum_active g at once) whose top eigenvector IS v,
and the scores (op 1).  Only the d x d algebra and the top-p selection (torch.topk) are not library kernels of this repo.
NumPy in -> NumPy out, CUDA tensors in -> CUDA tensors out.  No CPU fallback.

DEVIATION (quirk Q12): sever.py:26-27 reads `V = np.linalg.svd(G_cen)[-1]; v = V[:, 0]`.  NumPy returns
V^H, so that is the vector of first components of ALL right singular vectors, each with whatever sign LAPACK gave it --
not the "top right singular vector" of the comment, and not reproducible without LAPACK's SVD of the n x d gradient
matrix.  This module implements what the comment (and the SEVER paper) say, v = V^H[0, :]; it matches
`oracle.rlvi_np.sever_*(as_written=False)` to 1e-9, while the oracle's literal restatement is pinned on the reference.
"""
from __future__ import annotations

import torch

from . import ops
from ._host import as_device, to_caller
from .rlvi import _gpu_sym_solve
from .utils import _svd_flip_unit

__all__ = ["linear_regression", "pca"]


def _filter_step(X, active, n_active, c_args, eps, bufs):
    """sever.py:21-40 / :94-111 after the base learner: returns the new active-set size."""
    n, d = X.shape
    theta, alpha, b = c_args
    c, q, u = ops.sever_pass(X, theta, 0, alpha, b=b, active=active, out0=bufs[0], out1=bufs[1], out2=bufs[2])
    m = ops.split_moments(ops.weighted_moments(X, q, y=u), d)     # G = sum_active g g^T, Sy = sum_active g
    gbar = m["Sy"] / n_active                                     # np.mean(G_uncen, axis=0)
    scatter = m["G"] - n_active * torch.outer(gbar, gbar)         # G_cen^T G_cen
    _, evecs = torch.linalg.eigh(scatter)
    v = evecs[:, -1].contiguous()                                 # top right singular vector of G_cen (sign is irrelevant)
    tau, _, _ = ops.sever_pass(X, v, 1, float(gbar @ v), a=c, active=active, out0=bufs[3])
    p = int(eps / 2 * n_active)                                   # sever.py:34 / :107
    if p > 0:
        active[torch.topk(tau, p).indices] = 0.0                  # idx[p:] of argsort(-tau) stays
    return n_active - p


def linear_regression(X, y, eps, numiter=4):
    """sever.py:11-42.  theta of the LAST base fit is returned (the fit precedes the filter inside the loop)."""
    Xd, was_np = as_device(X)
    yd, _ = as_device(y, like=Xd)
    n, d = Xd.shape
    active = torch.ones(n, dtype=torch.float64, device=Xd.device)
    bufs = [torch.empty(n, dtype=torch.float64, device=Xd.device) for _ in range(4)]
    n_active = n
    theta = None
    for _ in range(numiter):
        m = ops.split_moments(ops.weighted_moments(Xd, active, y=yd), d)
        theta = _gpu_sym_solve(m["G"], m["Sy"], n_active)          # sever.py:20  lstsq on the active rows
        n_active = _filter_step(Xd, active, n_active, (theta, 2.0, yd), eps, bufs)   # g_i = 2 (x_i.theta - y_i) x_i
    return to_caller(theta, was_np)


def pca(samples, eps, numiter=4, theta_init=None):
    """sever.py:82-113.  Base learner = utils.pca on the active rows with unit weights (top principal direction of the
    column-centred active rows, sklearn's sign rule, unit norm)."""
    X, was_np = as_device(samples)
    n, d = X.shape
    active = torch.ones(n, dtype=torch.float64, device=X.device)
    bufs = [torch.empty(n, dtype=torch.float64, device=X.device) for _ in range(4)]
    n_active = n
    theta = None
    for k in range(numiter):
        if theta_init is not None and k == 0:
            theta = as_device(theta_init, like=X)[0].contiguous()
        else:
            m = ops.split_moments(ops.weighted_moments(X, active, power=2), d)
            mu = m["S1"] / n_active
            cov = (m["G"] - n_active * torch.outer(mu, mu)) / (n_active - 1)
            _, evecs = torch.linalg.eigh(cov)
            theta = _svd_flip_unit(evecs[:, -1]).contiguous()
        n_active = _filter_step(X, active, n_active, (theta, -2.0, None), eps, bufs)   # g_i = -2 (x_i.theta) x_i
    return to_caller(theta, was_np)
