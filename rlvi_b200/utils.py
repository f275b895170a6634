"""Drop-in for the reference's `standard-learning/utils.py` (per-sample losses + weighted M-steps):

    sigmoid(x)                                      utils.py:7-16
    cross_entropy(X, theta, y)                      utils.py:19-21
    clf_predict(X, theta, augment=True)             utils.py:24-29
    mm_log_reg(X, y, weights)                       utils.py:32-58
    sklearn_log_reg(X, y, weights, reg_coeff=1e2)   utils.py:61-73
    pca(samples, weights, theta=None)               utils.py:76-89
    covariance(samples, weights, mean=None)         utils.py:92-108

Same signatures and return values `(theta, losses)`.  NumPy in -> NumPy out, CUDA tensors in -> CUDA
tensors out.  Every pass over the N samples is a CUDA kernel of librlvi_b200.so; the d x d algebra
(inverse, eigen-decomposition, Cholesky) is done by torch on the device.  No CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from ._host import as_device, as_device_x, to_caller

__all__ = ["sigmoid", "cross_entropy", "clf_predict", "mm_log_reg", "sklearn_log_reg", "pca", "covariance"]


def sigmoid(x):
    """utils.py:7-16 -- overflow-free logistic function, any shape: one CUDA kernel (rlvi_sigmoid_f64) whatever the size
    (the N-sized sigmoid of the M-step itself lives inside rlvi_logistic_grad_f64)."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        t = x.to(torch.float64).contiguous()
        return ops.sigmoid(t).view(x.shape)
    scalar = np.ndim(x) == 0
    t, _ = as_device(np.atleast_1d(np.asarray(x, dtype=np.float64)))
    r = ops.sigmoid(t).cpu().numpy().reshape(np.shape(np.atleast_1d(x)))
    return float(r[0]) if scalar else r.reshape(np.shape(x))


def cross_entropy(X, theta, y):
    """utils.py:19-21 -- l = -y phi + phi + log1p(exp(-phi)), phi = X theta.  As in the reference, X is
    the design matrix theta applies to (mm_log_reg passes the ones-augmented matrix)."""
    Xd, was_np = as_device(X)
    td, _ = as_device(theta, like=Xd)
    yd, _ = as_device(y, like=Xd)
    losses, _, _ = ops.loss(ops.LOSS_LOGISTIC_CE, Xd, td, y=yd, intercept=False)
    return to_caller(losses, was_np)


def clf_predict(X, theta, augment=True):
    """utils.py:24-29 -- 0/1 predictions (int): sigmoid([1, X] theta) > 0.5, or sigmoid(X theta) > 0.5 with
    `augment=False`."""
    Xd, was_np = as_device(X)
    td, _ = as_device(theta, like=Xd)
    # softplus(phi) > log 2  <=>  phi > 0  <=>  sigmoid(phi) > 0.5
    sp, _, _ = ops.loss(ops.LOSS_SOFTPLUS, Xd, td, intercept=bool(augment))
    pred = (sp > math.log(2.0)).to(torch.int64)
    return to_caller(pred, was_np)


def _augmented_gram(X, y, weights, mom=None):
    """[[S0, S1^T], [S1, G]] = [1, X]^T diag(w) [1, X] from one statistics pass."""
    d = X.shape[1]
    mom = ops.weighted_moments(X, weights, y=None, out=mom)
    m = ops.split_moments(mom, d)
    A = torch.empty((d + 1, d + 1), dtype=torch.float64, device=X.device)
    A[0, 0] = m["S0"]
    A[0, 1:] = m["S1"]
    A[1:, 0] = m["S1"]
    A[1:, 1:] = m["G"]
    return A


def _spec_block(n, d):
    """How many iterations of a host-tested loop to issue before reading the stop tests back: small problems are
    launch / sync bound (run a few iterations ahead, discard the surplus), large ones are not (test every one)."""
    return 1 if n * d >= (1 << 22) else 4


def mm_log_reg(X, y, weights):
    """utils.py:32-58 -- MM logistic regression: Q = 1/4 [1,X]^T Pi [1,X] once (one statistics pass),
    then theta <- theta - Q^-1 [1,X]^T (pi * (sigmoid - y)) (one gradient pass each) until
    ||delta theta|| <= 1e-2; returns (theta [d+1] intercept first, cross-entropy losses).
    The stop test is evaluated on the device; the host reads the results back once per `_spec_block` steps and
    returns the FIRST iterate that met it, so the trajectory and the result are the reference's."""
    Xd, was_np = as_device(X)
    yd, _ = as_device(y, like=Xd)
    wd, _ = as_device(weights, like=Xd)
    n, d = Xd.shape
    Q_inv = torch.linalg.inv(0.25 * _augmented_gram(Xd, yd, wd))         # utils.py:36-38
    theta0 = torch.zeros(d + 1, dtype=torch.float64, device=Xd.device)   # utils.py:35
    g = torch.empty(d + 1, dtype=torch.float64, device=Xd.device)

    def step(theta):                                                      # utils.py:40-41,45,50
        ops.logistic_grad(Xd, yd, wd, theta, out=g)
        return theta - Q_inv @ g

    block = _spec_block(n, d)
    theta1 = step(theta0)
    n_grad = 1
    done = False
    while not done:
        thetas, norms = [theta1], [torch.linalg.norm(theta1 - theta0)]   # utils.py:47
        for _ in range(block - 1):
            thetas.append(step(thetas[-1]))
            norms.append(torch.linalg.norm(thetas[-1] - thetas[-2]))
        host = torch.stack(norms).cpu()                                   # one read-back per block
        for j in range(block):
            if not (float(host[j]) > 1e-2):                               # while-condition false: this iterate is returned
                theta1, done = thetas[j], True
                n_grad += j
                break
        if not done:
            theta0 = thetas[-1]
            theta1 = step(theta0)
            n_grad += block
    mm_log_reg.last_n_grad = n_grad
    losses, _, _ = ops.loss(ops.LOSS_LOGISTIC_CE, Xd, theta1, y=yd, intercept=True)   # utils.py:56
    return to_caller(theta1, was_np), to_caller(losses, was_np)


def sklearn_log_reg(X, y, weights, reg_coeff=1e2, xtol=1e-9, max_newton=60, theta0=None):
    """utils.py:61-73.  Reproduces the function's observable behaviour:
      * `weights` is normalised IN PLACE by its maximum (line 66, quirk Q3);
      * theta = [intercept, coef] minimises liblinear's objective for `LogisticRegression(C=reg_coeff,
        solver="liblinear")` with sample weights:  F = 1/2 ||[b, w]||^2 + C sum_i pi_i logloss_i
        (liblinear regularises the intercept too: intercept_scaling = 1);
      * losses = -log P(class 0 | x) = softplus(b + x.w) for every row whatever its label (lines 62-64).
    liblinear itself (third-party C++) is not re-implemented: the same strictly convex
    problem is solved on the device by damped NEWTON steps (liblinear's own method is a trust-region Newton).  One
    iteration = three passes over X: the cross-entropy pass (objective value + exp(-loss) for the curvature),
    the gradient pass (rlvi_logistic_grad_f64) and the statistics pass with the IRLS weights pi s (1 - s)
    (rlvi_irls_weights_f64 -> rlvi_weighted_moments_f64 = the exact Hessian).  Armijo back-tracking on F makes it
    globally convergent (nearly separable data, C = 100); it stops when a full Newton step no longer moves theta
    (||step|| <= xtol (1 + ||theta||): quadratic convergence puts the remaining error at ~xtol^2), far inside
    liblinear's own tolerance (1e-4).  `theta0` warm-starts the solve (rlvi.logistic_regression passes the previous EM
    iterate); the minimiser is unique, so the result does not depend on it.  Raises if `max_newton` is exhausted."""
    Xd, was_np = as_device(X)
    yd, _ = as_device(y, like=Xd)
    if isinstance(weights, np.ndarray):
        weights /= np.max(weights)                                        # utils.py:66, caller's array
        wd, _ = as_device(weights, like=Xd)
    else:
        wd, _ = as_device(weights, like=Xd)
        wd /= wd.max()
    n, d = Xd.shape
    dev = Xd.device
    C = float(reg_coeff)
    eye = torch.eye(d + 1, dtype=torch.float64, device=dev)
    theta = torch.zeros(d + 1, dtype=torch.float64, device=dev) if theta0 is None else as_device(theta0, like=Xd)[0].clone()
    e = torch.empty(n, dtype=torch.float64, device=dev)
    h = torch.empty(n, dtype=torch.float64, device=dev)
    g = torch.empty(d + 1, dtype=torch.float64, device=dev)
    ws = torch.empty(2, dtype=torch.float64, device=dev)
    mom = None
    passes = 0

    def objective(th):              # F(th); leaves e_i = exp(-logloss_i(th)) behind
        nonlocal passes
        ops.loss(ops.LOSS_LOGISTIC_CE, Xd, th, y=yd, intercept=True, weights=wd, want_losses=False, e_out=e, wsum_out=ws)
        passes += 1
        return float(0.5 * (th @ th) + C * ws[0])

    F = objective(theta)
    converged = False
    for _ in range(max_newton):
        ops.logistic_grad(Xd, yd, wd, theta, out=g)
        grad = theta + C * g
        ops.irls_weights(e, wd, out=h)
        mom = ops.weighted_moments(Xd, h, out=mom)
        passes += 2
        m = ops.split_moments(mom, d)
        A = torch.empty((d + 1, d + 1), dtype=torch.float64, device=dev)
        A[0, 0] = m["S0"]
        A[0, 1:] = m["S1"]
        A[1:, 0] = m["S1"]
        A[1:, 1:] = m["G"]
        step = torch.linalg.solve(eye + C * A, grad)
        slope = float(grad @ step)
        small = float(torch.linalg.norm(step)) <= xtol * (1.0 + float(torch.linalg.norm(theta)))
        t = 1.0
        while True:
            cand = theta - t * step
            Fc = objective(cand)
            if Fc <= F - 1e-4 * t * slope:              # Armijo
                break
            if small and Fc <= F + 1e-12 * abs(F):      # at the rounding floor of F: accept the (tiny) full step
                break
            if t < 1e-8:
                break
            t *= 0.5
        theta, F = cand, Fc
        if small and t == 1.0:
            converged = True
            break
    sklearn_log_reg.last_passes = passes
    if not converged and np.isfinite(F):
        raise RuntimeError(f"sklearn_log_reg: Newton iteration did not converge in {max_newton} steps")
    losses, _, _ = ops.loss(ops.LOSS_SOFTPLUS, Xd, theta, intercept=True)   # utils.py:70-71
    return to_caller(theta, was_np), to_caller(losses, was_np)


def _svd_flip_unit(v):
    """sklearn's v-based svd_flip (largest-|.| entry positive) + utils.py:85 normalisation."""
    k = torch.argmax(torch.abs(v))
    v = v * torch.sign(v[k])
    return v / torch.linalg.norm(v)


def pca(samples, weights, theta=None, precision=ops.TF32X3):
    """utils.py:76-89.  theta = top principal direction of the rows pi_i x_i after column-centring
    (what `PCA(n_components=1).fit(diag(pi) @ X)` returns: quirk Q4 -- weights pi, not sqrt(pi)), from
    G = sum pi_i^2 x_i x_i^T and m = sum pi_i x_i / N:  C = (G - N m m^T)/(N-1), top eigenvector, sklearn's
    sign rule, unit norm.  losses_i = ||x_i||^2 - (x_i . theta)^2 (uncentred).
    float32 `samples` stay float32 on the device (FP32-stored mode): the Gram runs on the TF32 tensor cores
    (`precision`), losses / weights / theta stay FP64."""
    X, was_np = as_device_x(samples)
    n, d = X.shape
    if theta is None:
        wd, _ = as_device(weights, like=X)
        mom = ops.weighted_moments(X, wd, power=2, precision=precision)
        m = ops.split_moments(mom, d)
        mu = m["S1"] / n
        Cm = (m["G"] - n * torch.outer(mu, mu)) / (n - 1)
        _, evecs = torch.linalg.eigh(Cm)
        theta_d = _svd_flip_unit(evecs[:, -1])
    else:
        theta_d, _ = as_device(theta, like=X)
    losses, _, _ = ops.loss(ops.LOSS_PCA, X, theta_d)
    return to_caller(theta_d, was_np), to_caller(losses, was_np)


def _gaussian_params(mu, cov):
    """[c, mu, U] for RLVI_LOSS_GAUSSIAN: U upper triangular with U^T U = cov^-1, c = log|cov| + d log 2pi.
    Raises ValueError("Singular covariance matrix") like utils.py:98-100."""
    d = mu.numel()
    sign, logabsdet = torch.linalg.slogdet(cov)
    if not (float(sign) > 0):
        raise ValueError("Singular covariance matrix")
    # cov = R R^T with R UPPER triangular (Cholesky of the index-reversed matrix), U = R^-1
    J = torch.arange(d - 1, -1, -1, device=cov.device)
    Lr, info = torch.linalg.cholesky_ex(cov[J][:, J])
    if int(info) != 0:
        raise ValueError("Singular covariance matrix")
    R = Lr[J][:, J]
    U = torch.linalg.solve_triangular(R, torch.eye(d, dtype=cov.dtype, device=cov.device), upper=True)
    c = logabsdet + d * math.log(2 * math.pi)
    return torch.cat([c.reshape(1), mu, torch.triu(U).reshape(-1)]).contiguous()


def covariance(samples, weights, mean=None):
    """utils.py:92-108.  mu = X^T pi / sum pi; cov = (X-mu)^T Pi (X-mu) / sum pi (a mean pass, then a centred
    statistics pass, as the reference centres before the contraction); losses = Gaussian NLL.  The `mean`
    argument is ignored, as in the reference (line 103)."""
    X, was_np = as_device(samples)
    wd, _ = as_device(weights, like=X)
    n, d = X.shape
    # pass 1: the weighted mean (utils.py:103); pass 2: statistics of the samples centred at it (utils.py:104-105)
    m1 = ops.split_moments(ops.weighted_moments(X, wd, want_gram=False), d)
    mu = (m1["S1"] / m1["S0"]).contiguous()
    m = ops.split_moments(ops.weighted_moments(X, wd, center=mu), d)
    delta = m["S1"] / m["S0"]                     # rounding-level residual of the centring
    cov = m["G"] / m["S0"] - torch.outer(delta, delta)
    cov = 0.5 * (cov + cov.T)
    params = _gaussian_params(mu, cov)
    losses, _, _ = ops.loss(ops.LOSS_GAUSSIAN, X, params)
    return to_caller(cov, was_np), to_caller(losses, was_np)
