"""NumPy restatement of the reference's standard-learning and online-learning RLVI path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are into /root/reference/.

The arithmetic deliberately keeps the reference's floating-point operation order wherever the result
feeds a stop test (the fixed point is cut off by `err < tol`, so parity is parity of the *iteration*),
while dropping the reference's N x N `np.diag` temporaries (`diag(w) @ A == w[:, None] * A` bit for
bit -- checked in tests/test_oracle_golden.py), so the oracle scales past N ~ 3e4.
"""
from __future__ import annotations

import numpy as np
from scipy import optimize as _opt
from scipy.linalg import lstsq as _lstsq


# --------------------------------------------------------------------------------------------------
# E-step: epsilon fixed point
# --------------------------------------------------------------------------------------------------
def fixed_point_trace(losses, tol=1e-3, maxiter=100):
    """standard-learning/rlvi.py:8-20 -- returns (pi, eps_last, n_iterations, err_last).

    pi^0 = 0.95; each pass: eps = 1 - mean(pi); rho = eps / (1 - eps); pi' = e / (rho + e) with
    e = exp(-loss); err = ||pi' - pi||_2; stop when err < tol.  The last computed pi' is returned
    (also when maxiter is exhausted).
    """
    losses = np.asarray(losses, dtype=np.float64)
    pi = np.full_like(losses, 0.95)
    new_pi = pi.copy()
    eps = np.nan
    err = np.nan
    k = 0
    for k in range(1, maxiter + 1):
        eps = 1 - np.mean(pi)
        rho = eps / (1 - eps)
        e = np.exp(-losses)
        new_pi = e / (rho + e)
        err = np.linalg.norm(new_pi - pi)
        pi = new_pi.copy()
        if err < tol:
            break
    return new_pi, eps, k, err


def update_weights(losses, tol=1e-3, maxiter=100):
    """standard-learning/rlvi.py:8-20."""
    return fixed_point_trace(losses, tol=tol, maxiter=maxiter)[0]


def shift_objective(losses, n_eff, s):
    """standard-learning/rlvi.py:34-39 -- (sum_i e^{-l_i+s} / (c + e^{-l_i+s}) - n_eff)^2, c = (n-n_eff)/n_eff."""
    n = len(losses)
    t = np.exp(-losses + s)
    return np.square(np.sum(t / ((n - n_eff) / n_eff + t)) - n_eff)


def update_weights_constrained(losses, n_eff, tol=1e-3, maxiter=100):
    """standard-learning/rlvi.py:23-43 -- a1, then the KKT shift when sum(pi) < n_eff (Brent, unbounded)."""
    losses = np.asarray(losses, dtype=np.float64)
    n = len(losses)
    pi = update_weights(losses, tol=tol, maxiter=maxiter)
    if np.sum(pi) < n_eff:
        shift = _opt.minimize_scalar(lambda s: shift_objective(losses, n_eff, s))["x"]
        t = np.exp(-losses + shift)
        pi = t / ((n - n_eff) / n_eff + t)
    return pi


def update_weights_online(losses, tol=1e-3, maxiter=100):
    """online-learning/main.py:45-58 -- pi^0 = 0.5, rho = avg/(1-avg), pi' = rho e / (1 + rho e);
    the break happens BEFORE pi is replaced; the result is normalised by max(pi') * n."""
    losses = np.asarray(losses, dtype=np.float64)
    e = np.exp(-losses)
    pi = np.full_like(losses, 0.5)
    new_pi = pi
    for _ in range(maxiter):
        avg = np.mean(pi)
        rho = avg / (1 - avg)
        new_pi = rho * e / (1 + rho * e)
        if np.linalg.norm(new_pi - pi) < tol:
            break
        pi = new_pi.copy()
    return new_pi / (np.max(new_pi) * len(new_pi))


def online_cross_entropy(log_proba, targets):
    """online-learning/main.py:84-85 -- algebraically -log_proba whatever the label (quirk Q11)."""
    return -targets * log_proba - (1 - targets) * log_proba


# --------------------------------------------------------------------------------------------------
# Per-sample losses (standard-learning/utils.py)
# --------------------------------------------------------------------------------------------------
def sigmoid(x):
    """standard-learning/utils.py:7-16 -- overflow-free logistic function."""
    x = np.asarray(x, dtype=np.float64)
    z = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0, z) / (1 + z)


def cross_entropy(X, theta, y):
    """standard-learning/utils.py:19-21 -- -y phi + phi + log1p(exp(-phi)), phi = X theta."""
    phi = X @ theta
    return -y * phi + phi + np.log1p(np.exp(-phi))


def softplus_loss(X, theta):
    """standard-learning/utils.py:62-64 -- the loss `sklearn_log_reg` reports: -log P(class 0 | x),
    label-independent (quirk Q3).  theta = [intercept, coef]; equals logaddexp(0, phi)."""
    phi = theta[0] + X @ theta[1:]
    return np.logaddexp(0.0, phi)


def pca_losses(samples, theta):
    """standard-learning/utils.py:77-79 -- ||x||^2 - (x . theta)^2 (uncentred)."""
    proj = samples @ theta
    return np.sum(samples ** 2, axis=1) - proj ** 2


def gaussian_losses(samples, mean, cov):
    """standard-learning/utils.py:93-101 -- 0.5 [(x-mu)^T cov^-1 (x-mu) + log|cov| + d log 2 pi]."""
    d = mean.shape[0]
    centered = samples - mean
    scaled = _lstsq(cov, centered.T)[0]
    quad = np.sum(centered * scaled.T, axis=1)
    sign, logabsdet = np.linalg.slogdet(cov)
    if sign <= 0:
        raise ValueError("Singular covariance matrix")
    return 0.5 * (quad + logabsdet + d * np.log(2 * np.pi))


# --------------------------------------------------------------------------------------------------
# Weighted sufficient statistics (what the M-steps need) -- the quantity the CUDA moments kernel emits
# --------------------------------------------------------------------------------------------------
def weighted_moments(X, w, y=None):
    """S0 = sum w, S1 = X^T w, Sy = X^T (w*y), G = X^T diag(w) X  (FP64).

    These are the contractions behind rlvi.py:48,56 (mean), rlvi.py:70-71,79-80 (normal equations of
    the sqrt(w)-scaled lstsq), utils.py:36-38 (MM majoriser Q), utils.py:103-105 (covariance) and,
    with w = pi^2, the Gram of the pi-scaled rows that utils.py:82-84 hands to sklearn PCA."""
    Xw = X * w[:, None]
    out = {"S0": np.sum(w), "S1": X.T @ w, "G": Xw.T @ X}
    if y is not None:
        out["Sy"] = Xw.T @ y
        out["Swy"] = w @ y
    return out


# --------------------------------------------------------------------------------------------------
# M-steps
# --------------------------------------------------------------------------------------------------
def mm_log_reg(X, y, weights):
    """standard-learning/utils.py:32-58 -- MM (quadratic majoriser) logistic regression.
    Returns (theta [d+1], intercept first; cross_entropy losses)."""
    Xa = np.hstack([np.ones((X.shape[0], 1)), X])
    theta0 = np.zeros(Xa.shape[1])
    Xt = 0.5 * Xa * np.sqrt(weights)[:, None]
    Q_inv = np.linalg.inv(Xt.T @ Xt)

    def step(theta):
        g = Xa.T @ (weights * (sigmoid(Xa @ theta) - y))
        return theta - Q_inv @ g

    theta1 = step(theta0)
    n_grad = 1
    while np.linalg.norm(theta1 - theta0) > 1e-2:
        theta0 = theta1
        theta1 = step(theta0)
        n_grad += 1
    mm_log_reg.last_n_grad = n_grad
    return theta1, cross_entropy(Xa, theta1, y)


def pca_direction(samples, weights):
    """standard-learning/utils.py:81-85 -- top principal direction of the rows pi_i * x_i after
    column-centring (what sklearn PCA(n_components=1).fit(diag(pi) @ X) returns), unit norm, with
    sklearn's sign rule (largest-|.| entry positive).  Restated through the covariance_eigh route
    sklearn >= 1.5 takes for N >= 10 d: C = (sum pi_i^2 x_i x_i^T - N m m^T)/(N-1), m = sum pi_i x_i / N."""
    n = samples.shape[0]
    Z = samples * weights[:, None]
    m = Z.mean(axis=0)
    C = (Z.T @ Z - n * np.outer(m, m)) / (n - 1)
    evals, evecs = np.linalg.eigh(C)
    v = evecs[:, -1]
    v = v * np.sign(v[np.argmax(np.abs(v))])
    return v / np.linalg.norm(v)


def pca_mstep(samples, weights, theta=None):
    """standard-learning/utils.py:76-89."""
    if theta is None:
        theta = pca_direction(samples, weights)
    return theta, pca_losses(samples, theta)


def covariance_mstep(samples, weights):
    """standard-learning/utils.py:92-108 (the `mean` argument is ignored there, line 103)."""
    mean = samples.T @ weights / np.sum(weights)
    centered = samples - mean
    # (C^T diag(w)) is a C-contiguous d x N array in the reference; keep that layout so BLAS sums in the same order
    cov = np.ascontiguousarray((weights[:, None] * centered).T) @ centered / np.sum(weights)
    return cov, gaussian_losses(samples, mean, cov)


# --------------------------------------------------------------------------------------------------
# Outer EM loops (standard-learning/rlvi.py)
# --------------------------------------------------------------------------------------------------
def _rel_change(new, old):
    return np.linalg.norm(new - old) / np.linalg.norm(old)


def mean(sample, maxiter=100, tol=1e-3):
    """standard-learning/rlvi.py:46-65."""
    def mstep(w):
        theta = w @ sample / np.sum(w)
        r2 = np.linalg.norm(theta - sample, axis=1) ** 2
        sigma2 = w @ r2 / np.sum(w)
        return theta, 0.5 * r2 / sigma2

    theta, losses = mstep(np.ones(sample.shape[0]))
    for _ in range(maxiter):
        w = update_weights(losses)
        prev = theta.copy()
        theta, losses = mstep(w)
        if _rel_change(theta, prev) <= tol:
            break
    return theta


def linear_regression(X, y, maxiter=100, tol=1e-3, trace=None):
    """standard-learning/rlvi.py:68-89 with `diag(sqrt(w)) @ A` written as a row scaling."""
    def mstep(w):
        sw = np.sqrt(w)
        theta = _lstsq(sw[:, None] * X, sw * y)[0]
        r2 = (y - X @ theta) ** 2
        sigma2 = w @ r2 / np.sum(w)
        return theta, 0.5 * r2 / sigma2

    theta, losses = mstep(np.ones(X.shape[0]))
    for _ in range(maxiter):
        w = update_weights(losses)
        prev = theta.copy()
        theta, losses = mstep(w)
        if trace is not None:
            trace.append((theta.copy(), w.copy()))
        if _rel_change(theta, prev) <= tol:
            break
    return theta


def logistic_regression_mm(X, y, maxiter=100, tol=1e-2):
    """standard-learning/rlvi.py:92-108 with the M-step the reference keeps as the commented
    alternative (lines 95, 102): utils.mm_log_reg."""
    theta, losses = mm_log_reg(X, y, np.ones(X.shape[0]))
    for _ in range(maxiter):
        w = update_weights(losses)
        prev = theta.copy()
        theta, losses = mm_log_reg(X, y, w)
        if _rel_change(theta, prev) <= tol:
            break
    return theta


def pca(sample, maxiter=100, tol=1e-2, theta_init=None):
    """standard-learning/rlvi.py:111-125."""
    theta, losses = pca_mstep(sample, np.ones(sample.shape[0]), theta_init)
    for _ in range(maxiter):
        w = update_weights(losses)
        prev = theta.copy()
        theta, losses = pca_mstep(sample, w)
        if _rel_change(theta, prev) <= tol:
            break
    return theta


def covariance(sample, eps, maxiter=100, tol=1e-2):
    """standard-learning/rlvi.py:128-144."""
    n = sample.shape[0]
    n_eff = n * (1 - eps)
    cov, losses = covariance_mstep(sample, np.ones(n))
    for _ in range(maxiter):
        w = update_weights_constrained(losses, n_eff)
        prev = cov.copy()
        cov, losses = covariance_mstep(sample, w)
        if np.linalg.norm(cov - prev, ord="fro") / np.linalg.norm(prev, ord="fro") <= tol:
            break
    return cov


# --------------------------------------------------------------------------------------------------
# Competitor weight rule on the same path (standard-learning/rrm.py)
# --------------------------------------------------------------------------------------------------
def rrm_update_weights(losses, eps):
    """standard-learning/rrm.py:12-33 (= online-learning/main.py:61-81): Robust Risk Minimization weights."""
    res = np.copy(losses)
    t = -np.log((1 - eps) * res.shape[0])
    cutoff = 1e-16

    def objective(xi):
        phi = np.exp(-res * np.exp(-xi))
        phi[phi < cutoff] = cutoff
        return np.exp(xi) * (np.log(np.sum(phi)) + t)

    alpha = np.exp(_opt.minimize_scalar(objective)["x"])
    phi = np.exp(-res / alpha)
    phi[phi < cutoff] = cutoff
    beta_over_alpha = np.log(np.sum(phi)) - 1
    return np.exp(-res / alpha) * np.exp(-beta_over_alpha - 1)


def rrm_linear_regression(X, y, eps, maxiter=100, tol=1e-3):
    """standard-learning/rrm.py:55-75 with `diag(sqrt(w)) @ A` written as a row scaling."""
    def fit(w):
        sw = np.sqrt(w)
        theta = _lstsq(sw[:, None] * X, sw * y)[0]
        return theta, (y - X @ theta) ** 2

    theta, losses = fit(np.ones(X.shape[0]) / X.shape[0])
    for _ in range(maxiter):
        w = rrm_update_weights(losses, eps)
        prev = theta.copy()
        theta, losses = fit(w)
        if _rel_change(theta, prev) <= tol:
            break
    return theta


def sever_filter_scores(Xs, c, as_written=True):
    """standard-learning/sever.py:22-31 / :95-104: gradients g_i = c_i x_i, centred; "top right singular vector"; scores.

    `as_written=True` restates line 26-27 literally: `V = np.linalg.svd(G_cen)[-1]; v = V[:, 0]`.  NumPy returns V^H, whose
    ROWS are the right singular vectors, so `V[:, 0]` is the vector of FIRST COMPONENTS of all d singular vectors (quirk
    Q12), and it depends on the sign LAPACK happens to give each of them.  `as_written=False` takes the top right singular
    vector the comment (and the SEVER paper) mean, `V[0, :]`, which is defined up to a sign the squared score removes."""
    G_uncen = c[:, None] * Xs
    G_cen = G_uncen - np.mean(G_uncen, axis=0)
    Vh = np.linalg.svd(G_cen)[-1]
    v = Vh[:, 0] if as_written else Vh[0, :]
    return (G_cen @ v) ** 2


def sever_linear_regression(X, y, eps, numiter=4, as_written=True):
    """standard-learning/sever.py:11-42."""
    Xs, ys = X, y
    theta = None
    for _ in range(numiter):
        theta = _lstsq(Xs, ys)[0]
        tau = sever_filter_scores(Xs, 2 * (Xs @ theta - ys), as_written)
        p = int(eps / 2 * Xs.shape[0])
        s = np.argsort(-tau)[p:]
        Xs, ys = Xs[s, :], ys[s]
    return theta


def sever_pca(samples, eps, numiter=4, theta_init=None, as_written=True):
    """standard-learning/sever.py:82-113 (base learner utils.pca with unit weights = pca_direction)."""
    Xs = samples
    theta = None
    for k in range(numiter):
        theta = theta_init if (theta_init is not None and k == 0) else pca_direction(Xs, np.ones(len(Xs)))
        tau = sever_filter_scores(Xs, -2 * (Xs @ theta), as_written)
        p = int(eps / 2 * Xs.shape[0])
        Xs = Xs[np.argsort(-tau)[p:], :]
    return theta


# --------------------------------------------------------------------------------------------------
# The benchmark's unit of work: one E+M step of the logistic model
# --------------------------------------------------------------------------------------------------
def em_step_logistic(X, y, theta, tol=1e-3, maxiter=100):
    """(1) loss pass  utils.py:19-21 on [1, X];  (2) fixed point  rlvi.py:8-20;
    (3) weighted statistics for the MM majoriser / normal equations  utils.py:36-38.
    Returns dict(losses, pi, eps, iters, S0, S1, G)."""
    phi = theta[0] + X @ theta[1:]
    losses = -y * phi + phi + np.log1p(np.exp(-phi))
    pi, eps, iters, err = fixed_point_trace(losses, tol=tol, maxiter=maxiter)
    out = weighted_moments(X, pi)
    out.update(losses=losses, pi=pi, eps=eps, iters=iters, err=err)
    return out
