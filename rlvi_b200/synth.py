"""Seeded synthetic inputs for the parity tests and the benchmark.

Host generators use `np.random.default_rng(seed)`; the device generators (large N) draw every block of 2^20
global rows from its own `torch.Generator` seeded by (seed, block index), so a row's values depend on its GLOBAL
index only: 1, 2, 4 or 8 ranks sharding the same N see exactly the same samples.  Distributions follow the reference's own
generators where one exists (standard-learning/main.py:44-167) generalised to d features.
"""
from __future__ import annotations

import numpy as np


def losses_mixture(n, seed=0, clean_scale=0.5, outlier_scale=50.0, frac=0.3):
    """70 % 0.5*chi2_1 + 30 % 50*chi2_1 losses (the mixture BASELINE.md section 2b times)."""
    rng = np.random.default_rng(seed)
    out = clean_scale * rng.chisquare(1, size=n)
    mask = rng.random(n) < frac
    out[mask] = outlier_scale * rng.chisquare(1, size=int(mask.sum()))
    return out


def linear_regression_data(n=40, d=10, eps=0.2, nu=2.5, seed=0):
    """X ~ U[-5,5]^{n x d}, theta* = 1, sigma = 0.25; corrupted rows get Student-t(nu) noise
    (standard-learning/main.py:69-85)."""
    rng = np.random.default_rng(seed)
    X = -5 + 10 * rng.random(size=(n, d))
    n2 = rng.binomial(n=n, p=eps)
    n1 = n - n2
    theta = np.ones(d)
    y = X @ theta
    y[:n1] += 0.25 * rng.normal(size=n1)
    u = rng.chisquare(df=nu, size=n2) / nu
    y[n1:] += rng.normal(size=n2) / np.sqrt(u)
    return np.ascontiguousarray(X), y


def mean_data(n=100, d=2, eps=0.2, seed=0):
    """Clean rows ~ N(200*1, S), corrupted rows heavy-tailed around the same mean
    (standard-learning/main.py:44-66 generalised to d)."""
    rng = np.random.default_rng(seed)
    n2 = rng.binomial(n=n, p=eps)
    n1 = n - n2
    S = 350.0 * np.eye(d) + 50.0
    L = np.linalg.cholesky(S)
    clean = 200.0 + rng.normal(size=(n1, d)) @ L.T
    nu = 2.5
    u = rng.chisquare(df=nu, size=(n2, 1)) / nu
    Lc = np.linalg.cholesky(2 * (nu / (nu - 2)) * S)
    bad = 200.0 + (rng.normal(size=(n2, d)) @ Lc.T) / np.sqrt(u)
    return np.ascontiguousarray(np.vstack([clean, bad]))


def logistic_data(n, d=64, corruption=0.3, seed=0):
    """Config C2-logistic: X ~ N(0,1), theta* ~ N(0,1/d), y ~ Bernoulli(sigmoid(X theta*)), then a
    uniformly random `corruption` fraction of labels flipped."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    theta = rng.normal(size=d) / np.sqrt(d)
    p = 1.0 / (1.0 + np.exp(-(X @ theta)))
    y = (rng.random(n) < p).astype(np.float64)
    flip = rng.random(n) < corruption
    y[flip] = 1.0 - y[flip]
    return np.ascontiguousarray(X), y, theta


def pca_data(n=200, d=2, eps=0.2, seed=0):
    """Clean rows = z*v + 0.25*N(0,I) along a fixed unit v; corrupted rows N(0,I)/sqrt(chi2_1.5/1.5)
    (standard-learning/main.py:126-143 generalised to d)."""
    rng = np.random.default_rng(seed)
    n2 = rng.binomial(n=n, p=eps)
    n1 = n - n2
    v = np.arange(1, d + 1, dtype=np.float64)
    v /= np.linalg.norm(v)
    clean = rng.normal(size=(n1, 1)) * 2.0 * v + 0.25 * rng.normal(size=(n1, d))
    nu = 1.5
    u = rng.chisquare(df=nu, size=(n2, 1)) / nu
    bad = rng.normal(size=(n2, d)) / np.sqrt(u)
    return np.ascontiguousarray(np.vstack([clean, bad])), v


def covariance_data(n=50, d=2, eps=0.2, seed=0, scale=1.0):
    """Clean rows ~ N(0, scale^2 R), R = 0.8*11^T + 0.2*I; corrupted rows the same draw divided by
    sqrt(chi2_1.5/1.5) (standard-learning/main.py:146-167 generalised to d)."""
    rng = np.random.default_rng(seed)
    n2 = rng.binomial(n=n, p=eps)
    n1 = n - n2
    R = 0.8 * np.ones((d, d)) + 0.2 * np.eye(d)
    L = scale * np.linalg.cholesky(R)
    clean = rng.normal(size=(n1, d)) @ L.T
    nu = 1.5
    u = rng.chisquare(df=nu, size=(n2, 1)) / nu
    bad = (rng.normal(size=(n2, d)) @ L.T) / np.sqrt(u)
    return np.ascontiguousarray(np.vstack([clean, bad])), R * scale ** 2


def pairflip_labels(labels, noise=0.45, nb_classes=100, seed=1):
    """Pair-flip label noise: class i -> i+1 (mod C) with probability `noise`
    (deep-learning/data_tools.py:156-177 defines the transition matrix; sampled here per label)."""
    rng = np.random.default_rng(seed)
    flip = rng.random(labels.shape[0]) < noise
    noisy = labels.copy()
    noisy[flip] = (labels[flip] + 1) % nb_classes
    return noisy, flip


def deep_batch(b=8192, c=100, seed=1):
    """Config C5 micro-batch: logits = 2*randn(b, c) FP32, uniform clean labels + pairflip 0.45."""
    rng = np.random.default_rng(seed)
    logits = (2.0 * rng.normal(size=(b, c))).astype(np.float32)
    clean = rng.integers(0, c, size=b)
    labels, _ = pairflip_labels(clean, 0.45, c, seed)
    return logits, labels.astype(np.int64)


def logistic_shard_torch(n, d, device, seed, corruption=0.3, chunk=1 << 22, dtype=None):
    """Device-side version of `logistic_data` for the benchmark (N = 2^26 per GPU): generated in
    chunks so the temporaries stay small.  Returns (X [n,d], y [n], theta* [d]) on `device`."""
    import torch

    dtype = dtype or torch.float64
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    theta = torch.randn(d, generator=g, device=device, dtype=torch.float64) / (d ** 0.5)
    X = torch.empty((n, d), device=device, dtype=dtype)
    y = torch.empty((n,), device=device, dtype=dtype)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        xb = X[s:s + m]
        xb.normal_(generator=g)
        p = torch.sigmoid(xb.to(torch.float64) @ theta)
        yb = (torch.rand(m, generator=g, device=device, dtype=torch.float64) < p).to(torch.float64)
        flip = torch.rand(m, generator=g, device=device, dtype=torch.float64) < corruption
        yb = torch.where(flip, 1.0 - yb, yb)
        y[s:s + m] = yb.to(dtype)
    return X, y, theta


ROW_BLOCK = 1 << 20      # rows per independently seeded block of the device generators


def theta_star_torch(d, device, seed):
    """theta* ~ N(0, 1/d) of the logistic workload: a function of (d, seed) only, identical on every rank."""
    import torch

    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) * 7919 + 17)
    return (torch.randn(d, generator=g, dtype=torch.float64) / (d ** 0.5)).to(device)


def logistic_rows_torch(row_lo, row_hi, d, device, seed, corruption=0.3, dtype=None):
    """Rows [row_lo, row_hi) of the GLOBAL synthetic logistic data set (config C2-logistic: X ~ N(0,1),
    y ~ Bernoulli(sigmoid(X theta*)), then `corruption` of the labels flipped), on `device`.  Block b of ROW_BLOCK
    rows is drawn from a generator seeded with (seed, b), so the result does not depend on how the rows are sharded.
    Returns (X [rows, d], y [rows], theta* [d])."""
    import torch

    dtype = dtype or torch.float64
    theta = theta_star_torch(d, device, seed)
    n = row_hi - row_lo
    X = torch.empty((n, d), device=device, dtype=dtype)
    y = torch.empty((n,), device=device, dtype=torch.float64)
    g = torch.Generator(device=device)
    b0, b1 = row_lo // ROW_BLOCK, (row_hi + ROW_BLOCK - 1) // ROW_BLOCK
    for b in range(b0, b1):
        g.manual_seed(int(seed) * 1000003 + b)
        xb = torch.empty((ROW_BLOCK, d), device=device, dtype=dtype).normal_(generator=g)
        p = torch.sigmoid(xb.to(torch.float64) @ theta)
        yb = (torch.rand(ROW_BLOCK, generator=g, device=device, dtype=torch.float64) < p).to(torch.float64)
        flip = torch.rand(ROW_BLOCK, generator=g, device=device, dtype=torch.float64) < corruption
        yb = torch.where(flip, 1.0 - yb, yb)
        lo, hi = max(row_lo, b * ROW_BLOCK), min(row_hi, (b + 1) * ROW_BLOCK)
        X[lo - row_lo:hi - row_lo] = xb[lo - b * ROW_BLOCK:hi - b * ROW_BLOCK]
        y[lo - row_lo:hi - row_lo] = yb[lo - b * ROW_BLOCK:hi - b * ROW_BLOCK]
        del xb, p, yb, flip
    return X, y, theta


def pca_rows_torch(row_lo, row_hi, d, device, seed, eps=0.2, dtype=None):
    """Rows [row_lo, row_hi) of the GLOBAL config-3 data set: clean rows z v + 0.25 N(0, I)
    along the fixed unit direction v ~ (1, 2, ..., d); a fraction `eps` of the rows replaced by N(0, I) / sqrt(chi2_1.5 /
    1.5) (standard-learning/main.py:126-143 generalised to d features; corruption by a per-row coin instead of a block
    of trailing rows so that every shard sees the same mixture).  Sharding-invariant like `logistic_rows_torch`.
    Returns (X [rows, d] `dtype` (default float32), v [d] float64)."""
    import torch

    dtype = dtype or torch.float32
    v = torch.arange(1, d + 1, dtype=torch.float64, device=device)
    v = v / v.norm()
    n = row_hi - row_lo
    X = torch.empty((n, d), device=device, dtype=dtype)
    g = torch.Generator(device=device)
    nu = 1.5
    b0, b1 = row_lo // ROW_BLOCK, (row_hi + ROW_BLOCK - 1) // ROW_BLOCK
    for b in range(b0, b1):
        g.manual_seed(int(seed) * 1000003 + b)
        xb = torch.empty((ROW_BLOCK, d), device=device, dtype=torch.float32).normal_(generator=g)
        z = torch.empty((ROW_BLOCK, 1), device=device, dtype=torch.float32).normal_(generator=g)
        bad = torch.rand((ROW_BLOCK, 1), generator=g, device=device) < eps
        # chi2_nu / nu from a Gamma(nu / 2, 2 / nu) draw: -log(U) has no closed form for nu = 1.5, so use the
        # sum-of-squares identity chi2_1.5 ~ Gamma(0.75, 2) via torch's gamma sampler on the host-independent generator
        u = torch._standard_gamma(torch.full((ROW_BLOCK, 1), nu / 2, device=device), generator=g) * (2.0 / nu)
        clean = 2.0 * z * v.to(torch.float32) + 0.25 * xb
        xb = torch.where(bad, xb / torch.sqrt(u.to(torch.float32)), clean)
        lo, hi = max(row_lo, b * ROW_BLOCK), min(row_hi, (b + 1) * ROW_BLOCK)
        X[lo - row_lo:hi - row_lo] = xb[lo - b * ROW_BLOCK:hi - b * ROW_BLOCK].to(dtype)
        del xb, z, bad, u, clean
    return X, v
