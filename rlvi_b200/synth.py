"""Seeded synthetic inputs for the parity tests and the benchmark.

Host generators use `np.random.default_rng(seed)`; the device generators (large N) use a
`torch.Generator` seeded with `seed + rank` per shard.  Distributions follow the reference's own
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
