"""Build-container only: the drop-in modules' host logic (kernels replaced by tests/abi_double.py) against the
UNMODIFIED reference functions imported from /root/reference, on seeded random contaminated problems of the
sizes the reference itself can handle (its N x N `np.diag` temporaries limit N to a few hundred here).
Skipped where the reference tree is absent (the GPU box)."""
import warnings

import numpy as np
import pytest
import torch

import abi_double
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")

CASES = [(0, 40, 1), (1, 60, 2), (2, 150, 5), (3, 300, 10), (4, 400, 3), (5, 200, 8)]


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture
def mods(monkeypatch):
    abi_double.install(monkeypatch)
    from rlvi_b200 import deep, online, rlvi, utils
    return rlvi, utils, deep, online


def problem(seed, n, d):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)) * rng.uniform(0.5, 3, size=d) + rng.normal(size=d)
    theta = rng.normal(size=d)
    out = rng.random(n) < 0.2
    Xc = X.copy()
    Xc[out] += rng.standard_t(1.5, size=(int(out.sum()), d)) * 5            # heavy-tailed contamination
    y = X @ theta + 0.3 * rng.normal(size=n)
    y[out] += rng.standard_t(1.5, size=int(out.sum())) * 10
    yl = (rng.random(n) < 1 / (1 + np.exp(-(X @ theta)))).astype(float)       # noisy labels: not separable
    yl[out] = 1 - yl[out]
    return rng, X, Xc, y, yl


@pytest.mark.parametrize("seed,n,d", CASES)
def test_standard_learning_functions(mods, seed, n, d):
    rlvi, utils, _, _ = mods
    R, RU = ref_shim.standard()
    rng, X, Xc, y, yl = problem(seed, n, d)
    w = rng.random(n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert relmax(rlvi.mean(Xc), R.mean(Xc.copy())) < 1e-12
        assert relmax(rlvi.linear_regression(X, y), R.linear_regression(X.copy(), y.copy())) < 1e-10
        # default M-step: liblinear on the reference side, MM on the same objective here -> liblinear's accuracy
        assert relmax(rlvi.logistic_regression(X, yl), R.logistic_regression(X.copy(), yl.copy())) < 5e-3
        a, b = utils.mm_log_reg(X, yl, w.copy()), RU.mm_log_reg(X.copy(), yl.copy(), w.copy())
        assert relmax(a[0], b[0]) < 1e-12 and relmax(a[1], b[1]) < 1e-12
        w_ours, w_ref = w.copy(), w.copy()
        a, b = utils.sklearn_log_reg(X, yl, w_ours), RU.sklearn_log_reg(X.copy(), yl.copy(), w_ref)
        assert np.array_equal(w_ours, w_ref) and relmax(a[0], b[0]) < 5e-3 and relmax(a[1], b[1]) < 5e-3
        if d >= 2:
            a, b = utils.pca(Xc, w), RU.pca(Xc.copy(), w.copy())
            assert relmax(a[0], b[0]) < 1e-10 and relmax(a[1], b[1]) < 1e-10
            t0 = np.ones(d) / np.sqrt(d)
            assert relmax(rlvi.pca(Xc, theta_init=t0), R.pca(Xc.copy(), theta_init=t0.copy())) < 1e-9
        a, b = utils.covariance(Xc, w), RU.covariance(Xc.copy(), w.copy())
        assert relmax(a[0], b[0]) < 1e-12 and relmax(a[1], b[1]) < 1e-11
        assert relmax(rlvi.covariance(Xc, 0.3), R.covariance(Xc.copy(), 0.3)) < 1e-6
        losses = np.abs(rng.normal(size=n)) * 3
        assert relmax(rlvi.update_weights(losses), R.update_weights(losses.copy())) < 1e-13
        assert relmax(rlvi.update_weights_constrained(losses, 0.8 * n),
                      R.update_weights_constrained(losses.copy(), 0.8 * n)) < 1e-6
        Xa = np.hstack([np.ones((n, 1)), X])
        th = rng.normal(size=d + 1)
        assert relmax(utils.cross_entropy(Xa, th, yl), RU.cross_entropy(Xa, th, yl)) < 1e-12
        assert np.array_equal(utils.clf_predict(X, th), RU.clf_predict(X, th))
        assert relmax(utils.sigmoid(X[:, 0]), RU.sigmoid(X[:, 0])) < 1e-15


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_deep_and_online_functions(mods, seed):
    _, _, deep, online = mods
    D, O = ref_shim.deep(), ref_shim.online()
    rng = np.random.default_rng(seed)
    n = 3000
    res = rng.exponential(1.0, size=n).astype(np.float32)
    res[rng.random(n) < 0.3] += 4.0
    r1, w1 = torch.from_numpy(res.copy()), torch.ones(n)
    r2, w2 = torch.from_numpy(res.copy()), torch.ones(n)
    deep.update_sample_weights(r1, w1)
    D.update_sample_weights(r2, w2)
    assert torch.allclose(r1, r2, rtol=0, atol=1e-6) and torch.allclose(w1, w2, rtol=0, atol=1e-5)
    assert abs(float(deep.false_negative_criterion(w1)) - float(D.false_negative_criterion(w2))) < 1e-5
    assert float(deep.false_negative_criterion(w2, alpha=0.2)) == float(D.false_negative_criterion(w2, alpha=0.2))
    lp = np.log(rng.uniform(0.05, 0.95, size=100))
    t = (rng.random(100) < 0.5).astype(float)
    ours, ref = online.cross_entropy(lp, t), O.cross_entropy(lp, t)
    assert np.array_equal(ours, ref)
    assert relmax(online.update_weights_rlvi(ours), O.update_weights_rlvi(ref)) < 1e-13
    assert relmax(online.update_weights_rlvi(ours, 1e-5, 7), O.update_weights_rlvi(ref, 1e-5, 7)) < 1e-13


def test_input_conventions_and_degenerate_designs(mods):
    """What NumPy callers may pass (lists, integer / float32 / Fortran-ordered arrays) and the designs lstsq
    handles silently (more features than samples, a zero column): same answers as the reference."""
    rlvi, utils, _, _ = mods
    R, RU = ref_shim.standard()
    rng = np.random.default_rng(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        X, y = rng.normal(size=(6, 10)), rng.normal(size=6)                  # underdetermined: minimum norm
        assert relmax(rlvi.linear_regression(X, y), R.linear_regression(X.copy(), y.copy())) < 1e-10
        X = rng.normal(size=(100, 4))
        X[:, 2] = 0                                                         # zero column
        y = X @ np.array([1.0, 2.0, 0.0, -1.0]) + 0.1 * rng.normal(size=100)
        assert relmax(rlvi.linear_regression(X, y), R.linear_regression(X.copy(), y.copy())) < 1e-10
        Xi = rng.integers(-5, 5, size=(50, 3))
        yi = Xi @ np.array([1, 2, 3]) + rng.integers(-1, 2, size=50)
        assert relmax(rlvi.linear_regression(Xi, yi), R.linear_regression(Xi.copy(), yi.copy())) < 1e-10
        losses = np.abs(rng.normal(size=30))
        assert relmax(rlvi.update_weights(list(losses)), R.update_weights(losses.copy())) < 1e-13
        assert relmax(rlvi.update_weights(losses, maxiter=1), R.update_weights(losses.copy(), maxiter=1)) < 1e-13
        assert relmax(rlvi.update_weights(losses, tol=10.0), R.update_weights(losses.copy(), tol=10.0)) < 1e-13
        X32 = rng.normal(size=(50, 3)).astype(np.float32)
        ours = rlvi.mean(X32)
        assert ours.dtype == np.float64 and relmax(ours, R.mean(X32.copy())) < 1e-12
        Xf = np.asfortranarray(rng.normal(size=(80, 4)))
        yf = Xf @ np.ones(4) + rng.normal(size=80)
        assert relmax(rlvi.linear_regression(Xf, yf), R.linear_regression(Xf.copy(), yf.copy())) < 1e-10
        same = np.full((20, 2), 3.0)                                         # sigma2 = 0: NaN on both sides (Q1)
        assert np.isnan(rlvi.mean(same)).all() and np.isnan(R.mean(same.copy())).all()
        Xc = rng.normal(size=(120, 3))
        for eps in (0.0, 0.05, 0.6):
            assert relmax(rlvi.covariance(Xc, eps), R.covariance(Xc.copy(), eps)) < 1e-6
        assert relmax(utils.sigmoid(Xc), RU.sigmoid(Xc)) < 1e-15             # any shape
