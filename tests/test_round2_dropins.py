"""Round-2 drop-in pieces against the goldens of the unmodified reference (oracle/make_golden_r02.py):
the Newton / IRLS M-step behind utils.sklearn_log_reg, rlvi.logistic_regression's default route end to end, the RRM
weight rule, config 1's Monte-Carlo replay, the N-sized sigmoid and the online cross-entropy.

Every test runs twice: on the CPU tier with the kernels replaced by the test double (tests/abi_double.py: the host
logic -- Newton iteration, line search, Brent, loops), and with `-m gpu` through librlvi_b200.so."""
import numpy as np
import pytest
import torch

import abi_double
from conftest import load_golden
from oracle import rlvi_np


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(params=["double", pytest.param("cuda", marks=pytest.mark.gpu)])
def mods(request, monkeypatch):
    if request.param == "double":
        abi_double.install(monkeypatch)
    else:
        from rlvi_b200 import _lib
        _lib.load()
        assert torch.cuda.is_available()
    from rlvi_b200 import online, rlvi, rrm, sever, utils

    class NS:
        pass
    ns = NS()
    ns.rlvi, ns.utils, ns.rrm, ns.online, ns.sever, ns.kind = rlvi, utils, rrm, online, sever, request.param
    return ns


def liblinear_objective(X, y, w, theta, C=100.0):
    phi = theta[0] + X @ theta[1:]
    return 0.5 * theta @ theta + C * np.sum(w * (np.logaddexp(0.0, phi) - y * phi))


@pytest.mark.parametrize("tag", ["sklearn_loss_n600_d5", "sklearn_sep_n800_d6"])
def test_newton_mstep_against_liblinear(mods, tag):
    """theta minimises liblinear's own objective at least as well as liblinear's answer (it stops at tol = 1e-4), agrees
    with it to liblinear's accuracy, and the Newton iteration needs few passes over X -- also on nearly separable data,
    where the fixed-curvature MM iteration of round 1 crawled."""
    g = load_golden(tag)
    w = g["w"].copy()
    theta, losses = mods.utils.sklearn_log_reg(g["X"], g["y"], w)
    assert np.array_equal(w, g["w_after"])                                  # in-place normalisation (quirk Q3)
    wn = g["w_after"]
    f_mine, f_ref = liblinear_objective(g["X"], g["y"], wn, theta), liblinear_objective(g["X"], g["y"], wn, g["theta"])
    assert f_mine <= f_ref * (1 + 1e-12)
    assert relmax(theta, g["theta"]) < 5e-3
    assert relmax(losses, rlvi_np.softplus_loss(g["X"], theta)) < 1e-12      # label-independent loss (quirk Q3)
    assert relmax(losses, g["losses"]) < 5e-3
    assert mods.utils.sklearn_log_reg.last_passes <= 40
    # the minimiser is unique: a warm start lands on the same theta
    theta2, _ = mods.utils.sklearn_log_reg(g["X"], g["y"], g["w"].copy(), theta0=theta + 0.3)
    assert relmax(theta2, theta) < 1e-8
    # first-order optimality of the regularised objective, checked independently in NumPy
    phi = theta[0] + g["X"] @ theta[1:]
    c = wn * (rlvi_np.sigmoid(phi) - g["y"])
    grad = theta + 100.0 * np.concatenate([[c.sum()], g["X"].T @ c])
    assert np.linalg.norm(grad) < 1e-6 * (1 + 100.0 * np.linalg.norm(np.concatenate([[np.abs(c).sum()], np.abs(g["X"]).T @ np.abs(c)])))


def test_logistic_regression_default_route_end_to_end(mods):
    """rlvi.logistic_regression with the reference's DEFAULT M-step (liblinear, rlvi.py:96,103) on a golden produced by
    the unmodified reference: the EM loop is driven by liblinear-accurate thetas on both sides."""
    g = load_golden("logreg_sklearn_n2000_d8")
    theta = mods.rlvi.logistic_regression(g["X"], g["y"])
    assert theta.shape == g["theta"].shape
    assert relmax(theta, g["theta"]) < 2e-2
    cos = theta[1:] @ g["theta"][1:] / np.linalg.norm(theta[1:]) / np.linalg.norm(g["theta"][1:])
    assert cos > 1 - 1e-4


def test_rrm_weights_and_linear_regression(mods):
    g = load_golden("rrm_weights_n2000")
    w = mods.rrm.update_weights(g["losses"], float(g["eps"]))
    assert isinstance(w, np.ndarray) and relmax(w, g["weights"]) < 1e-6      # Brent's own accuracy on alpha
    assert abs(w.sum() - g["weights"].sum()) < 1e-6
    g = load_golden("rrm_linreg_n300_d6")
    assert relmax(mods.rrm.linear_regression(g["X"], g["y"], float(g["eps"])), g["theta"]) < 1e-6
    X, _ = np.split(g["X"], [300], axis=0)
    assert mods.rrm.mean(g["X"], 0.2).shape == (6,)


def test_sever_filter(mods):
    """standard-learning/sever.py:11-42 and :82-113 on the statistics kernels (the active set as a 0/1 weight vector, the top
    right singular vector of the centred gradients as the top eigenvector of their d x d scatter).  The oracle restates the
    reference literally and is pinned bit for bit on its goldens; the drop-in implements the algorithm the reference's
    comment names -- line 27 of the reference takes a COLUMN of NumPy's V^H (quirk Q12, LAPACK-sign dependent, not
    reproducible without an SVD of the n x d gradient matrix) -- and is compared with the oracle's `as_written=False`."""
    g = load_golden("sever_linreg_n600_d8")
    X, y, eps = g["X"], g["y"], float(g["eps"])
    assert relmax(rlvi_np.sever_linear_regression(X, y, eps), g["theta"]) == 0.0
    th = mods.sever.linear_regression(X, y, eps)
    assert isinstance(th, np.ndarray) and relmax(th, rlvi_np.sever_linear_regression(X, y, eps, as_written=False)) < 1e-9
    g = load_golden("sever_pca_n500_d6")
    S, eps = g["samples"], float(g["eps"])
    assert relmax(rlvi_np.sever_pca(S, eps), g["theta"]) == 0.0
    assert relmax(mods.sever.pca(S, eps), rlvi_np.sever_pca(S, eps, as_written=False)) < 1e-9
    # theta_init replaces the first base fit only (sever.py:90-91)
    t0 = np.ones(6) / np.sqrt(6.0)
    assert relmax(mods.sever.pca(S, 0.3, theta_init=t0), rlvi_np.sever_pca(S, 0.3, theta_init=t0, as_written=False)) < 1e-9
    # eps small enough that nothing is filtered: the plain least-squares fit
    g = load_golden("sever_linreg_n600_d8")
    th = mods.sever.linear_regression(g["X"], g["y"], 1e-4)
    assert relmax(th, np.linalg.lstsq(g["X"], g["y"], rcond=None)[0]) < 1e-9


def test_config1_monte_carlo_replay(mods):
    """BASELINE.json configs[0] through the drop-in: the 100 problems of standard-learning/main.py:308-357 (epsilon =
    0.2), per-run theta against the reference's own estimate at 1e-9, and BASELINE.md section 2a's fingerprint
    (mean 0.02654 / median 0.02625) to three digits; RRM on the same problems."""
    g = load_golden("config1_linreg_fixed_eps")
    err = []
    for X, y, th in zip(g["X"], g["y"], g["theta_rlvi"]):
        mine = mods.rlvi.linear_regression(X, y)
        assert relmax(mine, th) < 1e-9
        err.append(np.linalg.norm(1.0 - mine) / np.sqrt(10.0))
    assert abs(np.mean(err) - 0.02654) < 5e-5 and abs(np.median(err) - 0.02625) < 5e-5
    for k in range(0, 100, 20):
        assert relmax(mods.rrm.linear_regression(g["X"][k], g["y"][k], 0.4), g["theta_rrm"][k]) < 1e-6


def test_sigmoid_any_size_and_online_cross_entropy(mods):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=100003) * 30, [0.0, -800.0, 800.0, -1e-300]])
    assert relmax(mods.utils.sigmoid(x), rlvi_np.sigmoid(x)) < 1e-15
    assert abs(mods.utils.sigmoid(0.3) - float(rlvi_np.sigmoid(0.3))) < 1e-16 and isinstance(mods.utils.sigmoid(0.3), float)
    m = rng.normal(size=(7, 5))
    assert mods.utils.sigmoid(m).shape == (7, 5)
    lp = -rng.random(100)
    t = (rng.random(100) < 0.5).astype(np.float64)
    assert np.array_equal(mods.online.cross_entropy(lp, t), rlvi_np.online_cross_entropy(lp, t))


def test_constrained_estep_outliers_whose_e_underflows(mods):
    """the Brent objective runs on the cached e = exp(-l), but the RETURNED weights use the literal
    exp(-l + s) of rlvi.py:42 -- far outliers (l > 708, e denormal or 0) keep their tiny but representable weights."""
    rng = np.random.default_rng(31)
    losses = np.concatenate([25.0 + 0.5 * rng.chisquare(1, size=2000), [730.0, 741.0]])
    n_eff = 0.9 * len(losses)
    pi = mods.rlvi.update_weights_constrained(losses, n_eff)
    ref = rlvi_np.update_weights_constrained(losses, n_eff)
    assert relmax(pi, ref) < 1e-6                               # Brent's own accuracy
    assert np.all(ref[-2:] > 0) and np.all(pi[-2:] > 0)
    assert np.max(np.abs(pi[-2:] / ref[-2:] - 1)) < 1e-5        # element-wise on the two outliers (denormal e would be ~1e-2 off)


def test_linear_regression_conditioning_bound(mods):
    """the M-step solves the d x d normal equations, so the error grows like cond(X)^2 eps where the
    reference's gelsd grows like cond(X) eps.  cond(X) = 1e3: still 1e-9; cond(X) = 1e5: bounded by 100 cond^2 eps."""
    rng = np.random.default_rng(32)
    n, d = 4000, 8
    U, _ = np.linalg.qr(rng.normal(size=(n, d)))
    V, _ = np.linalg.qr(rng.normal(size=(d, d)))
    for cond, bound in ((1e3, 1e-9), (1e5, 100 * 1e10 * 2.2e-16)):
        X = (U * np.geomspace(1.0, 1.0 / cond, d)) @ V.T * np.sqrt(n)
        y = X @ np.ones(d) + 0.1 * rng.normal(size=n)
        assert relmax(mods.rlvi.linear_regression(X, y), rlvi_np.linear_regression(X, y)) < bound


@pytest.mark.gpu
def test_wce_invalid_label_or_index_poisons_instead_of_reading_out_of_bounds():
    """a label outside [0, C) (ignore_index = -100, a corrupted label) or an index outside [0, n_train) raises
    in the reference; the kernel never touches memory out of bounds and turns the row -- and the batch loss -- into NaN."""
    from rlvi_b200 import ops
    dev = torch.device("cuda", 0)
    b, c, n_train = 64, 10, 100
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(b, c, generator=g).to(dev)
    labels = torch.randint(0, c, (b,), generator=g).to(dev)
    idx = torch.randperm(n_train, generator=g)[:b].to(dev)
    w = torch.rand(n_train, generator=g).to(dev)
    res = torch.full((n_train + 8,), 7.0, device=dev)[:n_train]          # guard band behind the view
    ok = ops.wce_fwd_bwd(logits, labels, w, res, indexes=idx, want_per_sample=True)
    assert torch.isfinite(ok["loss"]).all() and torch.isfinite(ok["dlogits"]).all()
    bad_labels = labels.clone()
    bad_labels[3], bad_labels[17] = -100, c + 5
    out = ops.wce_fwd_bwd(logits, bad_labels, w, res.clone(), indexes=idx, want_per_sample=True)
    assert torch.isnan(out["loss"]).all() and torch.isnan(out["per_sample"][[3, 17]]).all()
    assert torch.isnan(out["dlogits"][[3, 17]]).all()
    keep = torch.ones(b, dtype=torch.bool, device=dev)
    keep[[3, 17]] = False
    assert torch.equal(out["dlogits"][keep], ok["dlogits"][keep]) and torch.equal(out["per_sample"][keep], ok["per_sample"][keep])
    bad_idx = idx.clone()
    bad_idx[5] = n_train + 3
    out = ops.wce_fwd_bwd(logits, labels, w, res.clone(), indexes=bad_idx)
    assert torch.isnan(out["loss"]).all()
