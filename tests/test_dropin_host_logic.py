"""Host logic of the drop-in modules on the CPU tier: the outer EM loops, stop tests, Brent search, d x d
algebra, sign rules, in-place quirks and NumPy-in/NumPy-out conversion of rlvi_b200/{rlvi,utils,deep,online}.py,
run against the golden vectors of the unmodified reference with the kernels replaced by the test double in
tests/abi_double.py (which follows the contract in include/rlvi_b200.h).  The kernels themselves meet the same
goldens in tests/test_gpu_parity.py (-m gpu); nothing here is a product path.
"""
import numpy as np
import pytest
import torch

import abi_double
from conftest import load_golden
from oracle import deep_ref, rlvi_np

F64_TOL = 1e-9
F32_TOL = 1e-5


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture
def dropin(monkeypatch):
    abi_double.install(monkeypatch)
    from rlvi_b200 import deep, online, rlvi, utils

    class NS:
        pass
    ns = NS()
    ns.rlvi, ns.utils, ns.deep, ns.online = rlvi, utils, deep, online
    return ns


@pytest.mark.parametrize("tag", ["n40", "n1000", "n4096", "n16384"])
def test_update_weights(dropin, tag):
    g = load_golden("update_weights_" + tag)
    pi = dropin.rlvi.update_weights(g["losses"])
    assert isinstance(pi, np.ndarray) and pi.dtype == np.float64 and pi.shape == g["pi"].shape
    assert relmax(pi, g["pi"]) < 1e-12
    t = torch.from_numpy(g["losses"])
    out = dropin.rlvi.update_weights(t)                       # tensor in -> tensor out
    assert isinstance(out, torch.Tensor) and relmax(out.numpy(), g["pi"]) < 1e-12


def test_update_weights_keywords_and_negative_losses(dropin):
    g = load_golden("update_weights_neg")                      # negative losses, non-default tol / maxiter
    tol, maxiter = float(g["tol"]), int(g["maxiter"])
    assert relmax(dropin.rlvi.update_weights(g["losses"], tol=tol, maxiter=maxiter), g["pi"]) < 1e-12
    assert relmax(dropin.rlvi.update_weights(g["losses"], tol, maxiter), g["pi"]) < 1e-12       # positional order
    assert relmax(dropin.rlvi.update_weights(g["losses"], 1e-2, 3), rlvi_np.update_weights(g["losses"], 1e-2, 3)) < 1e-12


@pytest.mark.parametrize("tag", ["n50", "n3000"])
def test_update_weights_constrained(dropin, tag):
    g = load_golden("constrained_" + tag)
    pi = dropin.rlvi.update_weights_constrained(g["losses"], float(g["n_eff"]))
    assert relmax(pi, g["pi"]) < 1e-6                            # Brent's own accuracy (H4)
    assert abs(pi.sum() - float(g["n_eff"])) < 1e-4 * float(g["n_eff"]) or pi.sum() >= float(g["n_eff"])
    # branch not taken: sum(pi) >= n_eff returns the plain fixed point (quirk Q7)
    free = dropin.rlvi.update_weights(g["losses"])
    same = dropin.rlvi.update_weights_constrained(g["losses"], 0.5 * free.sum())
    assert np.array_equal(same, free)


@pytest.mark.parametrize("tag", ["n100_d2", "n768_d64"])
def test_mean(dropin, tag):
    g = load_golden("mean_" + tag)
    theta = dropin.rlvi.mean(g["X"])
    assert isinstance(theta, np.ndarray) and relmax(theta, g["theta"]) < F64_TOL


@pytest.mark.parametrize("tag", ["n40_d10", "n768_d64"])
def test_linear_regression(dropin, tag):
    g = load_golden("linreg_" + tag)
    assert relmax(dropin.rlvi.linear_regression(g["X"], g["y"]), g["theta"]) < F64_TOL


def test_linear_regression_rank_deficient_design_takes_the_minimum_norm_solution(dropin):
    """rlvi.py:71 uses lstsq (gelsd): a duplicated column gives the minimum-norm theta, which the normal-equation
    route reproduces through the eigen-decomposition pseudo-inverse (the Cholesky factorisation fails)."""
    rng = np.random.default_rng(3)
    X = rng.normal(size=(200, 4))
    X = np.hstack([X, X[:, :1]])                                  # column 4 == column 0
    y = X @ np.array([1.0, -2.0, 0.5, 3.0, 1.0]) + 0.01 * rng.normal(size=200)
    ref = rlvi_np.linear_regression(X, y)
    got = dropin.rlvi.linear_regression(X, y)
    assert relmax(got, ref) < 1e-6
    assert abs(got[0] - got[4]) < 1e-8 * abs(got[0])              # minimum norm: equal split over the twins


def test_linear_regression_badly_scaled_features_stay_full_rank(dropin):
    """Columns whose scales differ by 1e7 are independent, not rank-deficient: gelsd solves them, and so must
    the Gram route (the rank test looks at scale-free pivots)."""
    rng = np.random.default_rng(5)
    X = rng.normal(size=(300, 4)) * np.array([1.0, 1e-7, 1e3, 1.0])
    theta_true = np.array([1.0, 2e7, -3e-3, 0.5])
    y = X @ theta_true + 0.1 * rng.normal(size=300)
    ref = rlvi_np.linear_regression(X, y)
    got = dropin.rlvi.linear_regression(X, y)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-6


def test_mm_log_reg_and_logistic_regression(dropin):
    g = load_golden("mm_log_reg_n768_d64")
    assert relmax(dropin.utils.sigmoid(g["x_sig"]), g["sig"]) < 1e-15
    theta, losses = dropin.utils.mm_log_reg(g["X"], g["y"], g["w"])
    assert relmax(theta, g["theta"]) < F64_TOL and relmax(losses, g["losses"]) < F64_TOL
    Xa = np.hstack([np.ones((g["X"].shape[0], 1)), g["X"]])
    assert relmax(dropin.utils.cross_entropy(Xa, g["theta"], g["y"]), g["ce_at_theta"]) < 1e-12
    g2 = load_golden("logreg_mm_n1500_d8")
    assert relmax(dropin.rlvi.logistic_regression(g2["X"], g2["y"], mstep="mm"), g2["theta"]) < 1e-8


def test_sklearn_log_reg_quirks(dropin):
    g = load_golden("sklearn_loss_n600_d5")
    w = g["w"].copy()
    theta, losses = dropin.utils.sklearn_log_reg(g["X"], g["y"], w)
    assert np.array_equal(w, g["w_after"])                        # caller's weights normalised in place (Q3)
    assert relmax(theta, g["theta"]) < 5e-3                       # liblinear's own stopping accuracy
    assert relmax(losses, rlvi_np.softplus_loss(g["X"], theta)) < 1e-12   # label-independent loss (Q3)
    assert relmax(losses, g["losses"]) < 5e-3


def test_clf_predict(dropin):
    rng = np.random.default_rng(12)
    X, theta = rng.normal(size=(500, 7)), rng.normal(size=8)
    Xa = np.hstack([np.ones((500, 1)), X])
    out = dropin.utils.clf_predict(X, theta)
    assert out.dtype.kind == "i" and np.array_equal(out, np.array(rlvi_np.sigmoid(Xa @ theta) > 0.5, dtype=int))
    assert np.array_equal(dropin.utils.clf_predict(X, theta[1:], augment=False),
                          np.array(rlvi_np.sigmoid(X @ theta[1:]) > 0.5, dtype=int))


@pytest.mark.parametrize("tag", ["n400_d2", "n768_d64"])
def test_pca(dropin, tag):
    g = load_golden("pca_" + tag)
    theta, losses = dropin.utils.pca(g["X"], g["w"])
    assert relmax(theta, g["theta_mstep"]) < 1e-8 and relmax(losses, g["losses_mstep"]) < 1e-8
    _, l0 = dropin.utils.pca(g["X"], g["w"], g["theta_init"])
    assert relmax(l0, g["losses_init"]) < 1e-12
    assert relmax(dropin.rlvi.pca(g["X"], theta_init=g["theta_init"]), g["theta"]) < 1e-7


@pytest.mark.parametrize("tag", ["n50_d2", "n2048_d16"])
def test_covariance(dropin, tag):
    g = load_golden("cov_" + tag)
    cov, losses = dropin.utils.covariance(g["X"], g["w"])
    assert relmax(cov, g["cov_mstep"]) < F64_TOL and relmax(losses, g["losses_mstep"]) < F64_TOL
    assert relmax(dropin.rlvi.covariance(g["X"], float(g["eps"])), g["cov"]) < 1e-5
    with pytest.raises(ValueError, match="Singular covariance matrix"):
        dropin.utils.covariance(np.ones((20, 3)), np.ones(20))


def test_online(dropin):
    g = load_golden("online_n100")
    res = dropin.online.cross_entropy(g["log_proba"], g["targets"])            # main.py:84-85 (label-independent)
    assert np.array_equal(res, g["residuals"])
    w = dropin.online.update_weights_rlvi(res)
    assert isinstance(w, np.ndarray) and relmax(w, g["pi"]) < 1e-12
    w2, avg = dropin.online.update_weights_rlvi(res, return_avg=True)
    assert np.array_equal(w, w2) and 0.0 < avg < 1.0
    w3 = dropin.online.update_weights_rlvi(res, init_weight=avg)               # opt-in carry-over (Q11 extension)
    assert relmax(w3 / w3.sum(), w / w.sum()) < 1e-2


def test_deep_estep_threshold_and_mask(dropin):
    g = load_golden("deep_estep_n45000")
    res = torch.from_numpy(g["residuals"].copy())
    w = torch.ones(res.numel(), dtype=torch.float32)
    assert dropin.deep.update_sample_weights(res, w) is None
    assert relmax(res.numpy(), g["residuals_after"]) < F32_TOL
    assert np.max(np.abs(w.numpy() - g["weights"])) < F32_TOL
    thr = dropin.deep.false_negative_criterion(w)
    assert thr.dim() == 0 and abs(float(thr) - float(g["threshold"])) <= F32_TOL
    assert np.array_equal(dropin.deep.selection_mask(w, thr).numpy(), g["weights"] > float(thr))
    gw = load_golden("deep_threshold_wrap")
    assert float(dropin.deep.false_negative_criterion(torch.from_numpy(gw["weights"]))) == float(gw["threshold"])


def test_deep_weighted_ce_autograd(dropin):
    g = load_golden("deep_wce_b512_c100")
    b = g["logits"].shape[0]
    idx = np.random.default_rng(0).permutation(4 * b)[:b].astype(np.int64)
    weights = np.zeros(4 * b, dtype=np.float32)
    weights[idx] = g["batch_weights"]
    logits = torch.from_numpy(g["logits"]).requires_grad_(True)
    residuals = torch.zeros(4 * b, dtype=torch.float32)
    loss, correct = dropin.deep.weighted_cross_entropy(logits, torch.from_numpy(g["labels"]), torch.from_numpy(idx),
                                                       torch.from_numpy(weights), residuals)
    (2.0 * loss).backward()                                       # upstream gradient is honoured
    assert abs(float(loss) - float(g["loss"])) < F32_TOL * abs(float(g["loss"]))
    assert relmax(residuals.numpy()[idx], g["per_sample"]) < F32_TOL and not residuals.requires_grad
    assert np.max(np.abs(logits.grad.numpy() - 2.0 * g["dlogits"])) < 2 * F32_TOL * np.max(np.abs(g["dlogits"]))
    assert correct.dtype == torch.int32 and not correct.requires_grad


def test_train_rlvi_epoch(dropin):
    """One epoch of the drop-in `train_rlvi` against the same epoch written with the oracle's torch ops."""
    torch.manual_seed(0)
    n_train, c, bs = 512, 10, 128
    Xs, ys = torch.randn(n_train, 20), torch.randint(0, c, (n_train,))
    loader = [(Xs[i:i + bs], ys[i:i + bs], torch.arange(i, min(i + bs, n_train))) for i in range(0, n_train, bs)]

    def make():
        torch.manual_seed(1)
        m = torch.nn.Sequential(torch.nn.Linear(20, 32), torch.nn.ReLU(), torch.nn.Linear(32, c))
        return m, torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9)

    m1, o1 = make()
    res1, w1 = torch.zeros(n_train), torch.rand(n_train) * 0.5 + 0.5
    w2, res2 = w1.clone(), res1.clone()
    acc1, thr1 = dropin.deep.train_rlvi(loader, m1, o1, res1, w1, True, 0)
    m2, o2 = make()
    correct = 0.0
    for xb, yb, ib in loader:
        logits = m2(xb)
        correct += float((logits.argmax(1) == yb).float().sum() * (100.0 / yb.numel()))
        per = torch.nn.functional.cross_entropy(logits, yb, reduction="none")
        res2[ib] = per.detach()
        o2.zero_grad()
        (per * w2[ib]).mean().backward()
        o2.step()
    thr2 = deep_ref.epoch_tail(res2, w2, True, 0)
    assert isinstance(acc1, float) and abs(acc1 - correct / len(loader)) < 1e-9
    assert torch.allclose(res1, res2, rtol=1e-4, atol=1e-5) and torch.allclose(w1, w2, rtol=1e-4, atol=1e-5)
    assert abs(float(thr1) - float(thr2)) < 1e-4 and torch.equal(w1 > 0, w2 > 0)
    # overfit=False leaves the threshold object untouched (train_rlvi.py:100)
    _, thr3 = dropin.deep.train_rlvi(loader, m1, o1, res1, w1, False, 0)
    assert thr3 == 0 and isinstance(thr3, int)
