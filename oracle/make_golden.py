"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).

    python -m oracle.make_golden            # rewrites tests/golden/

The reference ships no golden vectors, so the fixtures are outputs of the
reference's own functions, imported from /root/reference through oracle/ref_shim.py, on seeded
inputs from rlvi_b200/synth.py.  Inputs are stored next to the outputs, so the fixtures do not
depend on NumPy's RNG stream staying stable.  Library versions used are recorded in
tests/golden/MANIFEST.json (the reference pins older ones: requirements.txt:1-7).
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from rlvi_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    return os.path.getsize(path)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_rlvi, ref_utils = ref_shim.standard()
    ref_deep = ref_shim.deep()
    ref_online = ref_shim.online()
    sizes = {}
    warnings.simplefilter("ignore")

    # ---- a1: rlvi.update_weights -------------------------------------------------------------
    for tag, n, seed in (("n40", 40, 1), ("n1000", 1000, 2), ("n4096", 4096, 3), ("n16384", 16384, 4)):
        losses = synth.losses_mixture(n, seed=seed)
        pi = ref_rlvi.update_weights(losses.copy())
        sizes["update_weights_" + tag] = save("update_weights_" + tag, losses=losses, pi=pi)
    # non-default tol / maxiter, and a loss vector with negative entries (Gaussian log-density)
    losses = synth.losses_mixture(777, seed=5) - 1.5
    pi = ref_rlvi.update_weights(losses.copy(), tol=1e-5, maxiter=7)
    sizes["update_weights_neg"] = save("update_weights_neg", losses=losses, pi=pi,
                                       tol=np.float64(1e-5), maxiter=np.int64(7))

    # ---- a2: rlvi.update_weights_constrained ---------------------------------------------------
    for tag, n, seed, eps in (("n50", 50, 6, 0.2), ("n3000", 3000, 7, 0.4)):
        losses = synth.losses_mixture(n, seed=seed) - 1.0
        n_eff = n * (1 - eps)
        pi = ref_rlvi.update_weights_constrained(losses.copy(), n_eff)
        sizes["constrained_" + tag] = save("constrained_" + tag, losses=losses, n_eff=np.float64(n_eff), pi=pi)

    # ---- a3: rlvi.mean ---------------------------------------------------------------------------
    for tag, n, d, seed in (("n100_d2", 100, 2, 8), ("n768_d64", 768, 64, 9)):
        X = synth.mean_data(n, d, 0.2, seed)
        theta = ref_rlvi.mean(X.copy())
        sizes["mean_" + tag] = save("mean_" + tag, X=X, theta=theta)

    # ---- a4: rlvi.linear_regression (config 1 shape + a d=64 case) -------------------------------
    for tag, n, d, seed in (("n40_d10", 40, 10, 10), ("n768_d64", 768, 64, 11)):
        X, y = synth.linear_regression_data(n, d, 0.2, 2.5, seed)
        theta = ref_rlvi.linear_regression(X.copy(), y.copy())
        sizes["linreg_" + tag] = save("linreg_" + tag, X=X, y=y, theta=theta)

    # ---- a6/a7: utils.sigmoid, utils.cross_entropy, utils.mm_log_reg ------------------------------
    X, y, _ = synth.logistic_data(768, 64, 0.3, seed=12)
    w = ref_rlvi.update_weights(synth.losses_mixture(768, seed=13))
    theta, losses = ref_utils.mm_log_reg(X.copy(), y.copy(), w.copy())
    Xa = np.hstack([np.ones((X.shape[0], 1)), X])
    x_sig = np.linspace(-800, 800, 4001)
    sizes["mm_log_reg_n768_d64"] = save(
        "mm_log_reg_n768_d64", X=X, y=y, w=w, theta=theta, losses=losses,
        ce_at_theta=ref_utils.cross_entropy(Xa, theta, y), x_sig=x_sig, sig=ref_utils.sigmoid(x_sig))

    # ---- a5: rlvi.logistic_regression with the reference's commented MM alternative ---------------
    # (rlvi.py:95,102).  The default liblinear M-step (utils.py:61-73) is third-party C++ and stays
    # out of scope; the alternative is enabled by rebinding the name rlvi.py calls.
    X, y, _ = synth.logistic_data(1500, 8, 0.2, seed=14)
    saved = ref_utils.sklearn_log_reg
    ref_utils.sklearn_log_reg = ref_utils.mm_log_reg
    try:
        theta = ref_rlvi.logistic_regression(X.copy(), y.copy())
    finally:
        ref_utils.sklearn_log_reg = saved
    sizes["logreg_mm_n1500_d8"] = save("logreg_mm_n1500_d8", X=X, y=y, theta=theta)

    # ---- a8: the loss sklearn_log_reg reports (softplus of the fitted scores) + weight mutation ---
    X, y, _ = synth.logistic_data(600, 5, 0.3, seed=15)
    w = np.linspace(0.1, 0.7, 600)
    w_in = w.copy()
    theta, losses = ref_utils.sklearn_log_reg(X.copy(), y.copy(), w_in)
    sizes["sklearn_loss_n600_d5"] = save("sklearn_loss_n600_d5", X=X, y=y, w=w, w_after=w_in,
                                         theta=theta, losses=losses)

    # ---- a9/a11: utils.pca, rlvi.pca ---------------------------------------------------------------
    for tag, n, d, seed in (("n400_d2", 400, 2, 16), ("n768_d64", 768, 64, 17)):
        X, _ = synth.pca_data(n, d, 0.2, seed)
        w = ref_rlvi.update_weights(synth.losses_mixture(n, seed=seed + 100))
        th1, l1 = ref_utils.pca(X.copy(), w.copy())
        init = np.ones(d)
        init[0] = 0.1
        init /= np.linalg.norm(init)
        th0, l0 = ref_utils.pca(X.copy(), np.ones(n), init.copy())
        theta = ref_rlvi.pca(X.copy(), theta_init=init.copy()) if n <= 2048 else th1
        sizes["pca_" + tag] = save("pca_" + tag, X=X, w=w, theta_mstep=th1, losses_mstep=l1,
                                   theta_init=init, losses_init=l0, theta=theta)

    # ---- a10/a11: utils.covariance, rlvi.covariance ------------------------------------------------
    for tag, n, d, seed, scale in (("n50_d2", 50, 2, 18, 1.0), ("n2048_d16", 2048, 16, 19, 0.25)):
        X, _ = synth.covariance_data(n, d, 0.2, seed, scale)
        w = ref_rlvi.update_weights(synth.losses_mixture(n, seed=seed + 100))
        cov1, l1 = ref_utils.covariance(X.copy(), w.copy())
        cov = ref_rlvi.covariance(X.copy(), eps=0.2)
        sizes["cov_" + tag] = save("cov_" + tag, X=X, w=w, cov_mstep=cov1, losses_mstep=l1, cov=cov,
                                   eps=np.float64(0.2))

    # ---- a15: online update_weights_rlvi + its "cross-entropy" -------------------------------------
    rng = np.random.default_rng(20)
    logp = np.log(rng.uniform(0.02, 0.98, size=100))
    tgt = (rng.random(100) < 0.5).astype(np.float64)
    res = ref_online.cross_entropy(logp, tgt)
    pi = ref_online.update_weights_rlvi(res.copy())
    sizes["online_n100"] = save("online_n100", log_proba=logp, targets=tgt, residuals=res, pi=pi)

    # ---- a12/a13/a14: deep path (CPU torch, FP32) ----------------------------------------------------
    import torch

    torch.manual_seed(1)
    n_train = 45000
    logits, labels = synth.deep_batch(512, 100, seed=1)
    lt, lb = torch.from_numpy(logits), torch.from_numpy(labels)
    lt_g = lt.clone().requires_grad_(True)
    per = torch.nn.functional.cross_entropy(lt_g, lb, reduction="none")       # train_rlvi.py:89
    bw = torch.rand(512)
    loss = (per * bw).mean()                                                     # train_rlvi.py:92-94
    loss.backward()
    sizes["deep_wce_b512_c100"] = save(
        "deep_wce_b512_c100", logits=logits, labels=labels, batch_weights=bw.numpy(),
        per_sample=per.detach().numpy(), loss=loss.detach().numpy(), dlogits=lt_g.grad.numpy())

    # residuals as an epoch would leave them: CE of random logits, clean ~ small, flipped ~ large
    rng = np.random.default_rng(21)
    residuals = np.abs(rng.normal(0.3, 0.3, size=n_train)).astype(np.float32)
    bad = rng.random(n_train) < 0.45
    residuals[bad] += rng.gamma(4.0, 1.0, size=int(bad.sum())).astype(np.float32)
    res_t = torch.from_numpy(residuals.copy())
    w_t = torch.ones(n_train)
    ref_deep.update_sample_weights(res_t, w_t)
    thr = ref_deep.false_negative_criterion(w_t)
    w_trunc = w_t.clone()
    w_trunc[w_trunc < thr] = 0                                                   # train_rlvi.py:103
    # second epoch: weights carried over (first-pass error is measured against them)
    res2 = np.maximum(residuals + rng.normal(0, 0.05, size=n_train).astype(np.float32), 0).astype(np.float32)
    res2_t = torch.from_numpy(res2.copy())
    w2_t = w_trunc.clone()
    ref_deep.update_sample_weights(res2_t, w2_t)
    thr2 = ref_deep.false_negative_criterion(w2_t)
    sizes["deep_estep_n45000"] = save(
        "deep_estep_n45000", residuals=residuals, residuals_after=res_t.numpy(), weights=w_t.numpy(),
        threshold=thr.numpy(), weights_truncated=w_trunc.numpy(), residuals2=res2,
        residuals2_after=res2_t.numpy(), weights2=w2_t.numpy(), threshold2=thr2.numpy())
    # Q9: nothing fits under beta -> index -1 wraps to the smallest weight
    w_small = torch.tensor([0.5, 0.4, 0.3, 0.45], dtype=torch.float32)
    sizes["deep_threshold_wrap"] = save("deep_threshold_wrap", weights=w_small.numpy(),
                                        threshold=ref_deep.false_negative_criterion(w_small).numpy())

    import scipy
    import sklearn

    manifest = {
        "generated_by": "python -m oracle.make_golden (imports /root/reference through oracle/ref_shim.py)",
        "versions": {"numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__,
                     "torch": torch.__version__},
        "reference_pins": "requirements.txt:1-7 (numpy 1.26.0, scipy 1.12.0, scikit-learn 1.5.0, torch 2.1.2)",
        "files_bytes": sizes,
    }
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(sizes, indent=1), "\ntotal", sum(sizes.values()))


if __name__ == "__main__":
    main()
