"""Round-2 additions to tests/golden/ -- again outputs of the UNMODIFIED reference, run in the build container.

TEST INFRASTRUCTURE (see oracle/__init__.py).

    python -m oracle.make_golden_r02

  config1_linreg_fixed_eps   the 100 Monte-Carlo problems of standard-learning/main.py:308-357 (epsilon = 0.2) exactly as
                             main() reaches them (module RNG seeded 0, test_mean and test_linear_regression run first),
                             with the reference's RLVI and RRM estimates for each -- BASELINE.md section 2a's row
  logreg_sklearn_n2000_d8    rlvi.logistic_regression end to end with the DEFAULT liblinear M-step (rlvi.py:92-108)
  sklearn_sep_n800_d6        utils.sklearn_log_reg on nearly separable data (the Newton M-step's hard case)
  rrm_weights_n2000          rrm.update_weights (rrm.py:12-33)
  rrm_linreg_n300_d6         rrm.linear_regression (rrm.py:55-75)
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import types
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


def load_standard_main():
    """standard-learning/main.py with matplotlib stubbed out, its sibling modules importable by bare name, run from a
    scratch directory (it creates ./plots at import).  The module RNG is seeded by the file itself (main.py:16-17)."""
    d = os.path.join(ref_shim.REF_ROOT, "standard-learning")

    class _Anything:
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __setitem__(self, k, v):
            pass

        def __getitem__(self, k):
            return _Anything()

        def __iter__(self):                      # `fig, ax = plt.subplots()`
            return iter((_Anything(), _Anything()))

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager"):
        mod = types.ModuleType(name)
        mod.__getattr__ = lambda attr, _a=_Anything(): _a        # any attribute is a no-op object
        sys.modules[name] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].font_manager = sys.modules["matplotlib.font_manager"]
    sys.path.insert(0, d)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        for k in ("utils", "rlvi", "rrm", "sever", "huber"):
            sys.modules.pop(k, None)
        spec = importlib.util.spec_from_file_location("_ref_standard_main", os.path.join(d, "main.py"))
        main = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(main)
    finally:
        os.chdir(cwd)
    return main


def config1(sizes):
    main = load_standard_main()
    warnings.simplefilter("ignore")
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    os.makedirs("plots", exist_ok=True)
    try:
        main.test_mean()                      # main.py:576-577: the draws that precede the experiment
        main.test_linear_regression()
        rec = {"X": [], "y": [], "rlvi": [], "rrm": []}
        ref_rlvi_fn, ref_rrm_fn = main.rlvi.linear_regression, main.rrm.linear_regression

        def rlvi_rec(X, y, *a, **k):
            th = ref_rlvi_fn(X, y, *a, **k)
            rec["X"].append(X.copy())
            rec["y"].append(y.copy())
            rec["rlvi"].append(th.copy())
            return th

        def rrm_rec(X, y, *a, **k):
            th = ref_rrm_fn(X, y, *a, **k)
            rec["rrm"].append(th.copy())
            return th

        main.rlvi.linear_regression, main.rrm.linear_regression = rlvi_rec, rrm_rec
        try:
            main.test_linear_regression_fixed_eps()
        finally:
            main.rlvi.linear_regression, main.rrm.linear_regression = ref_rlvi_fn, ref_rrm_fn
    finally:
        os.chdir(cwd)
    X = np.stack(rec["X"])
    y = np.stack(rec["y"])
    th_rlvi = np.stack(rec["rlvi"])
    th_rrm = np.stack(rec["rrm"])
    err = np.linalg.norm(1.0 - th_rlvi, axis=1) / np.sqrt(10.0)
    print("config 1, RLVI relative error: mean %.5f median %.5f over %d runs" % (err.mean(), np.median(err), len(err)))
    sizes["config1_linreg_fixed_eps"] = save("config1_linreg_fixed_eps", X=X, y=y, theta_rlvi=th_rlvi, theta_rrm=th_rrm)


def others(sizes):
    ref_rlvi, ref_utils = ref_shim.standard()
    warnings.simplefilter("ignore")
    # default liblinear route, end to end
    X, y, _ = synth.logistic_data(2000, 8, 0.25, seed=21)
    theta = ref_rlvi.logistic_regression(X.copy(), y.copy())
    sizes["logreg_sklearn_n2000_d8"] = save("logreg_sklearn_n2000_d8", X=X, y=y, theta=theta)
    # nearly separable M-step
    rng = np.random.default_rng(22)
    X = rng.normal(size=(800, 6))
    tstar = 6.0 * np.array([1.0, -0.5, 0.25, 0.0, 0.7, -0.3])
    y = (X @ tstar + 0.5 > 0).astype(np.float64)
    flip = rng.random(800) < 0.01
    y[flip] = 1 - y[flip]
    w = 0.2 + 0.8 * rng.random(800)
    w_in = w.copy()
    theta, losses = ref_utils.sklearn_log_reg(X.copy(), y.copy(), w_in)
    sizes["sklearn_sep_n800_d6"] = save("sklearn_sep_n800_d6", X=X, y=y, w=w, w_after=w_in, theta=theta, losses=losses)
    # RRM
    d = os.path.join(ref_shim.REF_ROOT, "standard-learning")
    rrm = ref_shim._load(os.path.join(d, "rrm.py"), "_ref_standard_rrm", {"utils": ref_utils})
    losses = synth.losses_mixture(2000, seed=23)
    sizes["rrm_weights_n2000"] = save("rrm_weights_n2000", losses=losses, eps=np.array(0.2),
                                      weights=rrm.update_weights(losses.copy(), 0.2))
    X, y = synth.linear_regression_data(300, 6, 0.2, seed=24)
    sizes["rrm_linreg_n300_d6"] = save("rrm_linreg_n300_d6", X=X, y=y, eps=np.array(0.4),
                                       theta=rrm.linear_regression(X.copy(), y.copy(), 0.4))


def main():
    sizes = {}
    others(sizes)
    config1(sizes)
    mpath = os.path.join(OUT, "MANIFEST.json")
    manifest = json.load(open(mpath))
    manifest["files_bytes"].update(sizes)
    manifest["generated_by_r02"] = "python -m oracle.make_golden_r02 (same library versions)"
    with open(mpath, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(sizes, indent=1))


if __name__ == "__main__":
    main()
