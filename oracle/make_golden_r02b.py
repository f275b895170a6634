"""Second round-2 batch of tests/golden/ -- outputs of the UNMODIFIED reference, run in the build container.

TEST INFRASTRUCTURE (see oracle/__init__.py).

    python -m oracle.make_golden_r02b

  sever_linreg_n600_d8   sever.linear_regression (standard-learning/sever.py:11-42), eps = 0.4 as main.py:271 calls it
  sever_pca_n500_d6      sever.pca (sever.py:82-113), eps = 0.3
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
from oracle.make_golden_r02 import OUT, save  # noqa: E402
from rlvi_b200 import synth  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    _, ref_utils = ref_shim.standard()
    d = os.path.join(ref_shim.REF_ROOT, "standard-learning")
    sever = ref_shim._load(os.path.join(d, "sever.py"), "_ref_standard_sever", {"utils": ref_utils})
    sizes = {}
    X, y = synth.linear_regression_data(600, 8, 0.2, seed=31)
    sizes["sever_linreg_n600_d8"] = save("sever_linreg_n600_d8", X=X, y=y, eps=np.array(0.4),
                                         theta=sever.linear_regression(X.copy(), y.copy(), eps=0.4))
    rng = np.random.default_rng(32)
    vdir = np.array([1.0, 0.5, -0.25, 0.0, 0.3, 0.1])
    vdir /= np.linalg.norm(vdir)
    S = 2.0 * rng.normal(size=(500, 1)) * vdir + 0.3 * rng.normal(size=(500, 6))
    out = rng.random(500) < 0.2
    S[out] = 3.0 * rng.standard_t(2.0, size=(int(out.sum()), 6))
    sizes["sever_pca_n500_d6"] = save("sever_pca_n500_d6", samples=S, eps=np.array(0.3),
                                      theta=sever.pca(S.copy(), eps=0.3))
    mpath = os.path.join(OUT, "MANIFEST.json")
    manifest = json.load(open(mpath))
    manifest["files_bytes"].update(sizes)
    manifest["generated_by_r02b"] = "python -m oracle.make_golden_r02b (same library versions)"
    with open(mpath, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(sizes, indent=1))


if __name__ == "__main__":
    main()
