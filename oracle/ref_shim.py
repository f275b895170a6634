"""Loader for the UNMODIFIED reference modules (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  `/root/reference` exists only in the build container;
nothing that runs on the GPU box may import this module (it raises if the tree is absent).

The reference's directories are flat script collections that import siblings by bare name, and
`utils` means a different file in standard-learning/ and deep-learning/ (SURVEY.md section 1), so each
module is loaded by file path under a private name with its sibling `utils` injected only for the
duration of the load.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("RLVI_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "standard-learning"))


def _load(path, name, siblings=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    saved = {}
    for k, v in (siblings or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    try:
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


_cache: dict = {}


def standard():
    """-> (rlvi module, utils module) of /root/reference/standard-learning."""
    if not available():
        raise RuntimeError("reference tree not present (this only runs in the build container)")
    if "standard" not in _cache:
        d = os.path.join(REF_ROOT, "standard-learning")
        utils = _load(os.path.join(d, "utils.py"), "_ref_standard_utils")
        rlvi = _load(os.path.join(d, "rlvi.py"), "_ref_standard_rlvi", {"utils": utils})
        _cache["standard"] = (rlvi, utils)
    return _cache["standard"]


def deep():
    """-> methods/train_rlvi.py module of /root/reference/deep-learning."""
    if not available():
        raise RuntimeError("reference tree not present (this only runs in the build container)")
    if "deep" not in _cache:
        d = os.path.join(REF_ROOT, "deep-learning")
        utils = _load(os.path.join(d, "utils.py"), "_ref_deep_utils")
        mod = _load(os.path.join(d, "methods", "train_rlvi.py"), "_ref_deep_train_rlvi", {"utils": utils})
        _cache["deep"] = mod
    return _cache["deep"]


def online():
    """-> a namespace holding online-learning/main.py's `update_weights_rlvi` and `cross_entropy`.

    main.py executes matplotlib imports, `os.makedirs('plots')` and `loadmat('./humanactivity.mat')`
    (file absent from the tree) at import time, so only the two pure functions are compiled out of the
    source text -- unmodified -- instead of importing the module."""
    if not available():
        raise RuntimeError("reference tree not present (this only runs in the build container)")
    if "online" not in _cache:
        import ast

        path = os.path.join(REF_ROOT, "online-learning", "main.py")
        with open(path) as fh:
            src = fh.read()
        tree = ast.parse(src)
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
                and n.name in ("update_weights_rlvi", "cross_entropy")]
        ns = types.ModuleType("_ref_online_main")
        import numpy as np

        ns.np = np
        code = compile(ast.Module(body=keep, type_ignores=[]), path, "exec")
        exec(code, ns.__dict__)
        _cache["online"] = ns
    return _cache["online"]
