"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python binding lists the same set, the drop-in modules keep the reference's signatures,
nothing in the product imports the oracle, and the product fails loudly (no CPU fallback)."""
import ast
import inspect
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rlvi_b200.h")
LIB = os.path.join(ROOT, "rlvi_b200", "librlvi_b200.so")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rlvi_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return LIB


def test_header_matches_binding_list():
    from rlvi_b200 import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    # the version script keeps everything else private
    assert all(s.startswith("rlvi_") for s in exported), sorted(exported)


def test_library_loads_and_host_only_calls_work(built):
    from rlvi_b200 import _lib
    lib = _lib.load()
    assert lib.rlvi_version() == 100
    assert lib.rlvi_moments_out_doubles(64) == 2 + 2 * 64 + 64 * 64
    assert lib.rlvi_fp_dist_inbox_doubles(8) == 24 * 8 + 2 * 8 + 2 * 8 * 8192   # fp slots + stats tags + stats slots
    assert isinstance(lib.rlvi_last_error(), bytes)


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of the two public structs as gcc sees the header == the ctypes mirrors in _lib.py."""
    import ctypes as C
    from rlvi_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "rlvi_b200.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rlvi_fp_result), offsetof(rlvi_fp_result, eps),\n'
        '         offsetof(rlvi_fp_result, rho), offsetof(rlvi_fp_result, sum_pi), offsetof(rlvi_fp_result, err),\n'
        '         offsetof(rlvi_fp_result, iters), offsetof(rlvi_fp_result, converged));\n'
        '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rlvi_fp_dist), offsetof(rlvi_fp_dist, rank),\n'
        '         offsetof(rlvi_fp_dist, world), offsetof(rlvi_fp_dist, n_global), offsetof(rlvi_fp_dist, inbox),\n'
        '         offsetof(rlvi_fp_dist, peer_inbox), offsetof(rlvi_fp_dist, call_index));\n'
        '  printf("%d %d\\n", RLVI_IPC_HANDLE_BYTES, RLVI_DIST_STATS_CAPACITY);\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    res = [int(v) for v in out[0].split()]
    dist = [int(v) for v in out[1].split()]
    R, D = _lib.FpResult, _lib.FpDist
    assert res == [C.sizeof(R)] + [getattr(R, f).offset for f in ("eps", "rho", "sum_pi", "err", "iters", "converged")]
    assert dist == [C.sizeof(D)] + [getattr(D, f).offset for f in ("rank", "world", "n_global", "inbox", "peer_inbox",
                                                                   "call_index")]
    from rlvi_b200.dist import ShardGroup
    assert [int(v) for v in out[2].split()] == [64, ShardGroup.STATS_CAPACITY]


def test_header_is_plain_c(tmp_path):
    """The header compiles as C99 (no C++-isms, no CUDA or torch types in the signatures)."""
    src = tmp_path / "inc.c"
    src.write_text('#include "rlvi_b200.h"\nint main(void) { return rlvi_version() < 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(tmp_path / "inc.o")], check=True)


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_sass_shows_the_blackwell_paths(built):
    """UBLKCP = cp.async.bulk (TMA engine) in the Gram kernel; DMMA = FP64 tensor pipe."""
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    assert "DMMA" in sass
    assert "SYNCS" in sass          # mbarrier


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rlvi_b200 import deep, online, rlvi, utils
    x = np.random.default_rng(0).random(16)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        rlvi.update_weights(x)
    with pytest.raises(RuntimeError):
        online.update_weights_rlvi(x)
    with pytest.raises(RuntimeError):
        utils.pca(np.ones((4, 2)), np.ones(4))
    with pytest.raises(TypeError):
        deep.update_sample_weights(torch.zeros(4), torch.ones(4))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from rlvi_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rlvi_b200")
    for fn in os.listdir(pkg):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names), fn


def _oracle_import_sites(path):
    """(function name or '<module>', line) of every `import oracle...` / `from oracle... import` in a file."""
    tree = ast.parse(open(path).read())
    sites = []

    def walk(node, scope):
        for child in ast.iter_child_nodes(node):
            s = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else scope
            names = []
            if isinstance(child, ast.Import):
                names = [a.name for a in child.names]
            elif isinstance(child, ast.ImportFrom):
                names = [child.module or ""]
            if any(n.split(".")[0] == "oracle" for n in names):
                sites.append((scope, child.lineno))
            walk(child, s)

    walk(tree, "<module>")
    return sites


def test_oracle_is_only_used_as_the_checker():
    """Outside tests/ and oracle/ itself, the oracle may be imported in exactly two places: __graft_entry__.smoke()
    (the on-device check) and bench.py's CPU leg (cpu_baseline / --impl reference).  Tools, the package and the B200
    arm of the bench never touch it."""
    allowed = {("__graft_entry__.py", "smoke"), ("bench.py", "cpu_em_step_bench")}
    found = set()
    for dirpath, dirnames, filenames in os.walk(ROOT):
        rel = os.path.relpath(dirpath, ROOT)
        dirnames[:] = [d for d in dirnames if d not in (".git", "gpurun_out", "__pycache__", "baseline")
                       and not (rel == "." and d in ("tests", "oracle"))]
        for fn in filenames:
            if fn.endswith(".py"):
                path = os.path.join(dirpath, fn)
                for scope, _ in _oracle_import_sites(path):
                    found.add((os.path.relpath(path, ROOT), scope))
    assert found == allowed, found


# ---- reference signatures ---------------------------------------------------
def sig(f):
    return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]


E_ = inspect.Parameter.empty


def test_standard_signatures():
    from rlvi_b200 import rlvi, utils
    assert sig(rlvi.update_weights) == [("losses", E_), ("tol", 1e-3), ("maxiter", 100)]
    assert sig(rlvi.update_weights_constrained) == [("losses", E_), ("n_eff", E_), ("tol", 1e-3), ("maxiter", 100)]
    assert sig(rlvi.mean) == [("sample", E_), ("maxiter", 100), ("tol", 1e-3)]
    assert sig(rlvi.linear_regression) == [("X", E_), ("y", E_), ("maxiter", 100), ("tol", 1e-3)]
    assert sig(rlvi.logistic_regression)[:4] == [("X", E_), ("y", E_), ("maxiter", 100), ("tol", 1e-2)]
    assert sig(rlvi.pca) == [("sample", E_), ("maxiter", 100), ("tol", 1e-2), ("theta_init", None)]
    assert sig(rlvi.covariance) == [("sample", E_), ("eps", E_), ("maxiter", 100), ("tol", 1e-2)]
    assert sig(utils.sigmoid) == [("x", E_)]
    assert sig(utils.cross_entropy) == [("X", E_), ("theta", E_), ("y", E_)]
    assert sig(utils.clf_predict) == [("X", E_), ("theta", E_), ("augment", True)]
    assert sig(utils.mm_log_reg) == [("X", E_), ("y", E_), ("weights", E_)]
    assert sig(utils.sklearn_log_reg)[:4] == [("X", E_), ("y", E_), ("weights", E_), ("reg_coeff", 1e2)]
    assert sig(utils.pca)[:3] == [("samples", E_), ("weights", E_), ("theta", None)]   # + keyword `precision` (FP32 mode)
    assert sig(utils.covariance) == [("samples", E_), ("weights", E_), ("mean", None)]


def test_deep_and_online_signatures():
    from rlvi_b200 import deep, online
    assert sig(deep.update_sample_weights) == [("residuals", E_), ("weights", E_), ("tol", 1e-3), ("maxiter", 40)]
    assert sig(deep.false_negative_criterion) == [("weights", E_), ("alpha", 0.05)]
    assert [n for n, _ in sig(deep.train_rlvi)][:7] == ["train_loader", "model", "optimizer", "residuals", "weights",
                                                        "overfit", "threshold"]
    assert sig(deep.train_rlvi)[7:] == [("cuda_graph", None)]      # extension, off by default
    assert sig(online.update_weights_rlvi)[:3] == [("losses", E_), ("tol", 1e-3), ("maxiter", 100)]
    assert sig(online.cross_entropy) == [("log_proba", E_), ("targets", E_)]


def test_all_dropin_signatures_match_the_reference_when_present():
    """Every public function of the four reference files on the path has a drop-in with the same leading
    argument names and the same defaults (extra keyword-only options are allowed after them)."""
    import importlib
    table = {"standard-learning/rlvi.py": "rlvi_b200.rlvi", "standard-learning/utils.py": "rlvi_b200.utils",
             "deep-learning/methods/train_rlvi.py": "rlvi_b200.deep"}
    if not os.path.exists("/root/reference/standard-learning/rlvi.py"):
        pytest.skip("reference tree not present (GPU box)")
    checked = 0
    for rel, modname in table.items():
        mod = importlib.import_module(modname)
        tree = ast.parse(open(os.path.join("/root/reference", rel)).read())
        for node in tree.body:
            if not isinstance(node, ast.FunctionDef):
                continue
            ours = getattr(mod, node.name)
            ref_args = [a.arg for a in node.args.args]
            ref_defaults = [ast.literal_eval(dflt) for dflt in node.args.defaults]
            mine = sig(ours)[:len(ref_args)]
            assert [n for n, _ in mine] == ref_args, (rel, node.name)
            if ref_defaults:
                assert [dv for _, dv in mine[-len(ref_defaults):]] == ref_defaults, (rel, node.name)
            checked += 1
    # online-learning/main.py: only the two RLVI functions are on the path
    from rlvi_b200 import online
    tree = ast.parse(open("/root/reference/online-learning/main.py").read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("update_weights_rlvi", "cross_entropy"):
            ours = getattr(online, node.name)
            ref_args = [a.arg for a in node.args.args]
            assert [n for n, _ in sig(ours)][:len(ref_args)] == ref_args
            checked += 1
    assert checked >= 18


def test_signatures_match_the_reference_when_present():
    ref = "/root/reference/standard-learning/rlvi.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present (GPU box)")
    from rlvi_b200 import rlvi
    tree = ast.parse(open(ref).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            ours = getattr(rlvi, node.name)
            ref_args = [a.arg for a in node.args.args]
            assert [n for n, _ in sig(ours)][:len(ref_args)] == ref_args, node.name
