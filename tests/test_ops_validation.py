"""Host logic of rlvi_b200.ops without a GPU: the argument checks that guard the raw-pointer C ABI.

The library handle is replaced by a recorder (no compute happens, nothing is compared numerically): the
tests assert that well-formed calls -- the exact shapes bench.py, smoke() and the drop-in modules use --
reach the C entry point with the right scalar arguments, and that every length / dtype / layout mismatch
is rejected on the host before a pointer is handed over.
"""
import ctypes as C

import pytest
import torch

from rlvi_b200 import _lib, ops


class _Recorder:
    def __init__(self):
        self.calls = []

    def rlvi_moments_out_doubles(self, d):
        return 2 + 2 * d + d * d

    def __getattr__(self, name):
        if not name.startswith("rlvi_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append((name, args))
            return 0
        return fn


class _Ctx:
    def __init__(self):
        self.lib = _Recorder()
        self.handle = C.c_void_p(1)


@pytest.fixture
def rec(monkeypatch):
    ctx = _Ctx()

    def chk(t, dtype, name):            # ops._chk minus the is_cuda requirement
        if t is None:
            return None
        if t.dtype != dtype or not t.is_contiguous():
            raise TypeError(f"{name} must be a contiguous tensor of dtype {dtype}")
        return t

    monkeypatch.setattr(ops, "_dev", lambda t: 0)
    monkeypatch.setattr(ops, "_stream", lambda dev: None)
    monkeypatch.setattr(ops, "_ctx", lambda dev: ctx)
    monkeypatch.setattr(ops, "_chk", chk)
    return ctx.lib


def f64(*shape):
    return torch.zeros(shape, dtype=torch.float64)


def f32(*shape):
    return torch.zeros(shape, dtype=torch.float32)


def test_cpu_tensors_are_refused_without_patching():
    with pytest.raises(TypeError, match="CUDA tensors only"):
        ops.fixed_point(f64(8))
    with pytest.raises(TypeError, match="CUDA tensors only"):
        ops.loss(ops.LOSS_PCA, f64(8, 4), f64(4))


def test_step_calls_reach_the_abi(rec):
    n, d = 64, 16
    X, y, params = f64(n, d), f64(n), f64(d + 1)
    _, e, _ = ops.loss(ops.LOSS_LOGISTIC_CE, X, params, y=y, intercept=True, want_losses=False, want_e=True)
    pi, res = ops.fixed_point(None, e_work=e)
    mom = ops.weighted_moments(X, pi)
    assert e.shape == (n,) and pi.shape == (n,) and res.shape == (5,) and mom.numel() == 2 + 2 * d + d * d
    names = [c[0] for c in rec.calls]
    assert names == ["rlvi_loss_f64", "rlvi_fixed_point_f64", "rlvi_weighted_moments_f64"]
    loss_args = rec.calls[0][1]
    assert loss_args[1:3] == (ops.LOSS_LOGISTIC_CE, 1) and loss_args[5:7] == (n, d)
    fp_args = rec.calls[1][1]
    assert fp_args[1] == ops.FP_STANDARD and fp_args[5] == n and fp_args[6] == 1e-3 and fp_args[7] == 100
    assert fp_args[2] is None                                  # losses == NULL: e_work is the input
    m_args = rec.calls[2][1]
    assert m_args[4:8] == (n, d, 1, 1)
    m = ops.split_moments(mom, d)
    assert m["S1"].shape == (d,) and m["Sy"].shape == (d,) and m["G"].shape == (d, d) and m["S0"].dim() == 0


def test_every_entry_point_accepts_its_documented_shapes(rec):
    n, d = 32, 8
    X, v = f64(n, d), f64(n)
    ops.fixed_point(v, scale=f64(1), pi0=0.7, variant=ops.FP_ONLINE)
    assert rec.calls[-1][0] == "rlvi_fixed_point_init_f64" and rec.calls[-1][1][8] == 0.7
    ops.fixed_point_deep(f32(n), f32(n))
    ops.shift_sum(v, 0.5, 0.25, pi_out=f64(n))
    ops.shift_sum_e(v, 1.5, 0.25)
    for kind, npar, icpt in ((ops.LOSS_LOGISTIC_CE, d, False), (ops.LOSS_SOFTPLUS, d + 1, True),
                             (ops.LOSS_SQRES, d, False), (ops.LOSS_SQDIST, d, False), (ops.LOSS_PCA, d, False),
                             (ops.LOSS_GAUSSIAN, 1 + d + d * d, False)):
        ops.loss(kind, X, f64(npar), y=v, intercept=icpt, weights=v)
    ops.weighted_moments(X, v, y=v, power=2, want_gram=False, center=f64(d))
    assert rec.calls[-1][0] == "rlvi_weighted_moments_centered_f64"
    ops.logistic_grad(X, v, v, f64(d + 1))
    idx = torch.zeros(4, dtype=torch.int64)
    out = ops.wce_fwd_bwd(f32(4, 10), idx, f32(100), f32(100), indexes=idx, want_per_sample=True, want_correct=True)
    assert out["dlogits"].shape == (4, 10) and out["per_sample"].shape == (4,) and out["correct"].shape == (2,)
    ops.fn_threshold(f32(n), alpha=0.1, prev_threshold=0.2, truncate=True)
    assert rec.calls[-1][1][3:6] == (pytest.approx(0.1), pytest.approx(0.2), 1)


@pytest.mark.parametrize("call, msg", [
    (lambda: ops.fixed_point(f64(8), out=f64(7)), "out must have 8"),
    (lambda: ops.fixed_point(f64(8), e_work=f64(9)), "e_work must have 8"),
    (lambda: ops.fixed_point(f64(8), scale=f64(2)), "scale must have 1"),
    (lambda: ops.fixed_point(f64(8), result=f64(4)), "result must have 5"),
    (lambda: ops.fixed_point(f64(0)), "empty"),
    (lambda: ops.fixed_point_deep(f32(8), f32(7)), "same length"),
    (lambda: ops.shift_sum(f64(8), 0.0, 1.0, pi_out=f64(4)), "pi_out must have 8"),
    (lambda: ops.shift_sum_e(f64(8), 1.0, 1.0, out=f64(2)), "out must have 1"),
    (lambda: ops.loss(ops.LOSS_PCA, f64(8), f64(8)), r"X must be \[n, d\]"),
    (lambda: ops.loss(ops.LOSS_PCA, f64(8, 4), f64(5)), "params must have 4"),
    (lambda: ops.loss(ops.LOSS_SQRES, f64(8, 4), f64(5), y=f64(8), intercept=True, losses_out=f64(6)),
     "losses_out must have 8"),
    (lambda: ops.loss(ops.LOSS_LOGISTIC_CE, f64(8, 4), f64(4), y=f64(9)), "y must have 8"),
    (lambda: ops.loss(ops.LOSS_GAUSSIAN, f64(8, 4), f64(1 + 4 + 16), intercept=True), "no intercept"),
    (lambda: ops.loss(ops.LOSS_SQDIST, f64(8, 4), f64(4), weights=f64(8), wsum_out=f64(1)), "wsum_out must have 2"),
    (lambda: ops.weighted_moments(f64(8, 4), f64(7)), "weights must have 8"),
    (lambda: ops.weighted_moments(f64(8, 4), f64(8), y=f64(3)), "y must have 8"),
    (lambda: ops.weighted_moments(f64(8, 4), f64(8), out=f64(10)), "out must have 26"),
    (lambda: ops.weighted_moments(f64(8, 4), f64(8), center=f64(5)), "center must have d"),
    (lambda: ops.logistic_grad(f64(8, 4), f64(8), f64(8), f64(4)), "params must have 5"),
    (lambda: ops.wce_fwd_bwd(f32(4, 10), torch.zeros(5, dtype=torch.int64), f32(9), f32(9)), "labels must have 4"),
    (lambda: ops.wce_fwd_bwd(f32(4, 10), torch.zeros(4, dtype=torch.int64), f32(9), f32(8)), "same length"),
    (lambda: ops.wce_fwd_bwd(f32(4, 10), torch.zeros(4, dtype=torch.int64), f32(3), f32(3)), "n_train >= batch"),
    (lambda: ops.fn_threshold(f32(8), out=f32(2)), "out must have 1"),
])
def test_length_mismatches_never_reach_the_abi(rec, call, msg):
    with pytest.raises(ValueError, match=msg):
        call()
    assert rec.calls == []


@pytest.mark.parametrize("call", [
    lambda: ops.fixed_point(f32(8)),                                       # FP64 entry point
    lambda: ops.fixed_point_deep(f64(8), f64(8)),                          # FP32 entry point
    lambda: ops.loss(ops.LOSS_PCA, f64(4, 8).t(), f64(4)),                 # non-contiguous X
    lambda: ops.weighted_moments(f64(8, 4), f32(8)),
    lambda: ops.wce_fwd_bwd(f32(4, 10), torch.zeros(4, dtype=torch.int32), f32(9), f32(9)),
])
def test_dtype_and_layout_mismatches_never_reach_the_abi(rec, call):
    with pytest.raises(TypeError):
        call()
    assert rec.calls == []


def test_contexts_are_per_device_and_stream(monkeypatch):
    made = []

    class FakeContext:
        def __init__(self, device):
            made.append(device)
            self.launches = 3

    monkeypatch.setattr(_lib, "Context", FakeContext)
    monkeypatch.setattr(_lib, "_contexts", {})
    a = _lib.context(0)
    assert _lib.context(0, 0) is a and _lib.context(0, None) is a
    b = _lib.context(0, 0x7f00)
    c = _lib.context(1)
    assert b is not a and c is not a and _lib.context(0, 0x7f00) is b
    assert made == [0, 0, 1]
    assert _lib.launches(0) == 6 and _lib.launches(1) == 3 and _lib.launches(2) == 0


def test_first_context_call_loads_the_library_without_deadlock():
    """context() holds the module lock while it builds a Context, whose constructor calls load(): with a
    cold module (library not loaded yet) that must not self-deadlock.  Runs in a child process with a
    timeout; without a GPU the expected outcome is the loud rlvi_ctx_create failure, not a hang."""
    import subprocess
    import sys
    code = ("import rlvi_b200._lib as L\n"
            "assert L._lib is None\n"
            "try:\n"
            "    L.context(0)\n"
            "    print('created')\n"
            "except RuntimeError as e:\n"
            "    print('refused:', e)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "created" in r.stdout or "refused" in r.stdout


def test_fp32_mode_calls_reach_the_f32_entry_points(rec):
    """float32 X selects rlvi_loss_f32 / rlvi_weighted_moments_f32 (per-sample vectors stay float64)."""
    n, d = 64, 16
    X = f32(n, d)
    ops.loss(ops.LOSS_PCA, X, f64(d), want_e=True)
    ops.weighted_moments(X, f64(n), power=2, precision=ops.TF32X1)
    names = [c[0] for c in rec.calls]
    assert names == ["rlvi_loss_f32", "rlvi_weighted_moments_f32"]
    args = rec.calls[1][1]
    assert args[4:9] == (n, d, 2, 1, ops.TF32X1)
    with pytest.raises(ValueError):
        ops.loss(ops.LOSS_GAUSSIAN, X, f64(1 + d + d * d))
    with pytest.raises(ValueError):
        ops.weighted_moments(X, f64(n), center=f64(d))
    with pytest.raises(TypeError):
        ops.weighted_moments(X, f32(n))          # weights must be float64


def test_host_entry_point_checks_its_output_buffers():
    """rlvi_em_step_logistic_host copies n doubles of pi and the statistics into raw host pointers: a short, float32
    or strided buffer must be refused before the call."""
    import numpy as np
    X, y, params = np.zeros((8, 4)), np.zeros(8), np.zeros(5)
    for bad in (np.zeros(7), np.zeros(8, dtype=np.float32), np.zeros(16)[::2]):
        with pytest.raises(ValueError, match="pi_out"):
            ops.em_step_logistic_host(X, y, params, pi_out=bad)
    with pytest.raises(ValueError, match="moments_out"):
        ops.em_step_logistic_host(X, y, params, moments_out=np.zeros(3))
