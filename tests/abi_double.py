"""TEST DOUBLE of the `rlvi_b200.ops` layer for the CPU tier -- not a fallback, never imported by the package.

The drop-in modules (`rlvi_b200/rlvi.py`, `utils.py`, `deep.py`, `online.py`) contain host logic of their own:
outer EM loops and their stop tests, SciPy's Brent around a device sum, the d x d algebra, sklearn's sign
rule, in-place quirks, NumPy-in/NumPy-out conversion.  `pytest -m "not gpu"` has no device to run the kernels
on, so `install(monkeypatch)` swaps every `ops.*` entry point for the plain torch-CPU function below that
follows the CONTRACT written in include/rlvi_b200.h (argument meaning, in-place effects, result struct), and
lets NumPy inputs stay on the CPU.  tests/test_dropin_host_logic.py then runs the drop-ins against the golden
vectors of the unmodified reference.  The real kernels are compared with the same goldens in
tests/test_gpu_parity.py (`-m gpu`).
"""
from __future__ import annotations

import numpy as np
import torch

from rlvi_b200 import _host, ops

_RES = np.dtype([("eps", "<f8"), ("rho", "<f8"), ("sum_pi", "<f8"), ("err", "<f8"), ("iters", "<i4"),
                 ("converged", "<i4")])


def _pack(result, eps, rho, sum_pi, err, iters, converged):
    rec = np.zeros(1, dtype=_RES)
    rec[0] = (eps, rho, sum_pi, err, iters, converged)
    packed = torch.from_numpy(rec.view(np.float64).copy())
    if result is None:
        return packed
    result.copy_(packed)
    return result


def fixed_point(losses=None, *, e_work=None, scale=None, variant=ops.FP_STANDARD, tol=1e-3, maxiter=100, out=None,
                result=None, dist=None, pi0=None):
    """rlvi_fixed_point_f64 / rlvi_fixed_point_init_f64 as include/rlvi_b200.h describes them."""
    assert dist is None and maxiter >= 1
    if losses is not None:
        s = 1.0 if scale is None else float(scale)
        e = torch.exp(-s * losses)
        if e_work is not None:
            e_work.copy_(e)
    else:
        e = e_work.clone()
    n = e.numel()
    start = pi0 if pi0 is not None else (0.95 if variant == ops.FP_STANDARD else 0.5)
    pi = torch.full_like(e, start)
    new = pi
    eps = rho = err = float("nan")
    converged, k = 0, 0
    for k in range(1, maxiter + 1):
        if variant == ops.FP_STANDARD:                    # rlvi.py:13-19
            eps = 1 - float(pi.mean())
            rho = eps / (1 - eps)
            new = e / (rho + e)
        else:                                             # online-learning/main.py:50-55
            avg = float(pi.mean())
            eps, rho = 1 - avg, avg / (1 - avg)
            new = rho * e / (1 + rho * e)
        err = float(torch.linalg.norm(new - pi))
        if err < tol:
            converged = 1
            if variant == ops.FP_STANDARD:
                pi = new
            break
        pi = new
    sum_pi = float(new.sum())
    if variant == ops.FP_ONLINE:
        new = new / (new.max() * n)                       # main.py:57
    if out is None:
        out = torch.empty_like(e)
    out.copy_(new)
    return out, _pack(result, eps, rho, sum_pi, err, k, converged)


def fixed_point_deep(residuals, weights, *, e_work=None, tol=1e-3, maxiter=40, result=None, dist=None):
    """rlvi_fixed_point_deep_f32: train_rlvi.py:14-38, in place on both tensors (FP32)."""
    assert dist is None
    residuals -= residuals.min()
    e = torch.exp(-residuals)
    avg = 0.95
    err, k, converged = float("nan"), 0, 0
    for k in range(1, maxiter + 1):
        ratio = avg / (1 - avg)
        new = ratio * e / (1 + ratio * e)
        err = float(torch.linalg.norm(new - weights))
        weights[:] = new
        avg = float(weights.mean())
        if err < tol:
            converged = 1
            break
    sum_pi = float(weights.double().sum())
    weights /= weights.max()
    return _pack(result, 1 - avg, avg / (1 - avg), sum_pi, err, k, converged)


def shift_sum(losses, shift, c, *, pi_out=None, out=None):
    t = torch.exp(-losses + shift)
    r = t / (c + t)
    if pi_out is not None:
        pi_out.copy_(r)
    out = torch.empty(1, dtype=torch.float64) if out is None else out
    out[0] = r.sum()
    return out


def shift_sum_e(e, scale_t, c, *, pi_out=None, out=None):
    t = e * scale_t
    r = t / (c + t)
    if pi_out is not None:
        pi_out.copy_(r)
    out = torch.empty(1, dtype=torch.float64) if out is None else out
    out[0] = r.sum()
    return out


def loss(kind, X, params, *, y=None, intercept=False, weights=None, want_losses=True, want_e=False, losses_out=None,
         e_out=None, wsum_out=None):
    X = X.double()                                        # float32 X = the FP32-stored mode: FP64 arithmetic on it
    d = X.shape[1]
    if kind in (ops.LOSS_LOGISTIC_CE, ops.LOSS_SOFTPLUS, ops.LOSS_SQRES):
        b, theta = (params[0], params[1:]) if intercept else (0.0, params)
        phi = X @ theta + b
        if kind == ops.LOSS_LOGISTIC_CE:
            l = -y * phi + phi + torch.log1p(torch.exp(-phi))
        elif kind == ops.LOSS_SOFTPLUS:
            l = torch.logaddexp(torch.zeros_like(phi), phi)
        else:
            l = (y - phi) ** 2
    elif kind == ops.LOSS_SQDIST:
        l = ((params - X) ** 2).sum(dim=1)
    elif kind == ops.LOSS_PCA:
        l = (X * X).sum(dim=1) - (X @ params) ** 2
    else:                                                 # GAUSSIAN: params = [c, mu, U upper, U^T U = cov^-1]
        c, mu, U = params[0], params[1:1 + d], torch.triu(params[1 + d:].view(d, d))
        z = (X - mu) @ U.T
        l = 0.5 * ((z * z).sum(dim=1) + c)
    if want_losses or losses_out is not None:
        losses_out = l.clone() if losses_out is None else losses_out.copy_(l)
    if want_e or e_out is not None:
        e_out = torch.exp(-l) if e_out is None else e_out.copy_(torch.exp(-l))
    if weights is not None:
        wsum = torch.stack([(weights * l).sum(), weights.sum()])
        wsum_out = wsum if wsum_out is None else wsum_out.copy_(wsum)
    return losses_out, e_out, wsum_out


def weighted_moments(X, weights, *, y=None, power=1, want_gram=True, out=None, center=None, precision=0):
    X = X.double()
    n, d = X.shape
    if out is None:
        out = torch.zeros(2 + 2 * d + d * d, dtype=torch.float64)
    Xc = X if center is None else X - center
    w = weights if power == 1 else weights * weights
    out[0] = w.sum()
    out[1] = (w * y).sum() if y is not None else 0.0
    out[2:2 + d] = Xc.T @ weights                        # first power even when power == 2
    out[2 + d:2 + 2 * d] = Xc.T @ (w * y) if y is not None else 0.0
    if want_gram:
        out[2 + 2 * d:] = ((Xc * w[:, None]).T @ Xc).reshape(-1)
    return out


def logistic_grad(X, y, weights, params, *, out=None):
    phi = X @ params[1:] + params[0]
    c = weights * (torch.sigmoid(phi) - y)
    g = torch.cat([c.sum().reshape(1), X.T @ c])
    return g if out is None else out.copy_(g)


def sigmoid(x, *, out=None):
    z = torch.exp(-torch.abs(x))
    r = torch.where(x >= 0, torch.ones_like(z), z) / (1 + z)
    return r if out is None else out.copy_(r)


def online_ce(log_proba, targets, *, out=None):
    r = -targets * log_proba - (1 - targets) * log_proba
    return r if out is None else out.copy_(r)


def irls_weights(e, weights, *, out=None):
    r = weights * e * (1 - e)
    return r if out is None else out.copy_(r)


def rrm_sum(losses, inv_alpha, cutoff=1e-16, *, norm=0.0, w_out=None, out=None):
    phi = torch.exp(-losses * inv_alpha)
    if w_out is not None:
        w_out.copy_(phi * norm)
    out = torch.empty(1, dtype=torch.float64) if out is None else out
    out[0] = torch.clamp(phi, min=cutoff).sum()
    return out


def sever_pass(X, u, op, scalar, *, a=None, b=None, active=None, out0=None, out1=None, out2=None):
    """rlvi_sever_pass_f64 as include/rlvi_b200.h describes it."""
    dot = X @ u
    act = torch.ones_like(dot) if active is None else active
    if op == 0:
        c = scalar * (dot - (b if b is not None else 0.0))
        o0, o1 = c, act * c * c
        o2 = torch.where(c != 0, 1.0 / c, torch.zeros_like(c))
    else:
        t = a * dot - scalar
        o0, o1, o2 = torch.where(act != 0, t * t, torch.full_like(t, -1.0)), None, None
    res = []
    for dst, val in ((out0, o0), (out1, o1), (out2, o2)):
        if val is None:
            res.append(dst)
        elif dst is None:
            res.append(val.clone())
        else:
            dst.copy_(val)
            res.append(dst)
    return tuple(res)


def wce_fwd_bwd(logits, labels, weights, residuals, *, indexes=None, want_grad=True, want_per_sample=False,
                want_correct=False):
    b = logits.shape[0]
    idx = torch.arange(b) if indexes is None else indexes
    lsm = torch.log_softmax(logits.detach(), dim=1)
    per = -lsm[torch.arange(b), labels]
    residuals[idx] = per
    w = weights[idx]
    dlogits = None
    if want_grad:
        dlogits = torch.exp(lsm)
        dlogits[torch.arange(b), labels] -= 1.0
        dlogits *= (w / b)[:, None]
    correct = None
    if want_correct:
        top = logits.detach().topk(min(5, logits.shape[1]), dim=1).indices
        correct = torch.tensor([int((top[:, 0] == labels).sum()), int((top == labels[:, None]).any(dim=1).sum())],
                               dtype=torch.int32)
    return {"loss": (per * w).mean().reshape(1), "dlogits": dlogits, "per_sample": per if want_per_sample else None,
            "correct": correct}


def fn_threshold(weights, alpha=0.05, prev_threshold=0.0, truncate=False, *, out=None):
    ws, _ = torch.sort(weights, descending=True)
    mass = torch.cumsum((1 - ws).double(), dim=0)
    beta = alpha * float((1 - weights).double().sum())
    k = int((mass <= beta).sum()) - 1                     # -1 wraps to the smallest weight (quirk Q9)
    thr = max(float(prev_threshold), float(ws[k]))
    thr32 = torch.tensor([thr], dtype=torch.float32)
    if truncate:
        weights[weights < thr32] = 0
    return thr32 if out is None else out.copy_(thr32)


def _as_device(a, like=None, dtype=torch.float64):
    """_host.as_device without the GPU: same return convention, tensors stay on the CPU."""
    if isinstance(a, torch.Tensor):
        return a.to(dtype).contiguous(), False
    np_dtype = {torch.float64: np.float64, torch.float32: np.float32, torch.int64: np.int64}[dtype]
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype)), True


def install(monkeypatch):
    """Swap the ops entry points and the host<->device glue of the drop-in modules for the doubles above."""
    from rlvi_b200 import deep, online, rlvi, rrm, sever, utils
    for name in ("fixed_point", "fixed_point_deep", "shift_sum", "shift_sum_e", "loss", "weighted_moments",
                 "logistic_grad", "wce_fwd_bwd", "fn_threshold", "sigmoid", "online_ce", "irls_weights", "rrm_sum",
                 "sever_pass"):
        monkeypatch.setattr(ops, name, globals()[name])
    for mod in (rlvi, utils, online, rrm, sever):
        monkeypatch.setattr(mod, "as_device", _as_device)
    for mod in (rlvi, utils):
        monkeypatch.setattr(mod, "as_device_x", lambda a, like=None: _as_device(
            a, like, torch.float32 if getattr(a, "dtype", None) in (torch.float32, np.float32, np.dtype("float32")) else torch.float64))
    monkeypatch.setattr(_host, "as_device", _as_device)
    return deep
