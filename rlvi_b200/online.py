"""Drop-in for the RLVI pieces of the reference's `online-learning/main.py`:

    update_weights_rlvi(losses, tol=1e-3, maxiter=100)     main.py:45-58
    cross_entropy(log_proba, targets)                      main.py:84-85

As in the reference every batch starts from pi = 0.5: nothing is carried across batches (quirk Q11).
NumPy in -> NumPy out; CUDA tensors in -> CUDA tensors out.  No CPU fallback.
"""
from __future__ import annotations

from . import ops
from ._host import as_device, to_caller

__all__ = ["update_weights_rlvi", "cross_entropy"]


def update_weights_rlvi(losses, tol=1e-3, maxiter=100, *, init_weight=None, return_avg=False):
    """main.py:45-58 -- pi0 = 0.5; rho = avg/(1-avg); pi' = rho e/(1 + rho e); stop when
    ||pi' - pi|| < tol; result divided by max(pi') * n.

    Extension (off by default, the reference restarts from 0.5 on every batch -- quirk Q11): pass the previous
    batch's mean posterior as `init_weight` to carry the corruption prior across batches, and
    `return_avg=True` to get `(weights, mean posterior of this batch)` back for the next call."""
    l, was_np = as_device(losses)
    pi, res = ops.fixed_point(l, variant=ops.FP_ONLINE, tol=tol, maxiter=maxiter, pi0=init_weight)
    if return_avg:
        r = ops.read_result(res)
        return to_caller(pi, was_np), r["sum_pi"] / l.numel()
    return to_caller(pi, was_np)


def cross_entropy(log_proba, targets):
    """main.py:84-85 -- -t log_proba - (1-t) log_proba (== -log_proba whatever the label, quirk Q11);
    the same two-term expression, one kernel (rlvi_online_ce_f64)."""
    lp, was_np = as_device(log_proba)
    t, _ = as_device(targets, like=lp)
    return to_caller(ops.online_ce(lp.contiguous(), t.contiguous()).view(lp.shape), was_np)
