"""CPU-torch restatement of the reference's deep-learning RLVI pieces (FP32).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are into /root/reference/deep-learning/.
For a floating-point kernel the task keeps a plain torch fp32 reference; these functions are that
reference, written from methods/train_rlvi.py's arithmetic.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F


@torch.no_grad()
def update_sample_weights(residuals, weights, tol=1e-3, maxiter=40):
    """methods/train_rlvi.py:14-38 -- in place on BOTH arguments.

    residuals -= min(residuals); E = exp(-residuals); avg = 0.95;
    loop: rho = avg/(1-avg); pi' = rho E/(1 + rho E); err = ||pi' - weights|| (first pass: the incoming
    weights); weights <- pi'; avg = mean(weights); stop when err < tol.  Finally weights /= max(weights).
    Returns the number of passes (the reference returns None)."""
    residuals.sub_(residuals.min())
    E = torch.exp(-residuals)
    avg = 0.95
    k = 0
    for k in range(1, maxiter + 1):
        rho = avg / (1 - avg)
        new = torch.div(rho * E, 1 + rho * E)
        err = torch.norm(new - weights)
        weights[:] = new
        avg = weights.mean()
        if err < tol:
            break
    weights.div_(weights.max())
    return k


def false_negative_criterion(weights, alpha=0.05):
    """methods/train_rlvi.py:41-49 -- threshold at a fixed type-II error mass.
    beta = alpha * sum(1-w); sort w descending; c = cumsum(1 - w_sorted); idx = #{c <= beta} - 1
    (idx = -1 wraps to the smallest weight: quirk Q9); returns w_sorted[idx] (0-dim tensor)."""
    beta = torch.sum(1 - weights) * alpha
    w_sorted, _ = torch.sort(weights, dim=0, descending=True)
    mass = torch.cumsum(1 - w_sorted, dim=0)
    return w_sorted[torch.sum(mass <= beta) - 1]


def weighted_ce(logits, labels, batch_weights):
    """methods/train_rlvi.py:89-94 -- per-sample CE, the pi-weighted mean, and its gradient w.r.t. the
    logits (what autograd produces at line 96): (softmax - onehot) * pi_i / B.
    Returns (per_sample_loss [B], scalar loss, dlogits [B, C])."""
    logits = logits.detach().clone().requires_grad_(True)
    per_sample = F.cross_entropy(logits, labels, reduction="none")
    loss = (per_sample * batch_weights).mean()
    loss.backward()
    return per_sample.detach(), loss.detach(), logits.grad.detach()


def epoch_tail(residuals, weights, overfit, threshold):
    """methods/train_rlvi.py:99-103 -- E-step after the epoch, then (if overfitting) truncation.
    In place on residuals / weights; returns the new threshold."""
    update_sample_weights(residuals, weights)
    if overfit:
        threshold = max(threshold, false_negative_criterion(weights))
        weights[weights < threshold] = 0
    return threshold
