"""Drop-in for the reference's `deep-learning/methods/train_rlvi.py` (FP32, CUDA tensors, in place):

    update_sample_weights(residuals, weights, tol=1e-3, maxiter=40) -> None      train_rlvi.py:14-38
    false_negative_criterion(weights, alpha=0.05) -> 0-dim tensor               train_rlvi.py:41-49
    train_rlvi(train_loader, model, optimizer, residuals, weights, overfit, threshold)
        -> (train_acc, threshold)                                                train_rlvi.py:52-106

plus `weighted_cross_entropy(logits, labels, indexes, weights, residuals)`, the fused autograd op that
replaces lines 85-94 (accuracy + per-sample CE + scatter + gather + weighted mean) and their backward
with ONE kernel launch.  `residuals` / `weights` are the caller-owned `[N_train]` FP32 CUDA tensors of
deep-learning/main.py:249-251.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["train_rlvi", "update_sample_weights", "false_negative_criterion", "weighted_cross_entropy",
           "selection_mask"]


@torch.no_grad()
def update_sample_weights(residuals, weights, tol=1e-3, maxiter=40):
    """train_rlvi.py:14-38 -- in place on BOTH tensors: one persistent kernel instead of ~8 launches and
    one host sync per iteration."""
    ops.fixed_point_deep(residuals, weights, tol=tol, maxiter=maxiter)


@torch.no_grad()
def false_negative_criterion(weights, alpha=0.05):
    """train_rlvi.py:41-49 -- returns a 0-dim tensor on the weights' device."""
    return ops.fn_threshold(weights, alpha=alpha, prev_threshold=0.0, truncate=False).reshape(())


@torch.no_grad()
def selection_mask(weights, threshold):
    """deep-learning/main.py:342 -- the samples the method currently treats as clean: weights > threshold."""
    return weights > threshold


class _WeightedCE(torch.autograd.Function):
    """loss = mean_i CE(logits_i, label_i) * weights[indexes_i]; residuals[indexes_i] = CE_i (detached,
    quirk Q8); d loss / d logits = (softmax - onehot) * weights[indexes_i] / B, computed in the forward
    launch and handed to autograd in backward."""

    @staticmethod
    def forward(ctx, logits, labels, indexes, weights, residuals):
        lg = logits.contiguous()
        if lg.dtype != torch.float32:
            lg = lg.float()
        need_grad = ctx.needs_input_grad[0]
        r = ops.wce_fwd_bwd(lg, labels, weights, residuals, indexes=indexes, want_grad=need_grad,
                            want_correct=True)
        ctx.in_dtype = logits.dtype
        if need_grad:
            ctx.save_for_backward(r["dlogits"])
        ctx.mark_non_differentiable(r["correct"])
        return r["loss"].reshape(()), r["correct"]

    @staticmethod
    def backward(ctx, grad_loss, _grad_correct):
        (dlogits,) = ctx.saved_tensors
        g = dlogits * grad_loss
        return g.to(ctx.in_dtype), None, None, None, None


def weighted_cross_entropy(logits, labels, indexes, weights, residuals):
    """Returns (loss 0-dim, correct int32[2] = rows whose label is in the top-1 / top-5 logits)."""
    return _WeightedCE.apply(logits, labels, indexes, weights, residuals)


def train_rlvi(train_loader, model, optimizer, residuals, weights, overfit, threshold):
    """train_rlvi.py:52-106 -- one epoch.  Same arguments, same in-place updates of `residuals` /
    `weights`, same return value `(train_acc, threshold)`."""
    device = weights.device
    train_total = 0
    train_correct = torch.zeros((), dtype=torch.float64, device=device)

    for (images, labels, indexes) in train_loader:
        images = images.to(device, non_blocking=True)                   # train_rlvi.py:81
        labels = labels.to(device, non_blocking=True)                   # train_rlvi.py:82
        indexes = torch.as_tensor(indexes).to(device=device, dtype=torch.int64, non_blocking=True)

        logits = model(images)                                          # train_rlvi.py:84
        loss, correct = weighted_cross_entropy(logits, labels, indexes, weights, residuals)   # :85-94
        train_total += 1
        train_correct += correct[0] * (100.0 / labels.size(0))          # utils.py:78 prec@1 in percent
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()

    update_sample_weights(residuals, weights)                           # train_rlvi.py:99
    if overfit:
        # train_rlvi.py:100-103: threshold = max(threshold, criterion); weights[weights < threshold] = 0
        new = ops.fn_threshold(weights, alpha=0.05, prev_threshold=float(threshold), truncate=True).reshape(())
        threshold = max(threshold, new)

    train_acc = float(train_correct) / float(train_total)
    return train_acc, threshold
