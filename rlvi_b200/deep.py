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
           "selection_mask", "GraphedBatchStep"]


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


class GraphedBatchStep:
    """The per-batch body of train_rlvi.py:84-97 -- model forward, fused weighted cross-entropy (residual scatter, weight
    gather, accuracy counts), backward, optimizer step -- captured ONCE as a CUDA graph for one batch shape and replayed
    per batch: a single graph launch instead of the model's and the optimizer's individual launches (SURVEY.md section 8f
    rank 3).  `optimizer` must be capturable (torch.optim.SGD, or Adam(..., capturable=True) -- the reference's Adam of
    deep-learning/main.py:263 with that flag).  The warm-up iterations torch needs before a capture run on the given
    batch and are undone (model, optimizer state and `residuals` are restored), so training is the same as eager."""

    def __init__(self, model, optimizer, residuals, weights, images, labels, indexes, warmup=3):
        import copy

        self.model, self.optimizer = model, optimizer
        self.images, self.labels, self.indexes = images.clone(), labels.clone(), indexes.clone()
        self.stream = torch.cuda.Stream(device=weights.device)
        saved_model = copy.deepcopy(model.state_dict())
        saved_opt = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in optimizer.state.items()}
        saved_res = residuals.clone()
        self.stream.wait_stream(torch.cuda.current_stream(weights.device))
        with torch.cuda.stream(self.stream):          # librlvi_b200 keeps one context per stream: warm it up on this one
            for _ in range(warmup):
                optimizer.zero_grad(set_to_none=True)
                loss, _ = weighted_cross_entropy(model(self.images), self.labels, self.indexes, weights, residuals)
                loss.backward()
                optimizer.step()
            # undo the warm-up IN PLACE: the optimizer's state tensors must exist before the capture (state created inside
            # it would be re-initialised by every replay), so they are kept and reset to what they held before
            model.load_state_dict(saved_model)
            for prm, st in optimizer.state.items():
                old = saved_opt.get(prm, {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
                    elif k in old:
                        st[k] = old[k]
            residuals.copy_(saved_res)
        torch.cuda.current_stream(weights.device).wait_stream(self.stream)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.loss, self.correct = weighted_cross_entropy(model(self.images), self.labels, self.indexes, weights,
                                                             residuals)
            self.loss.backward()
            optimizer.step()

    def matches(self, images, labels):
        return images.shape == self.images.shape and images.dtype == self.images.dtype and labels.shape == self.labels.shape

    def __call__(self, images, labels, indexes):
        """One training step on this batch; returns (loss, correct) -- tensors the NEXT call overwrites."""
        self.images.copy_(images, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.indexes.copy_(indexes, non_blocking=True)
        self.graph.replay()
        return self.loss, self.correct


def train_rlvi(train_loader, model, optimizer, residuals, weights, overfit, threshold, cuda_graph=None):
    """train_rlvi.py:52-106 -- one epoch.  Same arguments, same in-place updates of `residuals` /
    `weights`, same return value `(train_acc, threshold)`.  `cuda_graph` (an extension, default off): a dict the caller
    keeps across epochs; full-shape batches then run as one replayed CUDA graph (GraphedBatchStep), ragged ones eagerly."""
    device = weights.device
    train_total = 0
    train_correct = torch.zeros((), dtype=torch.float64, device=device)

    for (images, labels, indexes) in train_loader:
        images = images.to(device, non_blocking=True)                   # train_rlvi.py:81
        labels = labels.to(device, non_blocking=True)                   # train_rlvi.py:82
        indexes = torch.as_tensor(indexes).to(device=device, dtype=torch.int64, non_blocking=True)

        step = None
        if cuda_graph is not None:
            step = cuda_graph.get("step")
            if step is None:
                step = cuda_graph["step"] = GraphedBatchStep(model, optimizer, residuals, weights, images, labels, indexes)
            if not step.matches(images, labels):
                step = None
        if step is not None:
            _, correct = step(images, labels, indexes)                  # :84-97 as one graph launch
            train_total += 1
            train_correct += correct[0] * (100.0 / labels.size(0))
            continue
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
