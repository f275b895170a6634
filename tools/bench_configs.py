#!/usr/bin/env python
"""Measurements of the BASELINE.json configs that are not the headline bench line (configs[1] second model,
configs[2..4]) and of the drop-in outer loops.  bench.py imports `run_all` for its `configs` block (N = 1 only), so
these numbers land in the driver-run record; run as a script it prints one JSON object per line.

  2b  covariance model pieces at the headline shape (FP64, N = 2^26, d = 64): mean pass, centred Gram, Gaussian
      NLL, epsilon fixed point, one KKT-shift objective evaluation (utils.py:92-108, rlvi.py:23-43)
  3   robust PCA E+M step, FP32-stored X, N = 2^24, d = 512: rlvi_loss_f32(PCA) + fixed point +
      rlvi_weighted_moments_f32 (tcgen05 TF32 Gram), 3xTF32 and single-pass TF32
  4   online E-step on the 121-batch HAR-shaped stream (batches of 100 losses), reference restart and carry-over
  5a  fused weighted CE fwd+bwd + accuracy (8192 x 100 FP32)
  5b  per-epoch E-step + threshold + truncation (N_train = 45 000)
each next to the same step written with stock torch / NumPy calls in the order the reference makes them, on the same
GPU (or host).  Timing: CUDA events, 3 warm-ups, median.  Nothing under oracle/ is imported here.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rlvi_b200 import deep, ops, rlvi, synth, utils  # noqa: E402


@torch.no_grad()
def stock_epoch_tail(residuals, weights):
    """Baseline for config 5b: the per-epoch tail of train_rlvi.py (lines 14-38, 41-49, 100-103) as the ~10 stock
    torch launches + one host sync per pass the reference issues."""
    residuals.sub_(residuals.min())
    e = torch.exp(-residuals)
    avg = 0.95
    for _ in range(40):
        ratio = avg / (1 - avg)
        new = torch.div(ratio * e, 1 + ratio * e)
        err = torch.norm(new - weights)
        weights[:] = new
        avg = weights.mean()
        if err < 1e-3:
            break
    weights.div_(weights.max())
    beta = torch.sum(1 - weights) * 0.05
    w_sorted, _ = torch.sort(weights, dim=0, descending=True)
    threshold = w_sorted[torch.sum(torch.cumsum(1 - w_sorted, dim=0) <= beta) - 1]
    weights[weights < threshold] = 0
    return threshold


def numpy_online_estep(losses, tol=1e-3, maxiter=100):
    """Baseline for config 4: online-learning/main.py:45-58 in NumPy on the host."""
    e = np.exp(-losses)
    w = np.full_like(losses, 0.5)
    new = w
    for _ in range(maxiter):
        avg = np.mean(w)
        ratio = avg / (1 - avg)
        new = ratio * e / (1 + ratio * e)
        if np.linalg.norm(new - w) < tol:
            break
        w = new
    return new / (np.max(new) * len(new))


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def tensor_roofline(tf32_tflops, bf16_tflops=0.0, kernel="gram_tf32_pair_kernel<3>"):
    """The TF32 Gram against the tensor roofline that applies to it: TF32 runs at half the bf16 rate, and under tensor load
    the SM clock is power-managed, so the denominator is MEASURED_PEAKS.json's SUSTAINED cuBLAS bf16 figure (/ 2 for TF32).
    The default ~1e-6 mode issues its hi.hi product in TF32 and its two correction products in BF16: `frac` is then the
    share of the run time the tensor pipe needs for both at their own peaks, `achieved` the TF32-equivalent rate."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        peaks = json.load(open(path))
        bf16, src = float(peaks["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained; TF32 = half of it)"
    except Exception:
        bf16, src = 2250.0, "nominal dense bf16 2250 TFLOP/s, TF32 = half of it (MEASURED_PEAKS.json absent)"
    peak = bf16 / 2.0
    equiv = tf32_tflops + bf16_tflops / 2.0          # a BF16 flop costs half the tensor time of a TF32 flop
    return {"bound": "tensor", "kernel": kernel, "achieved": equiv, "peak": peak, "unit": "TFLOP/s (TF32-equivalent)",
            "frac": equiv / peak, "issued_tf32_TFLOPs": tf32_tflops, "issued_bf16_TFLOPs": bf16_tflops, "peak_source": src}


def config3(dev, log2n=24, d=512, seed=3):
    """Config 3 at its named size: one robust-PCA E+M step on FP32-stored X (32 GiB)."""
    n = 1 << log2n
    X, v = synth.pca_rows_torch(0, n, d, dev, seed)
    theta = torch.ones(d, dtype=torch.float64, device=dev)
    theta[0] = 0.1
    theta /= theta.norm()                                  # theta_init = normalised (0.1, 1, ..., 1)
    losses = torch.empty(n, dtype=torch.float64, device=dev)
    e = torch.empty_like(losses)
    pi = torch.empty_like(losses)
    res = torch.empty(5, dtype=torch.float64, device=dev)
    out3 = out1 = None

    def loss_pass():
        ops.loss(ops.LOSS_PCA, X, theta, losses_out=losses)

    def estep():
        ops.fixed_point(losses, e_work=e, out=pi, result=res)

    loss_pass()
    estep()
    iters = ops.read_result(res)["iters"]
    t_loss = timeit(loss_pass, reps=5)
    t_fp = timeit(estep, reps=5)
    out3 = ops.weighted_moments(X, pi, power=2, precision=ops.TF32X3)
    out1 = ops.weighted_moments(X, pi, power=2, precision=ops.TF32X1)
    t_g3 = timeit(lambda: ops.weighted_moments(X, pi, power=2, precision=ops.TF32X3, out=out3), reps=5)
    t_g1 = timeit(lambda: ops.weighted_moments(X, pi, power=2, precision=ops.TF32X1, out=out1), reps=5)
    G3 = ops.split_moments(out3, d)["G"]
    G1 = ops.split_moments(out1, d)["G"]
    # comparators on the same GPU: cuBLAS SGEMM (FP32) and TF32 of (pi^2 X)^T X in 2^20-row chunks
    w2 = (pi * pi).to(torch.float32)

    def blas(allow_tf32):
        torch.backends.cuda.matmul.allow_tf32 = allow_tf32
        acc = torch.zeros((d, d), dtype=torch.float32, device=dev)
        for s in range(0, n, 1 << 20):
            xb = X[s:s + (1 << 20)]
            acc.addmm_((xb * w2[s:s + (1 << 20), None]).T, xb)
        torch.backends.cuda.matmul.allow_tf32 = False
        return acc

    t_sgemm = timeit(lambda: blas(False), reps=3, warm=1)
    t_tf32 = timeit(lambda: blas(True), reps=3, warm=1)
    # parity of the statistic against an FP64 evaluation on the first 2^20 rows (cuBLAS DGEMM of the converted rows)
    m = 1 << min(20, log2n)
    Xh = X[:m].to(torch.float64)
    Gref = (Xh * (pi[:m] ** 2)[:, None]).T @ Xh
    o3 = ops.split_moments(ops.weighted_moments(X[:m], pi[:m].contiguous(), power=2, precision=ops.TF32X3), d)["G"]
    o1 = ops.split_moments(ops.weighted_moments(X[:m], pi[:m].contiguous(), power=2, precision=ops.TF32X1), d)["G"]
    scale = float(Gref.abs().max())
    step3 = t_loss + t_fp + t_g3
    pure3 = os.environ.get("RLVI_TF32_PURE3", "0") not in ("", "0")
    blk = 12 * 2 * n * 128 * 128                       # flop of one pass over the 12 issued 128 x 128 blocks
    out = {"config": f"3: robust PCA E+M step, FP32-stored X, N=2^{log2n}, d={d}", "n": n, "d": d,
           "fixed_point_passes": iters, "loss_f32_ms": t_loss, "loss_f32_GBps": n * (d * 4 + 8) / t_loss / 1e6,
           "fixed_point_ms": t_fp, "gram_tf32x3_ms": t_g3, "gram_tf32x1_ms": t_g1,
           "gram_tf32x3_GBps": n * (d * 4 + 8) / t_g3 / 1e6, "gram_tf32x1_GBps": n * (d * 4 + 8) / t_g1 / 1e6,
           "gram_useful_TFLOPs_x3": n * d * (d + 1) / t_g3 / 1e9, "gram_useful_TFLOPs_x1": n * d * (d + 1) / t_g1 / 1e9,
           # issued: the pair kernel computes 12 of the 16 128 x 128 blocks (the 10 of the upper triangle + the redundant
           # lower block of each diagonal pair), 3 passes for 3xTF32; 2 * rows * 128 * 128 flop per block
           # (default mode: the hi.hi pass in TF32, the two correction passes in BF16; RLVI_TF32_PURE3=1: all three in TF32)
           "gram_x3_mode": "pure 3xTF32" if pure3 else "TF32 hi.hi + BF16 corrections",
           "gram_issued_TFLOPs_x3": 3 * 12 * 2 * n * 128 * 128 / t_g3 / 1e9 if d == 512 else 3 * n * d * (d + 128) / t_g3 / 1e9,
           "gram_issued_TFLOPs_x1": 12 * 2 * n * 128 * 128 / t_g1 / 1e9 if d == 512 else n * d * (d + 128) / t_g1 / 1e9,
           "tensor_roofline_x3": (tensor_roofline((3 if pure3 else 1) * blk / t_g3 / 1e9, (0 if pure3 else 2) * blk / t_g3 / 1e9,
                                                  "gram_tf32_pair_kernel<3>" if pure3 else "gram_tf32_pair_kernel<2>")
                                  if d == 512 else None),
           "step_ms_tf32x3": step3, "samples_per_s_tf32x3": n / step3 * 1e3,
           "samples_per_s_tf32x1": n / (t_loss + t_fp + t_g1) * 1e3,
           "hbm_floor_ms_per_X_pass": n * d * 4 / 6538e6,
           "stock_torch_sgemm_fp32_ms": t_sgemm, "stock_torch_tf32_ms": t_tf32,
           "gram_rel_err_vs_fp64_first_2^20_rows": {"tf32x3": float((o3 - Gref).abs().max()) / scale,
                                                   "tf32x1": float((o1 - Gref).abs().max()) / scale},
           "x3_vs_x1_full_rel": float((G3 - G1).abs().max() / G3.abs().max())}
    del X, losses, e, pi, w2, Xh
    torch.cuda.empty_cache()
    return out


def config2_covariance(dev, X=None, log2n=26, d=64):
    """configs[1], second model: the covariance pieces at the headline shape (reuses the caller's X when given)."""
    if X is None:
        n = 1 << log2n
        X = torch.empty((n, d), device=dev, dtype=torch.float64)
        g = torch.Generator(device=dev).manual_seed(0)
        for s in range(0, n, 1 << 22):
            X[s:s + (1 << 22)].normal_(generator=g)
    n, d = X.shape
    g = torch.Generator(device=dev).manual_seed(1)
    w = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    mu = torch.zeros(d, dtype=torch.float64, device=dev)
    cov = torch.eye(d, dtype=torch.float64, device=dev)
    params = utils._gaussian_params(mu, cov)
    losses = torch.empty(n, dtype=torch.float64, device=dev)
    t_loss = timeit(lambda: ops.loss(ops.LOSS_GAUSSIAN, X, params, losses_out=losses), reps=5)
    t_mean = timeit(lambda: ops.weighted_moments(X, w, want_gram=False), reps=5)
    t_gram = timeit(lambda: ops.weighted_moments(X, w, center=mu), reps=5)
    acc = torch.empty(1, dtype=torch.float64, device=dev)
    e = torch.exp(-(losses - losses.min()))
    t_shift_e = timeit(lambda: ops.shift_sum_e(e, 1.35, 0.4, out=acc), reps=5)
    lz = losses - losses.min()
    pi = torch.empty_like(lz)
    t_fp = timeit(lambda: ops.fixed_point(lz, e_work=e, out=pi), reps=3)
    step = t_mean + t_gram + t_loss + t_fp + 15 * t_shift_e
    # comparator: the same three X passes with stock torch ops (cuBLAS DGEMM / GEMV) on 2^22-row chunks
    def stock():
        S = torch.zeros((d, d), dtype=torch.float64, device=dev)
        m1 = torch.zeros(d, dtype=torch.float64, device=dev)
        for s in range(0, n, 1 << 22):
            xb, wb = X[s:s + (1 << 22)], w[s:s + (1 << 22)]
            m1 += xb.T @ wb
        m1 /= w.sum()
        for s in range(0, n, 1 << 22):
            xb, wb = X[s:s + (1 << 22)], w[s:s + (1 << 22)]
            c = xb - m1
            S.addmm_((c * wb[:, None]).T, c)
            z = torch.linalg.solve_triangular(torch.linalg.cholesky(cov), c.T, upper=False)
            losses[s:s + (1 << 22)] = 0.5 * (z * z).sum(0)
        return S
    t_stock = timeit(stock, reps=2, warm=1)
    out = {"config": f"2b: covariance model pieces, FP64, N=2^{int(np.log2(n))}, d={d}", "n": n,
           "gaussian_loss_ms": t_loss, "gaussian_loss_GBps": n * (d + 1) * 8 / t_loss / 1e6,
           "gaussian_loss_tflops": n * 72 * 512 / 8 / t_loss / 1e9, "mean_pass_ms": t_mean,
           "mean_pass_GBps": n * (d + 1) * 8 / t_mean / 1e6, "centred_gram_ms": t_gram,
           "centred_gram_GBps": n * (d + 1) * 8 / t_gram / 1e6, "shift_sum_e_ms": t_shift_e,
           "shift_sum_e_GBps": n * 8 / t_shift_e / 1e6, "fixed_point_ms": t_fp,
           "em_step_ms_with_15_shift_evaluations": step, "samples_per_s": n / step * 1e3,
           "stock_torch_three_passes_ms": t_stock,
           "note": "one covariance E+M step = mean pass + centred Gram + Gaussian NLL + fixed point + ~15 Brent "
                   "evaluations of the KKT-shift sum (rlvi.py:34-42)"}
    del w, losses, e, lz, pi
    torch.cuda.empty_cache()
    return out


def config4(dev, batches=121, batch=100, seed=1):
    """The online stream: 121 batches of 100 per-sample losses (24 075 x 60 HAR-shaped data, half of it streamed in
    batches of 100: online-learning/main.py:226-231,286-297); one E-step kernel per batch."""
    rng = np.random.default_rng(seed)
    lh = [rng.exponential(0.7, size=batch) for _ in range(batches)]
    ld = [torch.from_numpy(a).to(dev) for a in lh]
    outs = [torch.empty(batch, dtype=torch.float64, device=dev) for _ in range(batches)]
    ew = torch.empty(batch, dtype=torch.float64, device=dev)
    res = torch.empty(5, dtype=torch.float64, device=dev)

    def stream():
        for l, o in zip(ld, outs):
            ops.fixed_point(l, e_work=ew, out=o, result=res, variant=ops.FP_ONLINE)

    t = timeit(stream, reps=10)
    t0 = time.perf_counter()
    for _ in range(5):
        for a in lh:
            numpy_online_estep(a)
    t_cpu = (time.perf_counter() - t0) / 5 * 1e3
    # what the reference's call pattern costs: NumPy losses in, NumPy weights out, per batch (H2D + kernel + D2H + sync)
    from rlvi_b200 import online
    t0 = time.perf_counter()
    for a in lh:
        online.update_weights_rlvi(a)
    t_call = (time.perf_counter() - t0) * 1e3
    return {"config": f"4: online E-step stream, {batches} batches of {batch} (online-learning/main.py:45-58)",
            "device_resident_stream_ms": t, "per_batch_us": t / batches * 1e3,
            "dropin_numpy_in_out_stream_ms": t_call, "numpy_cpu_stream_ms": t_cpu,
            "samples_per_s_device": batches * batch / t * 1e3, "samples_per_s_numpy_cpu": batches * batch / t_cpu * 1e3,
            "note": "single-CTA kernel per batch, launch-latency bound; the classifier update (sklearn SGD) is out of scope"}


def config5(dev):
    b, c, n_train = 8192, 100, 45000
    logits_np, labels_np = synth.deep_batch(b, c, seed=1)
    logits = torch.from_numpy(logits_np).to(dev)
    labels = torch.from_numpy(labels_np).to(dev)
    idx = torch.randperm(n_train, device=dev)[:b]
    weights = torch.rand(n_train, device=dev)
    residuals = torch.zeros(n_train, device=dev)

    def fused():
        ops.wce_fwd_bwd(logits, labels, weights, residuals, indexes=idx, want_correct=True)

    def stock():                                   # train_rlvi.py:85,89-96 with stock torch ops
        lg = logits.detach().requires_grad_(True)
        out = torch.softmax(lg, dim=1)
        _, pred = out.topk(5, 1, True, True)
        pred.t().eq(labels.view(1, -1).expand_as(pred.t()))
        loss = torch.nn.functional.cross_entropy(lg, labels, reduction="none")
        residuals[idx] = loss.detach()
        (loss * weights[idx]).mean().backward()

    t_f, t_s = timeit(fused, reps=30), timeit(stock, reps=30)
    a = {"config": "5a: weighted CE fwd+bwd + accuracy, 8192 x 100 FP32", "fused_kernel_ms": t_f, "stock_torch_ms": t_s,
         "speedup": t_s / t_f, "samples_per_s": b / t_f * 1e3}

    res0 = torch.from_numpy(np.random.default_rng(2).exponential(1.0, size=n_train).astype(np.float32)).to(dev)

    def ours():
        r, w = res0.clone(), torch.ones(n_train, device=dev)
        deep.update_sample_weights(r, w)
        ops.fn_threshold(w, truncate=True)

    def ref():
        r, w = res0.clone(), torch.ones(n_train, device=dev)
        stock_epoch_tail(r, w)

    t_o, t_r = timeit(ours, reps=20), timeit(ref, reps=20)
    bb = {"config": "5b: per-epoch E-step + threshold + truncation, N_train=45000 FP32", "kernels_ms": t_o,
          "reference_torch_ops_same_gpu_ms": t_r, "speedup": t_r / t_o}
    return a, bb


def loops(dev, log2n=22, d=64):
    n = 1 << log2n
    X, y, theta = synth.logistic_rows_torch(0, n, d, dev, seed=5)
    yl = (X @ torch.ones(d, dtype=torch.float64, device=dev)) + torch.randn(n, device=dev, dtype=torch.float64)
    out = []
    for name, fn in (("rlvi.mean", lambda: rlvi.mean(X)),
                     ("rlvi.linear_regression", lambda: rlvi.linear_regression(X, yl)),
                     ("rlvi.logistic_regression(mm)", lambda: rlvi.logistic_regression(X, y, mstep="mm")),
                     ("rlvi.logistic_regression(sklearn)", lambda: rlvi.logistic_regression(X, y)),
                     ("utils.mm_log_reg", lambda: utils.mm_log_reg(X, y, torch.ones(n, dtype=torch.float64, device=dev)))):
        fn()                                            # warm-up (scratch growth, cuSOLVER handles)
        torch.cuda.synchronize()
        l0 = ops.launch_count(dev.index or 0)
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out.append({"loop": name, "n": n, "d": d, "wall_ms": (time.perf_counter() - t0) * 1e3,
                    "library_launches": ops.launch_count(dev.index or 0) - l0})
    return out


def run_all(dev, X=None, config3_log2n=24, with_loops=True):
    """Everything above as one dict (bench.py's `configs` block).  A failing piece reports its error instead of
    taking the others down."""
    out = {}

    def guarded(key, fn):
        try:
            out[key] = fn()
        except Exception as exc:          # reporting only
            out[key] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    guarded("config2b_covariance", lambda: config2_covariance(dev, X))
    guarded("config4_online_stream", lambda: config4(dev))
    try:
        out["config5a_wce"], out["config5b_epoch_tail"] = config5(dev)
    except Exception as exc:
        out["config5a_wce"] = {"error": repr(exc)[:300]}
    if with_loops:
        guarded("dropin_loops", lambda: loops(dev))
    return out


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    res = run_all(dev)
    try:
        res["config3_pca_fp32"] = config3(dev)
    except Exception as exc:
        res["config3_pca_fp32"] = {"error": repr(exc)[:300]}
    for k, v in res.items():
        print(json.dumps({k: v}), flush=True)
