"""Bring-up / accuracy sweep of rlvi_weighted_moments_f32 and rlvi_loss_f32 against NumPy FP64 on the same float32
samples (run on the GPU box: python tools/tf32_bringup.py [quick])."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from rlvi_b200 import ops


def ref_moments(X32, w, y, power):
    X = X32.astype(np.float64)
    we = w * w if power == 2 else w
    out = {"S0": we.sum(), "S1": X.T @ w, "G": (X * we[:, None]).T @ X}
    if y is not None:
        out["Swy"] = we @ y
        out["Sy"] = X.T @ (we * y)
    return out


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def case(n, d, power, with_y, precision, seed=0, wscale=1.0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)).astype(np.float32)
    X[:, 0] += 3.0
    w = wscale * rng.random(n) ** 4
    y = rng.normal(size=n) if with_y else None
    dev = torch.device("cuda", 0)
    Xd = torch.from_numpy(X).to(dev)
    wd = torch.from_numpy(w).to(dev)
    yd = torch.from_numpy(y).to(dev) if with_y else None
    out = ops.weighted_moments(Xd, wd, y=yd, power=power, precision=precision)
    torch.cuda.synchronize()
    m = {k: v.cpu().numpy() for k, v in ops.split_moments(out, d).items()}
    r = ref_moments(X, w, y, power)
    res = {"n": n, "d": d, "power": power, "y": with_y, "prec": precision, "G": rel(m["G"], r["G"]),
           "S1": rel(m["S1"], r["S1"]), "S0": abs(m["S0"] - r["S0"]) / r["S0"],
           "sym": float(np.max(np.abs(m["G"] - m["G"].T)))}
    if with_y:
        res["Sy"] = rel(m["Sy"], r["Sy"])
        res["Swy"] = abs(m["Swy"] - r["Swy"]) / abs(r["Swy"])
    return res


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = [(4096, 128, 2, False), (1000, 64, 2, False), (5000, 256, 1, True), (20000, 512, 2, False),
             (777, 100, 1, True), (33, 512, 2, False), (300000, 384, 1, False)]
    if quick:
        cases = cases[:2]
    for c in cases:
        for prec in (ops.TF32X3, ops.TF32X1):
            print(json.dumps(case(*c, prec)), flush=True)
    print(json.dumps(dict(case(20000, 512, 2, True, ops.TF32X3, wscale=1e-20), wscale=1e-20)), flush=True)
    print(json.dumps(dict(case(20000, 256, 1, True, ops.TF32X3, wscale=1e-150), wscale=1e-150)), flush=True)
    # loss_f32 vs numpy
    rng = np.random.default_rng(1)
    n, d = 5000, 512
    X = rng.normal(size=(n, d)).astype(np.float32)
    th = rng.normal(size=d)
    th /= np.linalg.norm(th)
    dev = torch.device("cuda", 0)
    l, e, _ = ops.loss(ops.LOSS_PCA, torch.from_numpy(X).to(dev), torch.from_numpy(th).to(dev), want_e=True)
    X64 = X.astype(np.float64)
    lr = (X64 ** 2).sum(1) - (X64 @ th) ** 2
    print(json.dumps({"loss_f32_pca": rel(l.cpu().numpy(), lr), "e": rel(e.cpu().numpy(), np.exp(-lr))}), flush=True)
    if quick:
        return
    # timing at config-3 shape (reduced N)
    for log2n in (20, 22):
        n, d = 1 << log2n, 512
        Xd = torch.randn((n, d), device=dev, dtype=torch.float32)
        wd = torch.rand(n, device=dev, dtype=torch.float64)
        for prec in (ops.TF32X3, ops.TF32X1):
            out = None
            for _ in range(2):
                out = ops.weighted_moments(Xd, wd, power=2, precision=prec, out=out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                ops.weighted_moments(Xd, wd, power=2, precision=prec, out=out)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            flops = n * d * (d + 128) * (3 if prec == ops.TF32X3 else 1)   # 10 of 16 blocks, issued flops
            print(json.dumps({"time_n": n, "d": d, "prec": prec, "ms": ms, "GBps": n * d * 4 / ms / 1e6,
                              "issued_TFLOPs": flops / ms / 1e9}), flush=True)
        th = torch.randn(d, device=dev, dtype=torch.float64)
        lo = torch.empty(n, device=dev, dtype=torch.float64)
        for _ in range(2):
            ops.loss(ops.LOSS_PCA, Xd, th, losses_out=lo)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            ops.loss(ops.LOSS_PCA, Xd, th, losses_out=lo)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(json.dumps({"loss_f32_n": n, "ms": ms, "GBps": n * (d * 4 + 8) / ms / 1e6}), flush=True)
        del Xd, wd


if __name__ == "__main__":
    main()
