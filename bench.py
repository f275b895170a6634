#!/usr/bin/env python
"""bench.py -- RLVI E+M step throughput on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W     # N > 1 (the driver's launch)
    python bench.py --impl reference ...                           # the reference's CPU path (oracle port)

Workload (SURVEY.md section 8d, configs[1]): logistic model, N = 2^26 samples x d = 64 features FP64, 30 %
label corruption, synthetic, generated on the device.  One STEP = one E+M pass over all samples:
  (1) loss pass      rlvi_loss_f64(LOGISTIC_CE)   reads X, y -> writes e = exp(-loss)
  (2) E-step         rlvi_fixed_point_f64         K_fp passes over e until the reference's stop rule
                                                  (tol 1e-3, maxiter 100) fires, then writes pi
  (3) statistics     rlvi_weighted_moments_f64    reads X, pi -> S0, X^T pi, X^T Pi X  (+ all-reduce, N > 1)
N > 1 shards the samples (strong scaling: the same 2^26 samples over N GPUs -- every block of 2^20 global rows is
drawn from its own seeded generator, rlvi_b200.synth.logistic_rows_torch, so 1, 2, 4 and 8 ranks process identical
data and the `fixed_point` / `stats_checksum` blocks of their lines can be compared); the fixed point exchanges its
three partial sums per pass inside the kernel over NVLink peer memory, and the d x d statistics are summed by a
small kernel over the same peer windows (rlvi_stats_allreduce_f64; NCCL only carries the set-up and the timing
reductions).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rlvi_em_step_samples_per_sec"
UNIT = "samples/s"


def claim_stdout():
    """Reserve the process's stdout for the ONE JSON line: fd 1 is re-pointed at stderr for everything else
    (NCCL prints its version banner to stdout when NCCL_DEBUG is set, as it is on the GPU boxes)."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def emit(out, line):
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=26, help="total samples = 2^log2n (default: the named config)")
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--cpu-log2n", type=int, default=22, help="bounded CPU sample for cpu_baseline / reference arm")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (configs 2b, 3, 4, 5; N = 1 only)")
    ap.add_argument("--config3-log2n", type=int, default=24)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks: sample NVML during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference's E+M step (reference is pure Python/NumPy)
# --------------------------------------------------------------------------------------------------
def force_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must use every host core whatever the
    launcher did.  Called BEFORE NumPy is imported (bench.py imports it lazily), and again through threadpoolctl."""
    n = os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[k] = str(n)
    return n


def cpu_em_step_bench(log2n, d, steps, warmup, keep=False):
    ncpu = force_host_threads()
    import numpy as np
    from oracle import rlvi_np
    from rlvi_b200 import synth

    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=ncpu)
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        blas_threads = 1
    n = 1 << log2n
    X, y, theta = synth.logistic_data(n, d, seed=0)
    params = np.concatenate([[0.0], theta])
    ref = None
    for _ in range(warmup):
        ref = rlvi_np.em_step_logistic(X, y, params)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref = rlvi_np.em_step_logistic(X, y, params)
    dt = (time.perf_counter() - t0) / steps
    iters = ref["iters"]
    cb = {"value": n / dt, "unit": UNIT, "cores": int(blas_threads), "kind": "port",
          "sample": f"oracle.rlvi_np.em_step_logistic (NumPy restatement of rlvi.py:8-20 + utils.py:19-21,36-38) on "
                    f"N=2^{log2n} x d={d} FP64 of the same synthetic workload, {steps} step(s) of {dt:.2f} s, "
                    f"{iters} fixed-point passes; BLAS threads={blas_threads}, os.cpu_count()={os.cpu_count()}"}
    if keep:
        # the CPU side of `e2e_call`: the oracle's whole EM loop (rlvi.py:92-108 with utils.mm_log_reg) on the same
        # arrays, timed here so that the oracle is only ever touched by this CPU leg
        # (bounded sample: the reference's MM M-step has no iteration cap and takes hundreds of passes per EM iteration;
        # 2^19 rows is ~25 s of CPU work, the whole 2^22 sample was 197 s)
        n_loop = min(n, 1 << 19)
        t0 = time.perf_counter()
        th_ref = rlvi_np.logistic_regression_mm(X[:n_loop], y[:n_loop])
        t_loop = time.perf_counter() - t0
        return cb, dt, (X, y, params, ref, th_ref, t_loop, n_loop)
    return cb, dt


def workload_config(args, world):
    """The `config` both arms report (BASELINE.json configs[1])."""
    return {"workload": f"logistic E+M step (loss pass + epsilon fixed point tol 1e-3 + pi-weighted X^T Pi X), "
                        f"N=2^{args.log2n} samples x d={args.d} FP64, 30% label corruption",
            "n_total": 1 << args.log2n, "d": args.d, "fixed_point_tol": 1e-3, "fixed_point_maxiter": 100,
            "parallelism": f"sample-sharded dp{world}"}


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    force_host_threads()                      # before NumPy loads its BLAS: torchrun exports OMP_NUM_THREADS=1
    steps = max(1, min(args.steps, 5))
    warmup = 1 if args.warmup > 0 else 0
    cb, dt = cpu_em_step_bench(args.cpu_log2n, args.d, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args, args.gpus),
                           sample=f"each step runs a bounded sample: N=2^{args.cpu_log2n} rows of the same workload"),
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out, line)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args, out):
    import numpy as np
    import torch

    from rlvi_b200 import _lib, dist as rdist, ops, synth

    _lib.load()                                     # no CPU fallback: fail loudly
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_cpus(local) if world > 1 else None
    group = rdist.ShardGroup.create(dev) if world > 1 else None

    n_total = 1 << args.log2n
    d = args.d
    lo, hi = rdist.shard_bounds(n_total, rank, world)
    n = hi - lo
    X, y, theta = synth.logistic_rows_torch(lo, hi, d, dev, seed=1234)      # a function of the GLOBAL row index
    params = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), theta]).contiguous()
    e = torch.empty(n, dtype=torch.float64, device=dev)
    pi = torch.empty(n, dtype=torch.float64, device=dev)
    res = torch.empty(5, dtype=torch.float64, device=dev)
    mom = torch.zeros(_lib.load().rlvi_moments_out_doubles(d), dtype=torch.float64, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step(k=None):
        if k is not None:
            ev[k][0].record()
        ops.loss(ops.LOSS_LOGISTIC_CE, X, params, y=y, intercept=True, want_losses=False, want_e=True, e_out=e)
        if k is not None:
            ev[k][1].record()
        ops.fixed_point(None, e_work=e, out=pi, result=res, dist=group.fp_dist(n_total) if group else None)
        if k is not None:
            ev[k][2].record()
        ops.weighted_moments(X, pi, out=mom)
        if group:
            group.all_reduce(mom)
        if k is not None:
            ev[k][3].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    fp = ops.read_result(res)
    sampler = ClockSampler(local)
    launches0 = ops.launch_count(local)
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    sampler.start()
    if world > 1:
        # align the ranks on the GPU timeline: the CPUs leave the barrier above milliseconds apart, and a rank
        # that starts early would only spin in its first sharded fixed-point pass waiting for the others
        align = torch.zeros(1, device=dev)
        torch.distributed.all_reduce(align)
    t_begin.record()
    for k in range(args.steps):
        step(k)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ops.launch_count(local) - launches0
    total_ms = t_begin.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(t)
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # per-kernel averages (this rank) and the roofline of the dominant kernel
    kms = np.array([[ev[k][i].elapsed_time(ev[k][i + 1]) for i in range(3)] for k in range(args.steps)]).mean(axis=0)
    gaps = [ev[k][3].elapsed_time(ev[k + 1][0]) for k in range(args.steps - 1)]
    k_fp = fp["iters"]
    alg_bytes = {"loss_kernel": n * (d * 8 + 8 + 8), "fp_kernel_f64": n * 8 * (k_fp + 1) + n * 8,
                 "gram64_kernel": n * (d * 8 + 8)}
    names = list(alg_bytes)
    kernels = {nm: {"ms": float(kms[i]), "algorithmic_bytes": int(alg_bytes[nm]),
                    "GBps": alg_bytes[nm] / (kms[i] * 1e-3) / 1e9} for i, nm in enumerate(names)}
    dom = names[int(np.argmax(kms))]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["GBps"] / peak, "traffic": None, "peak_source": peak_src,
                "step_frac": sum(alg_bytes.values()) / (ms_per_step * 1e-3) / 1e9 / peak,   # per GPU
                "peak_nominal": 8000.0, "frac_nominal": kernels[dom]["GBps"] / 8000.0,
                "step_frac_nominal": sum(alg_bytes.values()) / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                "note": "fractions against BOTH denominators (SURVEY.md H7): `peak` = measured copy bandwidth, "
                        "`peak_nominal` = the ~8 TB/s BASELINE.json's north_star names",
                "kernels": kernels}
    # the statistics pass is bounded by the FP64 tensor pipe, not by HBM: report that too
    gram_flops = n * 36 * 512 / 4          # 36 DMMA.8x8x4 (512 flop) per 4 rows
    roofline["fp64_tensor"] = {"kernel": "gram64_kernel", "achieved": gram_flops / (kms[2] * 1e-3) / 1e12, "peak": 36.9,
                               "unit": "TFLOP/s", "frac": gram_flops / (kms[2] * 1e-3) / 1e12 / 36.9,
                               "peak_source": "measured DMMA.8x8x4 issue rate, tools/ubench_dmma.cu "
                                              "(profiles/r01_ubench_dmma.txt); not in MEASURED_PEAKS.json"}
    line_extra = {"measured_fp64_tflops": {"value": 36.9, "unit": "TFLOP/s", "what": "DMMA.8x8x4 issue-rate peak of this GPU "
                                           "model (4.0 clk per instruction per SM at 1.965 GHz x 148 SMs)",
                                           "source": "tools/ubench_dmma.cu run under gpurun, profiles/r01_ubench_dmma.txt"}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(dom)
        except Exception:
            pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args, world), fixed_point_passes=k_fp,
                           l2="inputs larger than L2 (X shard %.1f GiB streamed twice per step)" % (n * d * 8 / 2 ** 30)),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "fixed_point": {k: fp[k] for k in ("eps", "iters", "converged", "sum_pi")},
            "host": {"cpus": os.cpu_count(), "inter_step_gap_ms": float(np.mean(gaps)) if gaps else 0.0,
                     "cpus_local_to_gpu": numa}}
    line.update(line_extra)
    # the step's (all-reduced) statistics, normalised: identical data on every world size => comparable across N
    mh = ops.split_moments(mom, d)
    line["stats_checksum"] = {"trace_G_over_S0": float(torch.trace(mh["G"]) / mh["S0"]),
                              "sum_abs_G_over_S0": float(mh["G"].abs().sum() / mh["S0"]),
                              "S0": float(mh["S0"]),
                              "note": "same 2^%d global samples at every N (seeded per 2^20-row block)" % args.log2n}

    # ---- comparable-across-runs variant: exactly 32 fixed-point passes ---------------
    def step_k32():
        ops.loss(ops.LOSS_LOGISTIC_CE, X, params, y=y, intercept=True, want_losses=False, want_e=True, e_out=e)
        ops.fixed_point(None, e_work=e, out=pi, result=res, tol=-1.0, maxiter=32,
                        dist=group.fp_dist(n_total) if group else None)
        ops.weighted_moments(X, pi, out=mom)
        if group:
            group.all_reduce(mom)

    step_k32()
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        torch.distributed.all_reduce(torch.zeros(1, device=dev))
    k0.record()
    for _ in range(5):
        step_k32()
    k1.record()
    barrier()
    k32_ms = k0.elapsed_time(k1) / 5
    if world > 1:
        t = torch.tensor([k32_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        k32_ms = float(t)
    line["fixed_k32"] = {"fixed_point_passes": 32, "ms_per_step": k32_ms, "value": n_total / (k32_ms * 1e-3), "steps": 5}

    # ---- what a full logistic M-step adds to the step (timed and reported separately) --
    try:
        line["m_step_extras"] = m_step_extras(X, y, pi, params, mom, d)
    except Exception as exc:       # reporting only: never lose the headline line to it
        line["m_step_extras"] = {"error": repr(exc)[:200]}

    # ---- e2e: the same step through the host-buffer C-ABI call, H2D/D2H inside the timed region --------
    if not args.no_e2e:
        line["e2e"] = run_e2e(args, X, y, params, dev, world, rank, n_total, group)
    if world > 1:
        torch.distributed.barrier()
    # ---- the other BASELINE.json configs, each with its stock-op comparator (N = 1 only; reporting) ----------
    if rank == 0 and world == 1 and not args.no_configs:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs
            line["configs"] = bench_configs.run_all(dev, X=X if args.log2n >= 22 else None)
        except Exception as exc:
            line["configs"] = {"error": repr(exc)[:300]}
    del X, y, e, pi
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_configs:
        try:
            line["configs"]["config3_pca_fp32"] = bench_configs.config3(dev, log2n=args.config3_log2n)
        except Exception as exc:
            line["configs"]["config3_pca_fp32"] = {"error": repr(exc)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _, held = cpu_em_step_bench(args.cpu_log2n, d, 1, 1, keep=True)
        # ---- parity on the CPU leg's own data: the CUDA step against the oracle's outputs ----
        try:
            line["parity"] = parity_block(held, dev)
        except Exception as exc:
            line["parity"] = {"error": repr(exc)[:300]}
        # ---- the user-visible call: NumPy in, NumPy out, whole EM loop, against the oracle on the same data ------
        try:
            line["e2e_call"] = e2e_call_block(held, dev)
        except Exception as exc:
            line["e2e_call"] = {"error": repr(exc)[:300]}
    if rank == 0:
        emit(out, line)
    if world > 1:
        group.close()
        torch.distributed.destroy_process_group()


def parity_block(held, dev):
    """One CUDA E+M step on the host data the cpu_baseline leg used, compared with the oracle's outputs for that step.
    Tolerances are BASELINE.json's: 1e-9 relative for the FP64 statistics and epsilon, equal iteration count, raw
    posteriors at SURVEY.md H1's max(1e-9, 4 * 2^-53 / mean pi), identical selection masks."""
    import numpy as np
    import torch

    from rlvi_b200 import ops

    X, y, params, ref = held[:4]
    n, d = X.shape
    Xd, yd, pd = (torch.from_numpy(a).to(dev) for a in (X, y, params))
    _, e, _ = ops.loss(ops.LOSS_LOGISTIC_CE, Xd, pd, y=yd, intercept=True, want_losses=False, want_e=True)
    pi, res = ops.fixed_point(None, e_work=e)
    mom = ops.weighted_moments(Xd, pi)
    r = ops.read_result(res)
    m = {k: v.cpu().numpy() for k, v in ops.split_moments(mom, d).items()}
    pih = pi.cpu().numpy()

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

    mean_pi = float(ref["pi"].mean())
    pi_tol = max(1e-9, 4 * 2.0 ** -53 / mean_pi)
    out = {"n": int(n), "d": int(d), "iters": int(r["iters"]), "iters_ref": int(ref["iters"]),
           "iters_equal": bool(r["iters"] == ref["iters"]), "eps_rel": abs(r["eps"] - ref["eps"]) / abs(ref["eps"]),
           "G_rel": rel(m["G"] / m["S0"], ref["G"] / ref["S0"]), "S1_rel": rel(m["S1"] / m["S0"], ref["S1"] / ref["S0"]),
           "pi_rel": rel(pih, ref["pi"]), "pi_tol": pi_tol, "mean_pi": mean_pi,
           "pi_normalised_rel": rel(pih / pih.sum(), ref["pi"] / ref["pi"].sum()),
           "mask_equal": bool(np.array_equal(pih > 0.5 * pih.max(), ref["pi"] > 0.5 * ref["pi"].max())),
           "oracle": "oracle.rlvi_np.em_step_logistic on the cpu_baseline sample"}
    out["within_tolerance"] = bool(out["iters_equal"] and out["eps_rel"] <= 1e-9 and out["G_rel"] <= 1e-9 and
                                   out["S1_rel"] <= 1e-8 and out["pi_rel"] <= pi_tol and out["mask_equal"])
    return out


def e2e_call_block(held, dev):
    """`rlvi.logistic_regression(X_np, y_np, mstep="mm")` -- the call a user of the reference makes: NumPy in, NumPy
    out, the whole EM loop (H2D copy once, then iterate on the device) -- timed against the oracle's restatement of
    the same loop (rlvi.py:92-108 with utils.mm_log_reg) on the same host arrays and all host cores."""
    import numpy as np
    import torch

    from rlvi_b200 import ops, rlvi

    X, y, _, _, th_ref, t_cpu, n_cpu = held
    n, d = X.shape
    rlvi.logistic_regression(X[:4096], y[:4096], mstep="mm")          # warm-up: handles, scratch
    torch.cuda.synchronize()

    def timed(m):
        l0 = ops.launch_count(dev.index or 0)
        t0 = time.perf_counter()
        th = rlvi.logistic_regression(X[:m], y[:m], mstep="mm")
        return th, time.perf_counter() - t0, ops.launch_count(dev.index or 0) - l0

    th_s, t_s, _ = timed(n_cpu)                  # the oracle's sample: same arrays, theta compared
    th, t_gpu, launches = timed(n)               # the whole CPU-leg sample
    t0 = time.perf_counter()
    rlvi.logistic_regression(X, y)               # the reference's DEFAULT M-step route (Newton solve of liblinear's objective)
    t_default = time.perf_counter() - t0
    return {"call": "rlvi.logistic_regression(X_np, y_np, mstep='mm')", "n": int(n), "d": int(d), "seconds": t_gpu,
            "samples_per_s": n / t_gpu, "library_launches": int(launches),
            "h2d_bytes": int(n * (d + 1) * 8), "d2h_bytes": int((d + 1) * 8),
            "oracle_n": int(n_cpu), "oracle_seconds": t_cpu, "oracle_samples_per_s": n_cpu / t_cpu,
            "same_sample_seconds": t_s, "same_sample_samples_per_s": n_cpu / t_s, "speedup_vs_oracle_same_sample": t_cpu / t_s,
            "theta_rel_same_sample": float(np.max(np.abs(th_s - th_ref)) / np.max(np.abs(th_ref))),
            "default_route_seconds": t_default, "default_route_samples_per_s": n / t_default,
            "note": "X, y copied host->device once (pageable NumPy memory), EM iterations on the device, theta copied "
                    "back; the oracle is oracle.rlvi_np.logistic_regression_mm (all host cores) on the first "
                    f"2^{int(np.log2(n_cpu))} rows of the same arrays -- the reference's MM loop has no iteration cap"}


def m_step_extras(X, y, pi, params, mom, d):
    """The pieces of utils.mm_log_reg (utils.py:32-58) that are not part of the E+M step: one gradient pass
    X^T (pi (sigmoid - y)) over this rank's shard (the MM loop runs J of them), and the (d+1) x (d+1) inverse of the
    majoriser built from the step's statistics.  Rank-local (no collective), CUDA events, mean of 5."""
    import torch

    from rlvi_b200 import ops

    dev = X.device
    n = X.shape[0]

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    g = torch.empty(d + 1, dtype=torch.float64, device=dev)
    t_grad = timed(lambda: ops.logistic_grad(X, y, pi, params, out=g))
    m = ops.split_moments(mom, d)
    A = torch.empty((d + 1, d + 1), dtype=torch.float64, device=dev)
    A[0, 0] = m["S0"]
    A[0, 1:] = m["S1"]
    A[1:, 0] = m["S1"]
    A[1:, 1:] = m["G"]
    A = A / A.diagonal().max()                       # scale-free: the collapse regime makes the entries tiny
    t_inv = timed(lambda: torch.linalg.inv(0.25 * A))
    grad_bytes = n * (d * 8 + 8 + 8)
    return {"logistic_grad_pass_ms": t_grad, "logistic_grad_GBps": grad_bytes / (t_grad * 1e-3) / 1e9,
            "logistic_grad_algorithmic_bytes": int(grad_bytes), "majoriser_inverse_ms": t_inv,
            "note": "rlvi_logistic_grad_f64 over this rank's shard (one of the J passes of utils.mm_log_reg) and "
                    "torch.linalg.inv of the (d+1)^2 majoriser on the device; not included in `value`"}


def pin_to_gpu_cpus(index):
    """Multi-GPU runs: restrict this rank to the CPU cores NVML reports as local to its GPU, so that the pinned host
    buffers of the e2e leg are allocated on the GPU's NUMA node (eight ranks copying from one socket's memory share
    its bandwidth and the inter-socket link).  Returns the number of cores, or None when NVML does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_e2e(args, X, y, params, dev, world, rank, n_total, group):
    """Host buffers in, host buffers out: every step copies this rank's X/y shard host->device from pinned
    memory (inside rlvi_em_step_logistic_host), runs the three stages, and copies pi + statistics back."""
    import psutil
    import torch

    from rlvi_b200 import ops

    n, d = X.shape
    need = n * (d + 2) * 8
    avail = psutil.virtual_memory().available // max(world, 1)
    n_e2e = n
    while n_e2e * (d + 2) * 8 > 0.6 * avail and n_e2e > 1024:
        n_e2e //= 2
    Xh = torch.empty((n_e2e, d), dtype=torch.float64, pin_memory=True)
    yh = torch.empty(n_e2e, dtype=torch.float64, pin_memory=True)
    pih = torch.empty(n_e2e, dtype=torch.float64, pin_memory=True)
    Xh.copy_(X[:n_e2e])
    yh.copy_(y[:n_e2e])
    ph = params.cpu()
    torch.cuda.synchronize()
    if world > 1:
        # the host entry point is single-GPU (one shard per process); ranks run their shards side by side
        torch.distributed.barrier()
    steps = max(1, args.e2e_steps)
    if world > 1:     # every rank must run the same shard size for the global mean: agree on the smallest
        t = torch.tensor([n_e2e], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        n_e2e = int(t)
        Xh, yh, pih = Xh[:n_e2e], yh[:n_e2e], pih[:n_e2e]
    kw = dict(pi_out=pih, device=dev.index, group=group, n_global=n_e2e * world)
    ops.em_step_logistic_host(Xh, yh, ph, **kw)      # warm-up (allocates the resident copy)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = ops.em_step_logistic_host(Xh, yh, ph, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt = float(t)
    nm = out["moments"].size
    return {"value": n_e2e * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(n_e2e * (d + 1) * 8 + (d + 1) * 8),
            "d2h_bytes_per_step": int(n_e2e * 8 + nm * 8 + 40), "ms_per_step": dt * 1e3, "steps": steps,
            "n_per_gpu": int(n_e2e), "fixed_point_passes": out["result"]["iters"],
            "note": "rlvi_em_step_logistic_host: pinned host X,y -> device (chunked, overlapped with the loss kernel), "
                    "E-step, statistics, pi + statistics -> host"
                    + ("" if n_e2e == n else f"; host RAM limited the e2e shard to {n_e2e} samples")
                    + ("; N>1: rlvi_em_step_logistic_host_sharded -- global fixed point over NVLink peer windows, "
                       "statistics all-reduced over the same windows" if world > 1 else "")}


def main():
    args = parse_args()
    out = claim_stdout()
    if args.impl == "reference":
        run_reference(args, out)
    else:
        run_b200(args, out)


if __name__ == "__main__":
    main()
