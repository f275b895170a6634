"""bench.py's contract, checked on the CPU: the reference arm prints one JSON line with the agreed keys, and the
B200 arm refuses to run (non-zero exit, no JSON) when there is no GPU -- it never falls back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--cpu-log2n", "14", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "rlvi_em_step_samples_per_sec" and j["unit"] == "samples/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["value"] > 0 and j["steps"] == 2 and j["warmup"] == 1 and j["gpu_launches"] == 0
    assert "workload" in j["config"] and "N=2^26" in j["config"]["workload"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "N=2^14" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    r = run(["--impl", "reference", "--gpus", "2", "--cpu-log2n", "12", "--steps", "1", "--warmup", "0"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
    r0 = run(["--impl", "reference", "--gpus", "2", "--cpu-log2n", "12", "--steps", "1", "--warmup", "0"],
             env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r0.returncode == 0 and len([l for l in r0.stdout.splitlines() if l.startswith("{")]) == 1


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(["--steps", "1", "--warmup", "0", "--log2n", "12", "--no-e2e", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_both_arms_report_the_same_config_keys():
    """The driver compares the two arms: same metric / unit / workload string."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("workload_config(args") >= 2        # both arms build `config` from the same function


def test_stdout_carries_only_the_json_line():
    """Libraries that write to fd 1 (NCCL's version banner under NCCL_DEBUG) must not reach stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "out = bench.claim_stdout()\n"
            "os.write(1, b'NCCL version x.y\\n'); print('python noise')\n"
            "bench.emit(out, {'metric': 'm', 'value': 1})\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "m", "value": 1}\n'
    assert "NCCL version x.y" in r.stderr and "python noise" in r.stderr
    r = run(["--impl", "reference", "--cpu-log2n", "12", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0 and len(r.stdout.splitlines()) == 1 and r.stdout.startswith("{")


def test_m_step_extras_report(monkeypatch):
    """bench.py's separately-timed M-step pieces: keys, byte count, and that they are computed
    from the step's own statistics buffer -- run with the kernels replaced by the test double and mocked CUDA events."""
    sys.path.insert(0, ROOT)
    import abi_double
    import bench
    from rlvi_b200 import ops

    class Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 10.0

    abi_double.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    n, d = 500, 8
    g = torch.Generator().manual_seed(0)
    X = torch.randn(n, d, dtype=torch.float64, generator=g)
    y = (torch.rand(n, generator=g) < 0.5).double()
    pi = torch.rand(n, dtype=torch.float64, generator=g) * 1e-7          # collapse-regime magnitudes
    out = bench.m_step_extras(X, y, pi, torch.zeros(d + 1, dtype=torch.float64), ops.weighted_moments(X, pi), d)
    assert out["logistic_grad_pass_ms"] == 2.0 and out["majoriser_inverse_ms"] == 2.0      # 10 ms / 5 repetitions
    assert out["logistic_grad_algorithmic_bytes"] == n * (d * 8 + 16)
    assert abs(out["logistic_grad_GBps"] - n * (d * 8 + 16) / 2e-3 / 1e9) < 1e-9
