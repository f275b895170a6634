#!/bin/bash
# 8-GPU evidence run (second half of round 2): multi-GPU parity tests, the default bench line at N = 8, N = 4 device-resident
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02g_scale_8.json 2> gpurun_out/r02g_scale_8.err
echo "N=8 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02g_scale_4.json 2> gpurun_out/r02g_scale_4.err
echo "N=4 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f"gpurun_out/r02g_scale_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", d["ms_per_step"], "value", d["value"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()},
              d["fixed_point"], "k32", d.get("fixed_k32", {}).get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as ex:
        print(n, "failed", ex)
PY
