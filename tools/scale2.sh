#!/bin/bash
# usage: tools/scale2.sh <gpus> <log2n>  -- multi-GPU parity tests, then one bench line (device-resident only) at <gpus> GPUs
cd "$GRAFT_REPO_ROOT"
G=$1; L=$2
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --log2n $L --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2> gpurun_out/scale2_${G}_${L}.err | tail -1 > gpurun_out/scale2_${G}_${L}.json
python - "$G" "$L" <<'PY'
import json, sys
g, l = sys.argv[1:3]
d = json.loads(open(f"gpurun_out/scale2_{g}_{l}.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["config"]["n_total"], "ms/step", d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()},
      d["fixed_point"], d.get("fixed_k32"))
PY
