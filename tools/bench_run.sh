#!/bin/bash
# usage: tools/bench_run.sh <tag> [bench flags]  -- one plain bench.py run on the GPU box, wall time recorded, summary printed
cd "$GRAFT_REPO_ROOT"
TAG=$1; shift
S=$(date +%s)
timeout 850 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "rc=$? wall=$(( $(date +%s) - S ))s"
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
lines = open(f"gpurun_out/{tag}_bench.json").read().strip().splitlines()
d = json.loads(lines[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d.get("e2e", {}).get("value"))
print("roofline", {k: d["roofline"][k] for k in ("frac", "step_frac", "frac_nominal", "step_frac_nominal")})
print("config3", json.dumps(d.get("configs", {}).get("config3_pca_fp32")))
print("parity", json.dumps(d.get("parity")))
PY
