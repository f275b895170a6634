#!/bin/bash
# Final scaling run on one 8-GPU box: N = 1 (device-resident only) then N = 2, 4, 8 with the driver's default flags.
cd "$GRAFT_REPO_ROOT"
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scalef_1.log 2>&1
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scalef_$n.log 2>&1
  echo "N=$n rc=$?"
done
for n in 1 2 4 8; do tail -1 gpurun_out/scalef_$n.log | cut -c1-160; done
