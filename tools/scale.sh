#!/bin/bash
# scaling run as the driver does it: N = 1, 2, 4, 8 back to back on one box
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 250 -rf -x > gpurun_out/pytest_multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi8.log
timeout 120 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scale_1.log 2>&1
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > gpurun_out/scale_$n.log 2>&1
done
tail -3 gpurun_out/pytest_multi8.log
for n in 1 2 4 8; do tail -1 gpurun_out/scale_$n.log | cut -c1-200; done
