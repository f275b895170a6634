#!/bin/bash
# Final single-GPU evidence pass: bench lines (no profiler), then the ncu launch list + traffic at full size and
# one --set full capture of the three step kernels at N = 2^23 (ncu replays each kernel ~40 times).
cd "$GRAFT_REPO_ROOT"
timeout 400 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2>&1; echo "ref rc=$?"
timeout 300 python tools/bench_configs.py > gpurun_out/final_configs.jsonl 2>&1; echo "configs rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"loss_tma_kernel|fp_kernel|gram64" -c 16 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
CMD2="python bench.py --log2n 23 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"loss_tma_kernel|fp_kernel|gram64_kernel" -s 3 -c 3 -o gpurun_out/prof_final $CMD2 > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -c 600 gpurun_out/final_bench_n1.json
