#!/bin/bash
# Final single-GPU evidence pass (round 2): full GPU suite, bench lines (no profiler), then the ncu launch list + DRAM traffic
# at full size and one --set full capture of the three step kernels at N = 2^23 (ncu replays each kernel ~40 times).
cd "$GRAFT_REPO_ROOT"
R=r02h
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${R}_clocks.csv & SMI_PID=$!
timeout 400 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
kill $SMI_PID
timeout 200 python bench.py --impl reference > gpurun_out/${R}_bench_ref.json 2>&1; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
$CMD > gpurun_out/${R}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"loss_tma_kernel|fp_kernel|gram64" -c 16 --csv --log-file gpurun_out/${R}_launches_traffic_n26.csv $CMD > gpurun_out/${R}_ncu1.log 2>&1; echo "ncu1 rc=$?"
CMD2="python bench.py --log2n 23 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs"
$CMD2 > gpurun_out/${R}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"loss_tma_kernel|fp_kernel|gram64_kernel" -s 3 -c 3 -o gpurun_out/${R}_prof_n23 $CMD2 > gpurun_out/${R}_ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -c 400 gpurun_out/${R}_bench_n1.json
