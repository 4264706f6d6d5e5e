#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/launches.csv
timeout 1200 python -m pytest tests -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps|encoder|refine|scatter split" gpurun_out/pytest_gpu.log | tail -40
timeout 300 python scripts/profile_step.py 18000 5 0 > gpurun_out/profile_step_auto.log 2>&1; cat gpurun_out/profile_step_auto.log
timeout 420 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -12 gpurun_out/bench_n1.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spconv -s 470 -c 140 --csv \
    --log-file gpurun_out/conv_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
ls -la gpurun_out; du -sh gpurun_out
