#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/launches.csv
timeout 1200 python -m pytest tests -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps|encoder|refine" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python scripts/bench_layers.py 1.0 > gpurun_out/bench_layers_s1.log 2>&1; echo "exit $?" >> gpurun_out/bench_layers_s1.log
cat gpurun_out/bench_layers_s1.log
timeout 300 python scripts/profile_step.py 18000 5 0 > gpurun_out/profile_step_auto.log 2>&1; cat gpurun_out/profile_step_auto.log
timeout 420 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -12 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 452 -c 4 -o gpurun_out/prof_spconv_tc_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy \
    --clock-control none -k regex:k_spconv_tc -s 420 -c 57 -o gpurun_out/prof_spconv_tc_step -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
ls -la gpurun_out; du -sh gpurun_out
