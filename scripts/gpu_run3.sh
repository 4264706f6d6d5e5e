#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "rel err|passed|failed|Error|exit|median|step " gpurun_out/pytest_gpu.log | tail -70
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 150 -c 3 -o gpurun_out/prof_spconv_tc -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
