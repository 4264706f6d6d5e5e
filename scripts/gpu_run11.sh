#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_gpu.log | tail -12
LB2_TC_NSPLIT=0 timeout 600 python scripts/profile_layers.py 0 > gpurun_out/profile_layers_nosplit.log 2>&1
grep -E "===|conv total|up1.1|stage4" gpurun_out/profile_layers_nosplit.log
timeout 600 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers.log 2>&1; echo "exit $?" >> gpurun_out/profile_layers.log
grep -E "===|conv total|up1.1|stage4" gpurun_out/profile_layers.log
timeout 420 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json | cut -c1-300; tail -3 gpurun_out/bench_n1.err
