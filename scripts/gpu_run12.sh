#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout=150 --timeout-method=thread -p no:cacheprovider -k "persistent" > gpurun_out/pytest_persist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_persist.log
grep -E "passed|failed|Error|exit|Timeout" gpurun_out/pytest_persist.log | tail -8
timeout 900 python -m pytest tests -m gpu -q -s --timeout=300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers.log 2>&1; echo "exit $?" >> gpurun_out/profile_layers.log
grep -E "===|conv total" gpurun_out/profile_layers.log
timeout 420 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json | cut -c1-300; tail -3 gpurun_out/bench_n1.err
