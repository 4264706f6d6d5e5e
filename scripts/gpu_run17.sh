#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "persistent or nn_match or fused" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers.log 2>&1
grep -E "===|conv total|up1.1" gpurun_out/profile_layers.log
timeout 600 python -m pytest tests -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_gpu.log | tail -8
timeout 200 python scripts/profile_step.py 18000 5 0 > gpurun_out/profile_step_auto.log 2>&1; tail -5 gpurun_out/profile_step_auto.log
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json | cut -c1-300; tail -3 gpurun_out/bench_n1.err
