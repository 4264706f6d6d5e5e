#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "nn_match or row_order" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_kernels.py 25 2 > gpurun_out/profile_kernels_25.log 2>&1; grep -v Warn gpurun_out/profile_kernels_25.log | head -12
timeout 200 python scripts/profile_kernels.py 0 2 > gpurun_out/profile_kernels_0.log 2>&1; grep -v Warn gpurun_out/profile_kernels_0.log | head -12
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r29.json 2> gpurun_out/bench_r29.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_r29.json
