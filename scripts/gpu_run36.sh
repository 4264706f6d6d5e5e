#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_networks.py -m gpu -q -x -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_net.log 2>&1
rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_net.log; grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_net.log | tail -6
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_net.log; exit 1; fi
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_side1.json 2> gpurun_out/bench_side1.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_side1.json
LB2_SIDE_STREAM=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_side0.json 2> gpurun_out/bench_side0.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_side0.json
