#!/usr/bin/env bash
mkdir -p gpurun_out
# gate: risky new kernels first, short leash; stop the whole call if they hang or fail
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "persistent or farthest" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_gate.log; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_gpu.log | tail -12
timeout 200 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers.log 2>&1; echo "exit $?" >> gpurun_out/profile_layers.log
grep -E "===|conv total" gpurun_out/profile_layers.log
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json | cut -c1-300; tail -3 gpurun_out/bench_n1.err
