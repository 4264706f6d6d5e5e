#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "nn_match" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_kernels.py 25 2 > gpurun_out/profile_kernels_25.log 2>&1; grep -v Warn gpurun_out/profile_kernels_25.log | head -8
timeout 200 python scripts/profile_kernels.py 0 2 > gpurun_out/profile_kernels_0.log 2>&1; grep -v Warn gpurun_out/profile_kernels_0.log | head -8
timeout 700 python -m pytest tests -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_gpu.log | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cut -c1-300 gpurun_out/bench_n1.json
