#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "persistent or fused" > gpurun_out/pytest_gate.log 2>&1
echo "gate exit $?" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
