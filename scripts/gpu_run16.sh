#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "persistent or sparse_conv or row_order" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers.log 2>&1
grep -E "===|conv total" gpurun_out/profile_layers.log
