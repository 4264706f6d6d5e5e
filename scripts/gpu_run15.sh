#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "scatter" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit|scatter split" gpurun_out/pytest_gate.log | tail -8
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_gate.log; exit 1; fi
for lv in "2" "012" "x"; do
  LB2_SCATTER_LEVELS=$lv timeout 200 python scripts/profile_layers.py 0 > gpurun_out/profile_layers_sc$lv.log 2>&1
  echo "scatter levels=$lv"; grep -E "===|conv total|up2.1.1" gpurun_out/profile_layers_sc$lv.log
done
