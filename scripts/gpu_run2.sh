#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --timeout=300 -p no:cacheprovider -k "sparse_conv" > gpurun_out/pytest_tc_kernels.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc_kernels.log
grep -E "rel err|passed|failed|Error|error|exit" gpurun_out/pytest_tc_kernels.log | tail -60
timeout 600 python scripts/bench_layers.py 1.0 > gpurun_out/bench_layers_s1.log 2>&1; echo "exit $?" >> gpurun_out/bench_layers_s1.log
cat gpurun_out/bench_layers_s1.log
timeout 900 python -m pytest tests/test_gpu_networks.py -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/pytest_networks.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_networks.log
grep -E "rel err|passed|failed|Error|error|exit|median|step" gpurun_out/pytest_networks.log | tail -40
timeout 600 python scripts/profile_step.py 18000 5 0 > gpurun_out/profile_step_auto.log 2>&1; echo "exit $?" >> gpurun_out/profile_step_auto.log
cat gpurun_out/profile_step_auto.log
