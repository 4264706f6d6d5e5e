#!/usr/bin/env bash
# tc4 (Cout <= 128, two drain warpgroups) validation
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "persistent or fused or split or conv" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers_r23.log 2>&1
grep -E "===|conv total" gpurun_out/profile_layers_r23.log
LB2_TC_SMALL=0 timeout 200 python scripts/profile_layers.py 0 > gpurun_out/profile_layers_r23_base.log 2>&1
grep -E "===|conv total" gpurun_out/profile_layers_r23_base.log
timeout 500 python -m pytest tests/test_gpu_networks.py -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_net.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_net.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_net.log | tail -8
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r23.json 2> gpurun_out/bench_r23.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_r23.json
