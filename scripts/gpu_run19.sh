#!/usr/bin/env bash
# tc3 (N=256 register-total kernel) validation: gate on the equality tests, then A/B per-layer and bench
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "persistent or fused or split" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers_n256.log 2>&1
grep -E "===|conv total|up1.1|stage4.1" gpurun_out/profile_layers_n256.log
LB2_TC_N256=0 timeout 200 python scripts/profile_layers.py 0 49 > gpurun_out/profile_layers_base.log 2>&1
grep -E "===|conv total|up1.1|stage4.1" gpurun_out/profile_layers_base.log
timeout 500 python -m pytest tests/test_gpu_networks.py -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_net.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_net.log
grep -E "passed|failed|Error|exit|guided eps" gpurun_out/pytest_net.log | tail -8
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n256.json 2> gpurun_out/bench_n256.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_n256.json
