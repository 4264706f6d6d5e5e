#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -p no:cacheprovider -k "persistent or fused or split or conv" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc" >> gpurun_out/pytest_gate.log; grep -E "passed|failed|Error|exit" gpurun_out/pytest_gate.log | tail -5
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_gate.log; exit 1; fi
timeout 200 python scripts/profile_layers.py 0 25 > gpurun_out/profile_layers_r32.log 2>&1
grep -E "===|conv total|up1.1|stage3" gpurun_out/profile_layers_r32.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r32.json 2> gpurun_out/bench_r32.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_r32.json
