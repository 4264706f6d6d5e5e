#!/usr/bin/env bash
# final round-1 check: full GPU suite, smoke, bench with CPU baseline
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit" gpurun_out/pytest_gpu.log | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_n1.json
timeout 200 python scripts/profile_kernels.py 0 2 > gpurun_out/profile_kernels_0.log 2>&1; grep -v Warn gpurun_out/profile_kernels_0.log | head -14
