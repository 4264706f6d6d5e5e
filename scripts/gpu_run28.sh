#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_kernels.py 25 2 > gpurun_out/profile_kernels_25.log 2>&1; grep -v Warn gpurun_out/profile_kernels_25.log | head -8
timeout 200 python scripts/profile_kernels.py 0 2 > gpurun_out/profile_kernels_0.log 2>&1; grep -v Warn gpurun_out/profile_kernels_0.log | head -8
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r28.json 2> gpurun_out/bench_r28.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_r28.json
