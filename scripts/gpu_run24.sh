#!/usr/bin/env bash
mkdir -p gpurun_out
LB2_PROFILE_WASTE=1 LB2_SCATTER_LEVELS=x timeout 200 python scripts/profile_layers.py 0 25 49 > gpurun_out/profile_layers_noscatter.log 2>&1
grep -E "===|conv total|waste" gpurun_out/profile_layers_noscatter.log
LB2_SCATTER_LEVELS=x timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noscatter.json 2> gpurun_out/bench_noscatter.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_noscatter.json
