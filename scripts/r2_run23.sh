#!/usr/bin/env bash
# NN tree search with 32 x 32 -> 64 bit multiplies and 16-byte node loads: NN tests, per-level search times, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "nn or match or engine or step" -p no:cacheprovider 2>&1 | tail -3
timeout -k 10 600 python scripts/profile_timeline.py 10 3 2>&1 | grep -v Warn | grep "NN kernels\|ms/step;\|conv kernels"
timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_nn.json 2> gpurun_out/bench_nn.err; python -c "
import json;d=json.load(open('gpurun_out/bench_nn.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
