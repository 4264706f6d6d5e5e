#!/usr/bin/env bash
# fused head MLP (lb2_head_mlp): GPU suite, bench, and what runs between the last conv of a step and the first convs of the next
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_head$i.json 2> gpurun_out/bench_head$i.err; python -c "
import json;d=json.load(open('gpurun_out/bench_head$i.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"; done
TIMELINE_HEAD=150 timeout -k 10 600 python scripts/profile_timeline.py 10 3 2>&1 | grep -v Warn > gpurun_out/timeline_head.log; grep -A150 "activities around" gpurun_out/timeline_head.log | head -152
