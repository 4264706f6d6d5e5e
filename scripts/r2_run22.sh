#!/usr/bin/env bash
# NN matches in the order the gates need them (4, 0, 1, 2, 3; each from its level-4 ancestor, LB2_NN_SPLIT): GPU suite, A/B inside one box, timeline
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
}
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k compose -p no:cacheprovider 2>&1 | tail -3
run s1 LB2_NN_SPLIT=1
run s0 LB2_NN_SPLIT=0
run s1b LB2_NN_SPLIT=1
run s0b LB2_NN_SPLIT=0
timeout -k 10 600 python scripts/profile_timeline.py 10 3 > gpurun_out/timeline_step10.log 2>&1; echo "exit $?"; tail -32 gpurun_out/timeline_step10.log
