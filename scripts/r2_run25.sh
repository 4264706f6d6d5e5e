#!/usr/bin/env bash
# step start: stem right behind the level-0 map (LB2_LATE_FROM=1), NN matches in the order the gates need them (LB2_NN_SPLIT=1): GPU suite, A/B, timeline
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
run new
run old LB2_NN_SPLIT=0 LB2_LATE_FROM=2
run nn_only LB2_LATE_FROM=2
run late_only LB2_NN_SPLIT=0
run new2
run old2 LB2_NN_SPLIT=0 LB2_LATE_FROM=2
TIMELINE_HEAD=40 timeout -k 10 600 python scripts/profile_timeline.py 10 3 2>&1 | grep -v Warn > gpurun_out/timeline_head.log; grep "NN kernels\|ms/step;\|conv kernels" gpurun_out/timeline_head.log
