#!/usr/bin/env bash
# kernel maps of levels 2-4 on a side stream behind the stem / stage-1 convolutions: GPU suite + A/B
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],json.dumps(d['roofline']['by_class']),d['roofline']['all_conv']['gather_scatter_model_GBps'])"
}
run lm1
run lm0 LB2_LATE_MAPS=0
run lm1b
run lm0b LB2_LATE_MAPS=0
run lm1g0 LB2_GRAPHS=0
