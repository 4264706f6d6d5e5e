#!/usr/bin/env bash
# cost-sorted tile order (static LPT) + gate tables on a third stream: GPU suite, bench A/B
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
run to1
run to0 LB2_TILE_ORDER=0
run to1b
run to0b LB2_TILE_ORDER=0
run s0 LB2_SIDE_STREAM=0
