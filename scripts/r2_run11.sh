#!/usr/bin/env bash
# deeper gather ring for Cout <= 96, epilogue operand prefetch: GPU suite, bench x2, kernel time table of steps 0 and 25
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
run a LB2_TC_PAIR=2
run b LB2_TC_PAIR=2
run g0 LB2_TC_PAIR=2 LB2_GRAPHS=0
LB2_TC_PAIR=2 LB2_GRAPHS=0 timeout 300 python scripts/profile_kernels.py 0 2 > gpurun_out/profile_kernels_0.log 2>&1; grep -v Warn gpurun_out/profile_kernels_0.log | head -32
LB2_TC_PAIR=2 LB2_GRAPHS=0 timeout 300 python scripts/profile_kernels.py 25 2 > gpurun_out/profile_kernels_25.log 2>&1; grep -v Warn gpurun_out/profile_kernels_25.log | head -12
