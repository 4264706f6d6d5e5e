#!/usr/bin/env bash
# Morton A/B on one box + ncu --set full with source of the L3 layers up1.1.0.net.0 (384->256) and up1.1.0.net.3 (256->256)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],json.dumps(d['roofline']['by_class']),d['roofline']['all_conv']['gather_scatter_model_GBps'])"
}
run m0 LB2_TC_PAIR=2 LB2_MORTON_LEVELS=
run m234 LB2_TC_PAIR=2 LB2_MORTON_LEVELS=234
run m0b LB2_TC_PAIR=2 LB2_MORTON_LEVELS=
run m34 LB2_TC_PAIR=2 LB2_MORTON_LEVELS=34
timeout -k 10 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 12 -c 3 -o gpurun_out/prof_pair_l3 -f \
    env LB2_TC_PAIR=2 LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_pair_l3.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
