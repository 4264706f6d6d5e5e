#!/usr/bin/env bash
# ncu --set full of the level-0 decoder launches of k_spconv_tc_small<3> (up4: deconv 96->96, 128->96, 1x1 128->96, 96->96 +pre_add, 96->96, 96->96 +residual)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -k 10 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_small -s 19 -c 6 -o gpurun_out/prof_small_l0_r2 -f \
    env LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_small_l0.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
