#!/usr/bin/env bash
# ncu --set full of small-Cout conv launches (L1 32-ch and L0 96-ch) to find what bounds them
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_small -s 2 -c 2 -o gpurun_out/prof_small_l1 -f \
    env LB2_TC_PAIR=2 LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_small_l1.log 2>&1; echo "ncu l1 exit $?"
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_small -s 22 -c 3 -o gpurun_out/prof_small_l0 -f \
    env LB2_TC_PAIR=2 LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_small_l0.log 2>&1; echo "ncu l0 exit $?"
ls -la gpurun_out/*.ncu-rep
