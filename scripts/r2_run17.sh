#!/usr/bin/env bash
# ncu --set full of two Cout-128 pair launches: stage3.1.net.3 (128->128, level 3) and up2.1.0.net.0 (192->128, level 2)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 2 -c 1 -o gpurun_out/prof_pair128_l3 -f \
    env LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_p128a.log 2>&1; echo "ncu a exit $?"
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 18 -c 1 -o gpurun_out/prof_pair128_l2 -f \
    env LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_p128b.log 2>&1; echo "ncu b exit $?"
ls -la gpurun_out/*.ncu-rep
