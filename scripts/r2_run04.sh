#!/usr/bin/env bash
# ncu --set full capture of the CTA-pair kernel on the step-0 geometry: up1.1.0.net.0 (384->256), its 1x1 downsample, up1.1.0.net.3, up1.1.1.net.0
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 589 -c 4 -o gpurun_out/prof_pair_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed > gpurun_out/ncu_pair_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/ncu_pair_full.log
ls -la gpurun_out/*.ncu-rep
