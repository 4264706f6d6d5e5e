#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc_n256 -s 54 -c 4 -o gpurun_out/prof_spconv_n256_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
