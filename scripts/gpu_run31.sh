#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_spconv_tc_n256 -s 60 -c 6 -o /tmp/tc3 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc3.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/tc3.ncu-rep --page raw --csv > gpurun_out/tc3_raw.csv 2>/dev/null
for id in 0 1 2 3 4 5; do ncu -i /tmp/tc3.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > gpurun_out/tc3_src_$id.csv 2>/dev/null; done
ls -la gpurun_out | grep tc3
