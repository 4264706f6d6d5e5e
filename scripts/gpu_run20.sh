#!/usr/bin/env bash
# full ncu capture of the 48 tensor-core conv launches of one denoising step; export raw + source pages of selected layers
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_spconv_tc -s 336 -c 48 -o /tmp/step -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; echo "ncu exit $?"
ls -la /tmp/step.ncu-rep
ncu -i /tmp/step.ncu-rep --page raw --csv > gpurun_out/step_convs_raw.csv 2>/dev/null
for id in 0 2 13 21 25 28 40 47; do
  ncu -i /tmp/step.ncu-rep --page source --csv --launch-skip $id --launch-count 1 > gpurun_out/step_conv_src_$id.csv 2>/dev/null
done
ls -la gpurun_out | head -30; du -sh gpurun_out
