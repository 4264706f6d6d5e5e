#!/usr/bin/env bash
mkdir -p gpurun_out
for lv in "012" "2" "12" "x"; do
  LB2_SCATTER_LEVELS=$lv timeout 200 python scripts/profile_layers.py 0 > gpurun_out/profile_layers_sc$lv.log 2>&1
  echo "scatter levels=$lv"; grep -E "===|conv total" gpurun_out/profile_layers_sc$lv.log
done
