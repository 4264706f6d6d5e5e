#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_layers.py 0 25 49 > gpurun_out/profile_layers.log 2>&1; echo "exit $?" >> gpurun_out/profile_layers.log
cat gpurun_out/profile_layers.log
