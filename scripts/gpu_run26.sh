#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_kernels.py 25 2 > gpurun_out/profile_kernels_25.log 2>&1; tail -45 gpurun_out/profile_kernels_25.log
timeout 200 python scripts/profile_kernels.py 0 2 > gpurun_out/profile_kernels_0.log 2>&1; head -30 gpurun_out/profile_kernels_0.log
