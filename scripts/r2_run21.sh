#!/usr/bin/env bash
# what the step spends outside the conv kernels on its critical path (CUPTI timeline, graphs on)
mkdir -p gpurun_out
timeout -k 10 600 python scripts/profile_timeline.py 10 3 > gpurun_out/timeline_step10.log 2>&1; echo "exit $?"; tail -60 gpurun_out/timeline_step10.log
