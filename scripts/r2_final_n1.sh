#!/usr/bin/env bash
# final N = 1 lines at HEAD after the evidence run (bench.py's CPU leg got a warm-up step; the DRAM-traffic table was regenerated for these kernel sources)
mkdir -p gpurun_out
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.stderr.log ) 2>&1 | tail -3; echo "bench exit $?"
cut -c1-260 gpurun_out/bench_n1.json
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.stderr.log ) 2>&1 | tail -3
cut -c1-300 gpurun_out/bench_reference_arm.json
