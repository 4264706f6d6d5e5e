#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_edge.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_edge.log 2>&1
echo "edge exit $?" >> gpurun_out/pytest_edge.log; tail -25 gpurun_out/pytest_edge.log
