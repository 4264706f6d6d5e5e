#!/usr/bin/env bash
# first GPU session: environment facts, parity tests, step timing
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  echo "host cores: $(nproc)"; lscpu | grep -E "Model name|Socket|Thread|Core" | head -5
  free -g | head -2
} > gpurun_out/env.txt 2>&1
python -m pytest tests -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python scripts/profile_step.py 18000 3 1 > gpurun_out/profile_step.log 2>&1
echo "profile exit $?" >> gpurun_out/profile_step.log
cat gpurun_out/profile_step.log
