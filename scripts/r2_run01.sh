#!/usr/bin/env bash
# round-2 baseline at HEAD: GPU suite, smoke, bench at driver settings
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 700 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
( time timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_s20.json 2> gpurun_out/bench_n1_s20.err ) 2>&1 | tail -3; echo "bench exit $?"
cat gpurun_out/bench_n1_s20.json
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_s20.json 2> gpurun_out/bench_ref_s20.err ) 2>&1 | tail -3
cat gpurun_out/bench_ref_s20.json
