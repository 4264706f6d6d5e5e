#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
cut -c1-260 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
