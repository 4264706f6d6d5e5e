#!/usr/bin/env bash
# one point of the scaling curve: N ranks on one box, driver's launch line
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.stderr.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.stderr.log
fi
echo "bench N=$N exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'scan',d.get('scan_e2e'),d['clocks'])"
tail -5 gpurun_out/bench_n$N.stderr.log
