#!/usr/bin/env bash
# pair kernel after removing the GPU-scope membars: parity, then step time per kernel selection
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_conv_pair.py -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pair_tests.log 2>&1
echo "pair tests exit $?" >> gpurun_out/pair_tests.log
tail -3 gpurun_out/pair_tests.log
if grep -q "pair tests exit 0" gpurun_out/pair_tests.log; then
  for pm in 0 1 2; do
    LB2_TC_PAIR=$pm timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed > gpurun_out/bench_pair$pm.json 2> gpurun_out/bench_pair$pm.err
    echo "pair=$pm exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_pair$pm.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],json.dumps(d['roofline']['by_class']),json.dumps(d['roofline']['all_conv']))"
  done
fi
