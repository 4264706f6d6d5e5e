#!/usr/bin/env bash
# pair kernel (cta_group::2) bring-up: parity vs the oracle, then the whole GPU suite, then step time per kernel selection
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_conv_pair.py -q -x -s --timeout=120 -p no:cacheprovider > gpurun_out/pair_tests.log 2>&1
echo "pair tests exit $?" >> gpurun_out/pair_tests.log
grep -E "pair=|passed|failed|Error|exit" gpurun_out/pair_tests.log | tail -50
if grep -q "pair tests exit 0" gpurun_out/pair_tests.log; then
  timeout -k 10 700 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -4 gpurun_out/pytest_gpu.log
  for pm in 0 1 2; do
    LB2_TC_PAIR=$pm timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed > gpurun_out/bench_pair$pm.json 2> gpurun_out/bench_pair$pm.err
    echo "pair=$pm exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_pair$pm.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],json.dumps(d['roofline']['by_class']),json.dumps(d['roofline']['all_conv']))"
  done
fi
