#!/usr/bin/env bash
# lean activations + pair kernel: GPU suite (incl. full-size goldens), bench LB2_LEAN=0/1, then ncu --set full of the pair kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "worst|survive|passed|failed|Error|exit" gpurun_out/pytest_gpu.log | tail -12
for lean in 0 1; do
  LB2_LEAN=$lean LB2_TC_PAIR=2 timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed > gpurun_out/bench_lean$lean.json 2> gpurun_out/bench_lean$lean.err
  echo "lean=$lean exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_lean$lean.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],json.dumps(d['roofline']['by_class']),json.dumps(d['roofline']['all_conv']))"
done
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 589 -c 4 -o gpurun_out/prof_pair_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed > gpurun_out/ncu_pair_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/ncu_pair_full.log
ls -la gpurun_out/*.ncu-rep
