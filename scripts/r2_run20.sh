#!/usr/bin/env bash
# shared flag-driven epilogue (tc_common.cuh: epilogue_slabs): oracle cases, GPU suite, then A/B of the library before / after inside one box
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_conv_pair.py -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pair_tests.log 2>&1
echo "pair tests exit $?" >> gpurun_out/pair_tests.log
tail -3 gpurun_out/pair_tests.log
if grep -q "pair tests exit 0" gpurun_out/pair_tests.log; then
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() {  # name, library
  name=$1; cp lidiff_b200/_C/ab/$2.so lidiff_b200/_C/liblidiff_b200.so
  timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],json.dumps(d['roofline']['by_class']))"
}
run new1 new
run base1 base
run new2 new
run base2 base
cp lidiff_b200/_C/ab/new.so lidiff_b200/_C/liblidiff_b200.so
fi
