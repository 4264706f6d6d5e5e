#!/usr/bin/env bash
# pass-major order + evict-first epilogue stores: A/B on one box, GPU tests, then per-launch metrics of one step
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],json.dumps(d['roofline']['by_class']),d['roofline']['all_conv']['gather_scatter_model_GBps'])"
}
run cs0 LB2_TC_PAIR=2 LB2_STREAM_STORES=0
run cs1 LB2_TC_PAIR=2 LB2_STREAM_STORES=1
run cs0b LB2_TC_PAIR=2 LB2_STREAM_STORES=0
run cs1b LB2_TC_PAIR=2 LB2_STREAM_STORES=1
run cs1p1 LB2_TC_PAIR=1 LB2_STREAM_STORES=1
timeout -k 10 600 ncu --profile-from-start off --clock-control none -k regex:k_spconv -c 49 \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,lts__t_sector_hit_rate.pct \
  --csv --log-file gpurun_out/conv_metrics_step0.csv env LB2_TC_PAIR=2 LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_metrics.log 2>&1; echo "ncu metrics exit $?"
