#!/usr/bin/env bash
# graphs A/B with the parity fix, whole-scan mode, and per-launch ncu metrics of every conv of one step (step-0 geometry)
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('  ',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],json.dumps(d['roofline']['by_class']),d['roofline']['all_conv']['gather_scatter_model_GBps'])"
}
run p2l1g0 LB2_TC_PAIR=2 LB2_LEAN=1 LB2_GRAPHS=0
run p2l1g1 LB2_TC_PAIR=2 LB2_LEAN=1 LB2_GRAPHS=1
run p2l1g0b LB2_TC_PAIR=2 LB2_LEAN=1 LB2_GRAPHS=0
run p2l1g1b LB2_TC_PAIR=2 LB2_LEAN=1 LB2_GRAPHS=1
LB2_TC_PAIR=2 timeout -k 10 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_full.json'));print(d['value'],d['e2e'],d['scan_e2e'],d['fixed_geometry'],d['cpu_baseline'],d['engine'])"
timeout -k 10 600 ncu --profile-from-start off --clock-control none -k regex:k_spconv -c 49 \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,lts__t_sector_hit_rate.pct \
  --csv --log-file gpurun_out/conv_metrics_step0.csv env LB2_TC_PAIR=2 LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_metrics.log 2>&1; echo "ncu metrics exit $?"
tail -2 gpurun_out/ncu_metrics.log | cut -c1-300
