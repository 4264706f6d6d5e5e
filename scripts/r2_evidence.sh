#!/usr/bin/env bash
# round-2 evidence on one B200: GPU suite, smoke, bench at the driver's settings and over the full schedule, reference arm, T=1000,
# ncu launch list of one timed step, per-launch conv metrics, ncu --set full of the dominant kernel on the level-3 layers
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
nvidia-smi > gpurun_out/env.txt; lscpu | head -20 >> gpurun_out/env.txt
( time timeout -k 10 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider -s ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.stderr.log ) 2>&1 | tail -3; echo "bench exit $?"
cut -c1-260 gpurun_out/bench_n1.json
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_steps50.json 2> gpurun_out/bench_n1_steps50.stderr.log; echo "bench50 exit $?"
cut -c1-200 gpurun_out/bench_n1_steps50.json
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.stderr.log ) 2>&1 | tail -3
cut -c1-400 gpurun_out/bench_reference_arm.json
LB2_GRAPHS=0 timeout 900 python bench.py --gpus 1 --T 1000 --steps 20 --warmup 3 --no-cpu-baseline --no-fixed --no-scan > gpurun_out/bench_T1000.json 2> gpurun_out/bench_T1000.stderr.log; echo "T1000 exit $?"
cut -c1-200 gpurun_out/bench_T1000.json
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launch_list_step.csv \
    env LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout -k 10 600 ncu --profile-from-start off --clock-control none -k regex:k_spconv -c 49 \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,lts__t_sector_hit_rate.pct \
  --csv --log-file gpurun_out/ncu_conv_launch_metrics.csv env LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_metrics.log 2>&1; echo "ncu metrics exit $?"
timeout -k 10 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc_pair -s 12 -c 3 -o gpurun_out/prof_pair_l3 -f \
    env LB2_GRAPHS=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fixed --no-scan --profiler-range > gpurun_out/ncu_pair_l3.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -20; du -sh gpurun_out
