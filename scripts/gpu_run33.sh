#!/usr/bin/env bash
# round-1 evidence: full GPU suite, smoke, bench (both arms), ncu launch list + full capture of the dominant kernel + conv DRAM traffic
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/launches.csv gpurun_out/tc3_src_*.csv gpurun_out/step_conv_src_*.csv gpurun_out/step_convs_raw.csv
timeout 700 python -m pytest tests -m gpu -q -s --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|exit" gpurun_out/pytest_gpu.log | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cut -c1-200 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
cut -c1-300 gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc_n256 -s 84 -c 5 -o gpurun_out/prof_spconv_n256_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:k_spconv -s 300 -c 98 --csv --log-file gpurun_out/conv_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
ls -la gpurun_out | head -40; du -sh gpurun_out
