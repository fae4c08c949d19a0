#!/bin/bash
# Final-round captures of the frequency-domain sampler kernel with speculative evaluation (run under gpurun).
set -x
TAG=${1:-r01_final}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:rjmcmc -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_chain_benchsize_metrics.csv \
    python scripts/profile_chain.py 4096 0 > gpurun_out/${TAG}_chain_benchsize.log 2>&1
ls -la gpurun_out | tail -6
