#!/bin/bash
# ncu captures of the time-domain sampler kernel (run under gpurun).  Writes into gpurun_out/.
set -x
TAG=${1:-r01_tdem}
# 1. launch list of the skytem bench command (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --workload skytem --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# 2. DRAM traffic + key counters at bench size (4096 chains to termination)
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:rjmcmc -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_chain_benchsize_metrics.csv \
    python scripts/profile_chain.py 4096 0 32 skytem > gpurun_out/${TAG}_chain_benchsize.log 2>&1
# 3. full-set capture (with source) on a bounded run: one wave of 148 x 16 chains x 1000 iterations
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rjmcmc -s 1 -c 1 -o gpurun_out/${TAG}_chain_full \
    python scripts/profile_chain.py 2368 1000 32 skytem > gpurun_out/${TAG}_chain_full.log 2>&1
ls -la gpurun_out | tail -12
