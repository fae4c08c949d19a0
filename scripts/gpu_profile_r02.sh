#!/bin/bash
# Round-2 captures of the final build (run under gpurun): bench launch list, bench-size counters + DRAM traffic of both
# sampler kernels, one ncu --set full capture (with source) of a full wave.  Writes into gpurun_out/.
TAG=${1:-r02_final}
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:rjmcmc -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_chain_benchsize_metrics.csv \
    python scripts/profile_chain.py 4096 0 > gpurun_out/${TAG}_chain_benchsize.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:rjmcmc -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_tdem_chain_benchsize_metrics.csv \
    python scripts/profile_chain.py 4096 0 32 skytem > gpurun_out/${TAG}_tdem_chain_benchsize.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rjmcmc -s 1 -c 1 -f -o gpurun_out/${TAG}_wave \
    python scripts/profile_chain.py 2368 1000 > gpurun_out/${TAG}_wave.log 2>&1
ncu -i gpurun_out/${TAG}_wave.ncu-rep --page raw --csv > gpurun_out/${TAG}_wave_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_wave.ncu-rep --page source --csv > gpurun_out/${TAG}_wave_source.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG}
