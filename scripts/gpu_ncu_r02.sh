#!/bin/bash
# ncu --set full capture of one sampler launch (run under gpurun).  usage: gpu_ncu_r02.sh <tag> <B> <NIT> [workload] [lib]
TAG=$1; B=${2:-2368}; NIT=${3:-1000}; WL=${4:-resolve}; LIB=${5:-}
mkdir -p gpurun_out
[ -n "$LIB" ] && export GBP_LIB_PATH=$PWD/$LIB
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rjmcmc -s 1 -c 1 -f -o gpurun_out/${TAG} \
    python scripts/profile_chain.py $B $NIT 32 $WL > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
ls -la gpurun_out/${TAG}.ncu-rep
