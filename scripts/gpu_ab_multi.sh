#!/bin/bash
# full wave / lone chain of several builds of the library, interleaved: usage gpu_ab_multi.sh <tag> <lib.so>...
TAG=$1; shift
mkdir -p gpurun_out
{
for rep in 1 2 3; do
  for L in "$@"; do GBP_LIB_PATH=$L python scripts/gpu_perf_r02.py $L wave,lone; done
done
} > gpurun_out/${TAG}.log 2>&1
grep -o '"tag[^}]*lone_us_per_iter": [0-9.]*' gpurun_out/${TAG}.log
