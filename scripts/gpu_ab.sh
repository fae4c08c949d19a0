#!/bin/bash
# A/B of library builds on one GPU: scripts/gpu_ab.sh <what> <tag>=<lib.so> ...   (results: gpurun_out/ab_<tag>.json)
mkdir -p gpurun_out
WHAT=$1; shift
for kv in "$@"; do
  tag=${kv%%=*}; lib=${kv#*=}
  GBP_LIB_PATH=$PWD/$lib timeout 600 python scripts/gpu_perf_r02.py $tag $WHAT > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -5 gpurun_out/ab_$tag.err
  cat gpurun_out/ab_$tag.json
done
