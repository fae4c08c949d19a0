#!/bin/bash
# A/B of two library builds through bench.py on the same box, interleaved: scripts/gpu_ab_bench.sh <libA.so> <libB.so> [bench args]
A=$1; B=$2; shift 2
for rep in 1 2; do
  for lib in "$A" "$B"; do
    GBP_LIB_PATH=$PWD/$lib python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-forward-only --no-e2e "$@" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'Meps %.2f' % (d['value']/1e6), 'ms/step %.0f' % d['ms_per_step'], 'alone %.0f' % d['roofline']['kernel_ms_alone'])"
  done
done
