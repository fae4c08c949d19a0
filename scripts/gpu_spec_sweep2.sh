#!/bin/bash
# sweep of the speculation threshold / helper count on the current build (bench batch to termination)
for cfg in "24 12" "8 12" "3 12" "1 12" "0 12" "3 6" "1 15" "3 15"; do
  set -- $cfg
  GBP_SPEC_MIN_REJECTIONS=$1 GBP_SPEC_HELPERS=$2 python scripts/gpu_timeline2.py "min_rej=$1 helpers=$2" 2>&1 | tail -1
done
