#!/bin/bash
# 8-GPU lines for profiles/: the bench (configs[1] per GPU), the mixed flight line (configs[4] family) and a reduced
# configs[2] (8192 soundings per GPU x n_markov_chains 100000).  usage: gpurun --gpus 8 -- bash scripts/gpu_scale8.sh
N=${1:-8}; TAG=${2:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N"
$RUN --steps 3 --warmup 3 --no-forward-only 2> gpurun_out/${TAG}_bench_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu.json
$RUN --workload mixed --steps 1 --warmup 1 2> gpurun_out/${TAG}_mixed_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_mixed_${N}gpu.json
$RUN --soundings 8192 --chains 100000 --streams 1 --steps 1 --warmup 1 --no-e2e --no-forward-only 2> gpurun_out/${TAG}_config3_reduced_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_config3_reduced_${N}gpu.json
for f in bench mixed config3_reduced; do python -c "
import json; d=json.load(open('gpurun_out/${TAG}_${f}_${N}gpu.json')); print('$f', 'Meps %.1f' % (d['value']/1e6), 'ms/step %.0f' % d['ms_per_step'], d.get('collation'), d.get('e2e') and round(d['e2e']['value']/1e6,1))"; done
