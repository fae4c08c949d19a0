"""Throughput of the frequency-domain sampler against resident chains per SM (equal-length chains, CUDA-event kernel time)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import _lib, ops
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
system = ops.resolve_system_struct()
opt = ops.make_options(n_markov_chains=10000)
NIT = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
res = {}
for per_sm in (1, 2, 4, 8, 12, 16):
    B = 148 * per_sm
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.fdem_forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous()
    for rep in range(2):
        r = ops.rjmcmc_run(system, opt, data, t["height"], seed=rep, max_iterations=NIT, precision=32, outputs=("hitmap", "scalars"))
        torch.cuda.synchronize()
    ms = ops.last_kernel_ms()
    res[per_sm] = dict(ms=ms, Meps=B * NIT / ms / 1e3, us_per_iter=ms * 1e3 / NIT)
print(json.dumps(res))
