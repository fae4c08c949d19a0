"""Makespan diagnostics for the bench workload: distribution of per-chain total iterations."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops, _lib
from geobipy_b200.synthetic import synthetic_batch
system = ops.resolve_system_struct()
opt = ops.make_options(n_markov_chains=10000)
dev = torch.device("cuda")
B = 4096
sb = synthetic_batch(0, B)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.fdem_forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
d = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous()
for rep in range(2):
    r = ops.rjmcmc_run(system, opt, d, t["height"], seed=20261017 + rep, precision=32, outputs=("ncells_hist", "scalars"))
    torch.cuda.synchronize()
s = r["scalars"].cpu().numpy()
tot = s[:, _lib.S_TOTAL_ITER]
ms = ops.last_kernel_ms()
print("kernel ms", ms, "total iters", tot.sum(), "evals/s", tot.sum() / ms * 1e3)
print("total-iter percentiles 50/90/99/99.9/max:", np.percentile(tot, [50, 90, 99, 99.9, 100]))
print("resets hist", np.bincount(s[:, _lib.S_N_RESETS].astype(int)), "burned", s[:, _lib.S_BURNED_IN].sum(), "failed", s[:, _lib.S_FAILED].sum())
print("makespan/max-chain us per iter", ms * 1e3 / tot.max(), " mean-chain/makespan utilisation", tot.mean() / tot.max())
srt = np.sort(tot)[::-1]
print("top 10 chain lengths", srt[:10])
