"""Short rjMCMC run for ncu (one wave of chains, bounded iterations)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
NIT = int(sys.argv[2]) if len(sys.argv) > 2 else 300
PREC = int(sys.argv[3]) if len(sys.argv) > 3 else 32
WORKLOAD = sys.argv[4] if len(sys.argv) > 4 else "resolve"
dev = torch.device("cuda")
if WORKLOAD == "skytem":
    system = ops.skytem_survey_struct()
    opt = ops.make_options(n_markov_chains=10000, **ops.SKYTEM_OPTIONS)
    sb = synthetic_batch(0, B, max_depth=400.0, n_channels=45)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    tc = ops.tdem_window_operator(system)[3]
    add = torch.tensor(np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / tc), device=dev)
    data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + add ** 2)).contiguous()
else:
    system = ops.resolve_system_struct()
    opt = ops.make_options(n_markov_chains=10000)
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.fdem_forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous()
outs = ("hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "misfit_trace", "accept_trace", "scalars")
for rep in range(2):
    r = ops.rjmcmc_run(system, opt, data, t["height"], seed=rep, max_iterations=NIT, precision=PREC, outputs=outs)
    torch.cuda.synchronize()
    its = float(r["scalars"][:, 0].sum())
    print("rep", rep, "B", B, "iters", its, "kernel ms", ops.last_kernel_ms(), "evals/s", its / (ops.last_kernel_ms() * 1e-3))
