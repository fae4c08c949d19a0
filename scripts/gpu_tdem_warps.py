"""Resident chains per SM for the time-domain kernel (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
system = ops.skytem_survey_struct(); opt = ops.make_options(n_markov_chains=10000, **ops.SKYTEM_OPTIONS)
tc = ops.tdem_window_operator(system)[3]
add = torch.tensor(np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / tc), device=dev)
sb = synthetic_batch(0, 8192, max_depth=400.0, n_channels=45)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + add ** 2)).contiguous(); h = t["height"]
for warps in (16, 12, 8, 16, 12):
    os.environ["GBP_TDEM_WARPS"] = str(warps)
    for B in (4096, 8192):
        for rep in range(2):
            r = ops.rjmcmc_run(system, opt, data[:B].contiguous(), h[:B].contiguous(), seed=20261017, precision=32, outputs=("scalars", "hitmap"))
            torch.cuda.synchronize()
        its = float(r["scalars"][:, 24].sum()); ms = ops.last_kernel_ms()
        print("tdem warps", warps, "B", B, "kernel ms %.1f" % ms, "evals/s %.4g" % (its / ms * 1e3), flush=True)
