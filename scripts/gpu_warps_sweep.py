"""Resident chains per SM (frequency-domain fp32 kernel): full waves and the bench-size batch (run under gpurun).
Historical: needs the 20 / 24 / 28-warp instantiations of round 1 behind GBP_FDEM_WARPS (commit 43c46c2 .. 795c10d); the
current library builds the 16-warp kernel only and ignores the variable."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000)
sb = synthetic_batch(0, 16384)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(); h = t["height"]
for warps in (28, 20, 16):
    os.environ["GBP_FDEM_WARPS"] = str(warps)
    for B, nit in ((8192, 0), (16384, 0), (2048, 0)):
        d = data[:B]
        hh = h[:B]
        for rep in range(2):
            r = ops.rjmcmc_run(system, opt, d.contiguous(), hh.contiguous(), seed=20261017, max_iterations=nit, precision=32, outputs=("scalars", "hitmap"))
            torch.cuda.synchronize()
            its = float(r["scalars"][:, 24].sum()); ms = ops.last_kernel_ms()
            print("warps", warps, "B", B, "max_it", nit, "kernel ms %.1f" % ms, "evals/s %.4g" % (its / ms * 1e3), flush=True)
