"""Team mode (GBP_TEAM = 2, 4) must give bit-identical results to GBP_TEAM = 1 (fp32 production kernel)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import _lib, ops
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
system = ops.resolve_system_struct()
opt = ops.make_options(n_markov_chains=10000)
for B, nit in ((300, 600), (2500, 300), (37, 1500)):
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.fdem_forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous()
    ref = None
    for team in ("1", "2", "4", "8"):
        os.environ["GBP_TEAM"] = team
        r = ops.rjmcmc_run(system, opt, data, t["height"], seed=5, max_iterations=nit, precision=32)
        torch.cuda.synchronize()
        r = {k: v.cpu().numpy() for k, v in r.items()}
        if ref is None:
            ref = r
        else:
            bad = [k for k in ref if not np.array_equal(ref[k], r[k], equal_nan=True)]
            print("B", B, "team", team, "identical" if not bad else "DIFFERENT: %s" % bad, "ms", ops.last_kernel_ms())
            if bad:
                d = np.flatnonzero((ref["scalars"] != r["scalars"]).any(axis=1))
                print("  chains differing:", d[:10], len(d))
print("done")
