"""Sweep of the speculation parameters on the bench-size batch (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops, _lib
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000)
sb = synthetic_batch(0, 4096)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
d = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(); h = t["height"]
for helpers, minrej in ((0, 24), (12, 24), (12, 8), (12, 64), (6, 24), (27, 24), (27, 8)):
    os.environ["GBP_SPEC_HELPERS"] = str(helpers); os.environ["GBP_SPEC_MIN_REJECTIONS"] = str(minrej)
    for rep in range(2):
        r = ops.rjmcmc_run(system, opt, d, h, seed=20261017, precision=32, outputs=("scalars", "hitmap"))
        torch.cuda.synchronize()
    sc = r["scalars"]; its = float(sc[:, 24].sum()); ms = ops.last_kernel_ms()
    tot = sc[:, 24].cpu().numpy(); spec = sc[:, _lib.S_N_SPECULATED].cpu().numpy()
    long = tot > 22000
    print("helpers", helpers, "min_rej", minrej, "kernel ms", round(ms, 1), "evals/s %.4g" % (its / ms * 1e3), "speculated fraction all %.3f, of long chains (%d) %.3f" % (spec.sum() / tot.sum(), long.sum(), spec[long].sum() / max(tot[long].sum(), 1)), "max chain", tot.max(), flush=True)
