"""Speculative-evaluation check (run under gpurun): bit-identity with the sequential sampler and speed, small
batches (every chain gets helpers) and the bench-size batch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch
which = sys.argv[1] if len(sys.argv) > 1 else "resolve"
dev = torch.device("cuda")
if which == "skytem":
    system = ops.skytem_survey_struct(); opt = ops.make_options(n_markov_chains=10000, **ops.SKYTEM_OPTIONS)
    tc = ops.tdem_window_operator(system)[3]
    add = torch.tensor(np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / tc), device=dev)
    kw = dict(max_depth=400.0, n_channels=45)
else:
    system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000); add = 5.0; kw = {}
def data_for(B):
    sb = synthetic_batch(0, B, **kw)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    return (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + add ** 2)).contiguous(), t["height"]
cases = ((1, 3000), (148, 3000), (592, 3000), (4144, 2000), (4096, 0)) if len(sys.argv) < 3 else ((int(sys.argv[2]), int(sys.argv[3])),)
for B, nit in cases:
    d, h = data_for(B)
    ref = None
    for helpers in (0, 12):
        os.environ["GBP_SPEC_HELPERS"] = str(helpers)
        for rep in range(2):
            r = ops.rjmcmc_run(system, opt, d, h, seed=7, max_iterations=nit, precision=32, outputs=("scalars", "hitmap", "ncells_hist", "accept_trace"))
            torch.cuda.synchronize()
        its = float(r["scalars"][:, 24].sum()); ms = ops.last_kernel_ms()
        same = ""
        if ref is None:
            ref = r
        else:
            same = "identical: hitmap %s scalars %s trace %s" % (torch.equal(ref["hitmap"], r["hitmap"]), torch.equal(ref["scalars"], r["scalars"]), torch.equal(ref["accept_trace"], r["accept_trace"]))
        print(which, "helpers", helpers, "B", B, "max_it", nit, "iters", its, "kernel ms", round(ms, 2), "evals/s %.4g" % (its / ms * 1e3), same, flush=True)
