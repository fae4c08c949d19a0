"""Cost of sampling the sensor height (solve_z) on the RESOLVE workload (run under gpurun).

Same 4096 synthetic soundings as BASELINE configs[1], inverted (a) with the height fixed and (b) with solve_z
(prior +-1 m, proposal std 0.1 m; speculation is off on that path), to the reference's termination rule, and as
full waves of equal-length chains (2368 soundings x 2000 iterations)."""
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import _lib, ops
from geobipy_b200.synthetic import synthetic_batch

dev = torch.device("cuda")
system = ops.resolve_system_struct()
out = {}
for B, nit in ((148 * 16, 2000), (4096, 0)):
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous()
    for name, kw in (("fixed", {}), ("solve_z", dict(solve_height=1, max_height_change=1.0, height_prop_var=0.01))):
        opt = ops.make_options(n_markov_chains=10000, **kw)
        for rep in range(2):
            r = ops.rjmcmc_run(system, opt, data, t["height"], seed=rep, max_iterations=nit, precision=32,
                               outputs=("scalars", "hitmap") + (("height_hist",) if kw else ()))
            torch.cuda.synchronize()
        s = r["scalars"]
        its = float(s[:, _lib.S_TOTAL_ITER].sum())
        ms = ops.last_kernel_ms()
        rec = dict(B=B, max_iterations=nit, iterations=its, kernel_ms=round(ms, 2), evals_per_s=its / ms * 1e3,
                   burned_in=float((s[:, _lib.S_BURNED_IN] == 1).double().mean()),
                   acceptance=float(s[:, _lib.S_N_ACCEPT].sum() / s[:, _lib.S_ITER].sum()))
        if kw:
            rec["mean_abs_height_change_m"] = float((s[:, _lib.S_CUR_HEIGHT] - t["height"]).abs().mean())
        out["%s_B%d_it%d" % (name, B, nit)] = rec
        print(name, rec, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "height_perf.json"), "w"), indent=1)
