"""Time-domain kernel throughput check (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch
import oracle_py as O
dev = torch.device("cuda")
system = ops.skytem_survey_struct()
opt = ops.make_options(n_markov_chains=10000, **ops.SKYTEM_OPTIONS)
tc = ops.tdem_window_operator(system)[3]
add = torch.tensor(np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / tc), device=dev)
def data_for(B):
    sb = synthetic_batch(0, B, max_depth=400.0, n_channels=45)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    return (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + add ** 2)).contiguous(), t["height"], sb
# accuracy of the fp32 forward / Jacobian against the fp64 path
d, h, sb = data_for(256)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
p64, J64 = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64, sensitivity=True)
p32, J32 = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=32, sensitivity=True)
sn = torch.sqrt((0.05 * p64) ** 2 + add ** 2)
print("fp32 vs fp64: median rel", float(torch.median(torch.abs(p32 / p64 - 1))), "max |d|/(2e-4|ref|+0.01 sn)", float((torch.abs(p32 - p64) / (2e-4 * torch.abs(p64) + 0.01 * sn)).max()),
      "J row-rel max", float((torch.abs(J32 - J64).amax(2) / torch.abs(J64).amax(2)).max()), flush=True)
for warps in (16,):
    os.environ["GBP_TDEM_WARPS"] = str(warps)
    for B, nit in ((148 * warps, 1000), (4096, 0)):
        d, h, _ = data_for(B)
        for rep in range(2):
            r = ops.rjmcmc_run(system, opt, d, h, seed=rep, max_iterations=nit, precision=32, outputs=("scalars", "hitmap"))
            torch.cuda.synchronize()
        its = float(r["scalars"][:, 24].sum()); ms = ops.last_kernel_ms()
        print("warps", warps, "B", B, "max_it", nit, "iters", its, "kernel ms", round(ms, 2), "evals/s %.4g" % (its / ms * 1e3), flush=True)
