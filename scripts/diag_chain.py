"""Diagnostics: chain outcome statistics + lone-warp speed."""
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
def data_for(B):
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.fdem_forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    return (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(), t["height"]
outs = ("ncells_hist", "scalars")
for prec, B in ((32, 1184), (64, 1184)):
    d, h = data_for(B)
    r = ops.rjmcmc_run(system, opt, d, h, seed=3, precision=prec, outputs=outs)
    torch.cuda.synchronize()
    s = r["scalars"].cpu().numpy()
    it = s[:, _lib.S_ITER]
    print("prec", prec, "B", B, "ms", ops.last_kernel_ms(), "failed", int(s[:, _lib.S_FAILED].sum()), "burned", int(s[:, _lib.S_BURNED_IN].sum()),
          "resets>0", int((s[:, _lib.S_N_RESETS] > 0).sum()), "iters mean", it.mean(), "min", it.min(), "max", it.max(),
          "acc", s[:, _lib.S_N_ACCEPT].sum() / it.sum(), "failed&notburned", int(((s[:, _lib.S_FAILED] > 0) & (s[:, _lib.S_BURNED_IN] == 0)).sum()),
          "iters<10000", int((it < 10000).sum()))
    nc = r["ncells_hist"].sum(dim=0).cpu().numpy()
    print("   ncells", nc[:10] / nc.sum())
    if prec == 32:
        bi = s[:, _lib.S_BURNED_IN] > 0
        for name, m in (("burned", bi), ("not burned", ~bi)):
            nn = r["ncells_hist"][torch.tensor(m, device=dev)].sum(dim=0).cpu().numpy()
            print("   ", name, "mean k", (nn * np.arange(nn.size)).sum() / nn.sum(), "fwd/it", s[m, _lib.S_N_FORWARD].sum() / it[m].sum())
# lone-warp speed
for B in (148, 148 * 4, 148 * 8):
    d, h = data_for(B)
    r = ops.rjmcmc_run(system, opt, d, h, seed=3, precision=32, outputs=outs, max_iterations=3000)
    torch.cuda.synchronize()
    its = float(r["scalars"][:, 0].sum())
    print("lone-warp test B", B, "ms", ops.last_kernel_ms(), "evals/s", its / ops.last_kernel_ms() * 1e3, "us/iter/chain", ops.last_kernel_ms() * 1e3 / 3000)
