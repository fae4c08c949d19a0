"""Sweep of scheduling knobs (environment variables read at every launch) on the bench batch, one process.
usage: python scripts/gpu_tail_sweep.py "K1=v K2=v" "K1=v ..." ...   (each argument = one configuration)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops, _lib
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
B = int(os.environ.get("SWEEP_B", "4096"))
system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000)
sb = synthetic_batch(0, B)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
d = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(); h = t["height"]
ref = None
for cfg in sys.argv[1:]:
    keys = []
    for kv in cfg.split():
        k, v = kv.split("="); os.environ[k] = v; keys.append(k)
    ms = []
    for rep in range(4):
        r = ops.rjmcmc_run(system, opt, d, h, seed=20261017 + (rep % 2), precision=32, outputs=("scalars", "ncells_hist"))
        torch.cuda.synchronize()
        if rep >= 2:
            ms.append(ops.last_kernel_ms())
    its = float(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
    chk = int(r["ncells_hist"].sum()) ^ int((r["scalars"][:, _lib.S_N_ACCEPT].sum()).item())
    if ref is None:
        ref = chk
    print("%-60s ms (2 seeds) %s  Meps %.2f  identical %s" % (cfg, np.round(ms).astype(int), its / ms[-1] / 1e3, chk == ref), flush=True)
    for k in keys:
        os.environ.pop(k, None)
