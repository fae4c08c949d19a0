"""Finish-time distribution of the bench-size batch under the current environment (GBP_TEAM*, GBP_SPEC_*)."""
import os, sys
os.environ["GBP_DEBUG_TIMELINE"] = "1"
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
for rep in range(2):
    r = ops.rjmcmc_run(system, opt, d, h, seed=20261017, precision=32, outputs=("scalars",))
    torch.cuda.synchronize()
ft = np.zeros(4096); _lib.check(_lib.load().gbp_debug_finish_times(ft.ctypes.data, 4096))
tot = r["scalars"][:, 24].cpu().numpy(); ms = ops.last_kernel_ms()
q = np.percentile(ft, [10, 25, 50, 75, 90, 95, 97, 99, 100])
print(sys.argv[1] if len(sys.argv) > 1 else "", "kernel ms %.0f  Meps %.2f" % (ms, tot.sum() / ms / 1e3), "finish pct [10,25,50,75,90,95,97,99,100] ms:", np.round(q).astype(int), "spec iters", int(ops.debug_counters()[8]) // 2)
