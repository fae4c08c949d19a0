"""Progress of the longest chains of the bench batch: time per 2048 iterations (GBP_DEBUG_TIMELINE progress stamps)."""
import os, sys
os.environ["GBP_DEBUG_TIMELINE"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops, _lib
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
B = 4096
system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000)
sb = synthetic_batch(0, B)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
d = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(); h = t["height"]
for rep in range(2):
    r = ops.rjmcmc_run(system, opt, d, h, seed=20261017, precision=32, outputs=("scalars",))
    torch.cuda.synchronize()
ft = np.zeros(B); _lib.check(_lib.load().gbp_debug_finish_times(ft.ctypes.data, B))
pt = np.zeros((B, 32)); _lib.check(_lib.load().gbp_debug_progress_times(pt.ctypes.data, B))
s = r["scalars"].cpu().numpy()
tot = s[:, _lib.S_TOTAL_ITER]; acc = s[:, _lib.S_N_ACCEPT]
print("kernel ms %.0f" % ops.last_kernel_ms())
np.set_printoptions(linewidth=250)
for i in np.argsort(ft)[-14:]:
    p = pt[i]; ok = p >= 0
    dt = np.diff(p[ok]) / 2.048   # us per iteration in each 2048-iteration block
    print("chain %4d start %4.0f end %4.0f iters %5d acc %.2f resets %d | us/iter per block:" % (i, p[0], ft[i], tot[i], acc[i] / tot[i], s[i, _lib.S_N_RESETS]), np.round(dt).astype(int))
# the bulk: median us/iter of first-wave chains in their first 4 blocks
first = pt[:, 0] < 5.0
blk = (pt[first, 4] - pt[first, 0]) / (4 * 2.048)
print("first-wave chains: us/iter over their first 8192 iterations: median %.1f p10 %.1f p90 %.1f" % tuple(np.percentile(blk[blk > 0], [50, 10, 90])))
late = np.argsort(pt[:, 0])[-200:]
print("last 200 chains to START: start times p10/p50/p90/max", np.round(np.percentile(pt[late, 0], [10, 50, 90, 100])).astype(int), "their iterations p50/max", np.percentile(tot[late], [50, 100]))
