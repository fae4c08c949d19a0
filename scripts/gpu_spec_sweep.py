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
for helpers, minrej, idle_all in ((2, 24, 99), (3, 24, 99), (4, 24, 99), (6, 24, 99), (12, 24, 99)):
    os.environ["GBP_SPEC_HELPERS"] = str(helpers); os.environ["GBP_SPEC_MIN_REJECTIONS"] = str(minrej); pass
    import ctypes
    for rep in range(2):
        _lib.load().gbp_debug_counters(None, 1)
        r = ops.rjmcmc_run(system, opt, d, h, seed=20261017, precision=32, outputs=("scalars", "hitmap"))
        torch.cuda.synchronize()
    dc = (ctypes.c_ulonglong * 16)()
    _lib.load().gbp_debug_counters(dc, 0)
    dc = [float(x) for x in dc]
    if dc[0] > 0:
        us = lambda c: c / 1965.0
        print("   wake-ups %d: GO->awake %.1f us, state copy %.1f us | spec steps %d: %.1f us each | stopped helper waits %.1f us | per round: owner waits %.1f us, adopt/copy %.1f us" % (
            dc[0], us(dc[1] / dc[0]), us(dc[2] / dc[0]), dc[3], us(dc[4] / max(dc[3], 1)), us(dc[5] / dc[0]),
            us(dc[6] / max(dc[9], 1)), us(dc[7] / max(dc[9], 1))), flush=True)
    sc = r["scalars"]; its = float(sc[:, 24].sum()); ms = ops.last_kernel_ms()
    tot = sc[:, 24].cpu().numpy(); spec = np.zeros_like(tot)
    long = tot > 22000
    ck = sc[:, 5].cpu().numpy(); nf = sc[:, 9].cpu().numpy(); nsn = sc[:, 10].cpu().numpy(); acc = sc[:, 8].cpu().numpy()
    print("   long chains: mean current k %.2f (all %.2f), forwards/iter %.2f (all %.2f), jacobians/iter %.2f (all %.2f), acceptance %.3f (all %.3f), resets %.2f" % (
        ck[long].mean(), ck.mean(), nf[long].sum() / tot[long].sum(), nf.sum() / tot.sum(), nsn[long].sum() / tot[long].sum(), nsn.sum() / tot.sum(),
        acc[long].sum() / tot[long].sum(), acc.sum() / tot.sum(), sc[:, 19].cpu().numpy()[long].mean()))
    print("helpers", helpers, "min_rej", minrej, "idle_all", idle_all, "kernel ms", round(ms, 1), "evals/s %.4g" % (its / ms * 1e3), "speculated fraction all %.3f, of long chains (%d) %.3f" % (spec.sum() / tot.sum(), long.sum(), spec[long].sum() / max(tot[long].sum(), 1)), "max chain", tot.max(),
          "| rounds %d, iterations/round %.2f, us/round %.1f (at 1.965 GHz)" % (dc[9], spec.sum() / max(dc[9], 1), dc[10] / max(dc[9], 1) / 1965.0), flush=True)
