"""Speculative-step cost vs concurrency on one SM (run under gpurun)."""
import os, sys, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops, _lib
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000)
for B in (1, 4):
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    d = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(); h = t["height"]
    for helpers, minrej in ((0, 0), (1, 0), (2, 0), (4, 0), (8, 0), (12, 0), (27, 0), (12, 24)):
        os.environ["GBP_SPEC_HELPERS"] = str(helpers); os.environ["GBP_SPEC_MIN_REJECTIONS"] = str(minrej)
        for rep in range(2):
            _lib.load().gbp_debug_counters(None, 1)
            r = ops.rjmcmc_run(system, opt, d, h, seed=3, max_iterations=4000, precision=32, outputs=("scalars",))
            torch.cuda.synchronize()
        dc = (ctypes.c_ulonglong * 16)(); _lib.load().gbp_debug_counters(dc, 0); dc = [float(x) for x in dc]
        sc = r["scalars"]; ms = ops.last_kernel_ms()
        print("B", B, "helpers", helpers, "min_rej", minrej, "kernel ms %.1f" % ms, "us/iter %.1f" % (ms * 1e3 / 4000), "speculated %.2f" % (dc[8] / sc[:, 24].sum().item()),
              "rounds", int(dc[9]), "| spec steps %d at %.1f us, wake %.1f us" % (dc[3], dc[4] / max(dc[3], 1) / 1965.0, dc[1] / max(dc[0], 1) / 1965.0), flush=True)
