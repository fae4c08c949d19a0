"""Do consecutive batches overlap when they are launched on alternating streams?  K batches of the bench workload,
serial (one stream) vs two streams with separate result buffers; wall time by CUDA events across all streams."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import ops, _lib
from geobipy_b200.synthetic import synthetic_batch
dev = torch.device("cuda")
B, K = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4
system = ops.resolve_system_struct(); opt = ops.make_options(n_markov_chains=10000)
sb = synthetic_batch(0, B)
t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
clean = ops.forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
d = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous(); h = t["height"]
outs = ops.DEFAULT_OUTPUTS
shapes = ops.chain_buffer_shapes(opt, B)
tdt = {np.int32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}
for ns in (1, 2, 3):
    bufs = [{n: torch.zeros(shapes[n][0], dtype=tdt[shapes[n][1]], device=dev) for n in outs} for _ in range(ns)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    ops.rjmcmc_run(system, opt, d, h, seed=1, precision=32, outputs=outs, buffers=bufs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    its = []
    for i in range(K):
        with torch.cuda.stream(streams[i % ns]):
            r = ops.rjmcmc_run(system, opt, d, h, seed=20261017 + i, precision=32, outputs=outs, buffers=bufs[i % ns])
            its.append(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tot = float(sum(x.item() for x in its))
    print("streams %d: %d batches in %.0f ms = %.0f ms per batch, %.2f M evals/s" % (ns, K, dt * 1e3, dt * 1e3 / K, tot / dt / 1e6), flush=True)
    del bufs
    torch.cuda.empty_cache()
