"""Round-2 kernel A/B harness (one B200): accuracy of the fp32 forward / Jacobian against the fp64 kernels, then the
three timing regimes of the frequency-domain sampler: a full wave of equal-length chains (2368 x 2000 iterations), one
chain per SM (148 x 2000: the tail regime) and the bench batch (4096 soundings to termination).  CUDA-event kernel times.
usage: [GBP_LIB_PATH=...] python scripts/gpu_perf_r02.py [tag] [what=acc,wave,lone,bench]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from geobipy_b200 import _lib, ops
from geobipy_b200.synthetic import synthetic_batch

tag = sys.argv[1] if len(sys.argv) > 1 else "cur"
what = (sys.argv[2] if len(sys.argv) > 2 else "acc,wave,lone,bench").split(",")
dev = torch.device("cuda")
system = ops.resolve_system_struct()
out = {"tag": tag, "lib": _lib.LIB_PATH}

if "acc" in what:
    rng = np.random.default_rng(7)
    n = 4096
    nl = rng.integers(1, 31, n).astype(np.int32)
    nl[:1024] = rng.integers(1, 5, 1024)
    sig = 10.0 ** rng.uniform(-4.0, 1.0, (n, 30))
    thk = np.exp(rng.uniform(np.log(1.0), np.log(150.0), (n, 30)))
    alt = rng.uniform(20.0, 60.0, n)
    p64, J64 = ops.fdem_forward(system, nl, sig, thk, alt, precision=64, sensitivity=True)
    p32, J32 = ops.fdem_forward(system, nl, sig, thk, alt, precision=32, sensitivity=True)
    f32 = ops.fdem_forward(system, nl, sig, thk, alt, precision=32)
    out["fwd_err"] = float(np.max(np.abs(f32 - p64) / (np.abs(p64) + 10.0)))       # relative to |d| + 10 ppm
    out["fwdJ_err"] = float(np.max(np.abs(p32 - p64) / (np.abs(p64) + 10.0)))
    out["fwd_abs_ppm"] = float(np.max(np.abs(f32 - p64)))
    out["J_err"] = float(np.max(np.abs(J32 - J64).max(axis=(1, 2)) / np.abs(J64).max(axis=(1, 2))))
    # forward-only throughput
    for L in (3, 10, 30):
        B = 65536
        t = {"nl": torch.full((B,), L, dtype=torch.int32, device=dev), "sig": torch.tensor(np.tile(sig[:64, :], (B // 64, 1)), device=dev),
             "thk": torch.tensor(np.tile(thk[:64, :], (B // 64, 1)), device=dev), "alt": torch.tensor(np.tile(alt[:64], B // 64), device=dev)}
        for sens in (False, True):
            for _ in range(2):
                ops.fdem_forward(system, t["nl"], t["sig"], t["thk"], t["alt"], precision=32, sensitivity=sens)
                torch.cuda.synchronize()
            out["fwd%s_L%d_per_s" % ("J" if sens else "", L)] = B / (ops.last_kernel_ms() * 1e-3)


def chains(B, nit, reps=2, seed0=0):
    opt = ops.make_options(n_markov_chains=10000)
    sb = synthetic_batch(0, B)
    t = {k: torch.tensor(v, device=dev) for k, v in sb.items()}
    clean = ops.fdem_forward(system, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64)
    data = (clean + t["noise"] * torch.sqrt((0.05 * clean) ** 2 + 25.0)).contiguous()
    outs = ("hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "misfit_trace", "accept_trace", "scalars")
    best = None
    for rep in range(reps):
        r = ops.rjmcmc_run(system, opt, data, t["height"], seed=seed0 + rep, max_iterations=nit, precision=32, outputs=outs)
        torch.cuda.synchronize()
        its = float(r["scalars"][:, _lib.S_TOTAL_ITER].sum())
        ms = ops.last_kernel_ms()
        if rep > 0 or reps == 1:
            best = (its, ms, float(r["scalars"][:, _lib.S_N_ACCEPT].sum() / its))
    del r
    torch.cuda.empty_cache()
    return best

if "wave" in what:
    its, ms, acc = chains(2368, 2000)
    out["wave_Meps"], out["wave_ms"], out["wave_acc"] = its / ms / 1e3, ms, acc
if "lone" in what:
    its, ms, acc = chains(148, 2000)
    out["lone_us_per_iter"], out["lone_ms"] = ms * 1e3 / 2000, ms
if "bench" in what:
    its, ms, acc = chains(4096, 0, reps=2, seed0=20261017)
    out["bench_Meps"], out["bench_ms"], out["bench_acc"], out["bench_iters"] = its / ms / 1e3, ms, acc, its
print(json.dumps(out))
