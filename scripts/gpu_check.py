"""Exploratory GPU check (run under gpurun): parity numbers + quick throughput."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch

sysc = ops.resolve_system_struct(); osys = O.make_system()
g = np.load(os.path.join(ROOT, "tests/golden/fdem_random_models.npz"))
nl = g["nlayers"]; sig = np.nan_to_num(g["sigma"], nan=1.0); thk = np.nan_to_num(g["thickness"], nan=1.0, posinf=1e30)
for prec in (64, 32):
    t = time.time()
    pred, J = ops.fdem_forward(sysc, nl, sig, thk, g["height"], precision=prec, sensitivity=True)
    pf = ops.fdem_forward(sysc, nl, sig, thk, g["height"], precision=prec)
    ref = g["forward"]; refJ = np.nan_to_num(g["sensitivity"], nan=0.0)
    e = np.abs(pred - ref) / (np.abs(ref) + 1.0)
    ef = np.abs(pf - ref) / (np.abs(ref) + 1.0)
    eJ = np.abs(J - refJ).max(axis=(1, 2)) / np.abs(refJ).max(axis=(1, 2))
    print("prec", prec, "fwd(sens pass) max", e.max(), "fwd max", ef.max(), "J max", eJ.max(), "worst L", nl[eJ.argmax()], "t", time.time() - t, flush=True)

B = 16
batch = synthetic_batch(0, B)
data = np.zeros((B, 12))
for b in range(B):
    L = int(batch["nlayers"][b])
    clean = O.fdem_forward(osys, batch["height"][b], batch["sigma"][b, :L], batch["thickness"][b, :L])
    data[b] = clean + batch["noise"][b] * np.sqrt((0.05 * clean) ** 2 + 25.0)
NIT = 600
opt = ops.make_options(n_markov_chains=2000); oo = O.resolve_options(n_markov_chains=2000)
t = time.time()
res = ops.rjmcmc_run(sysc, opt, data, batch["height"], seed=11, max_iterations=NIT, precision=64)
print("gpu f64 chain time", time.time() - t, "kernel ms", ops.last_kernel_ms(), flush=True)
same = 0
for b in range(B):
    r = O.run_chain(osys, oo, data[b], batch["height"][b], 11, b, max_iterations=NIT)
    s, q = res["scalars"][b], r["scalars"]
    hm_eq = np.array_equal(res["hitmap"][b], r["hitmap"])
    tr_eq = np.array_equal(res["accept_trace"][b], r["accept_trace"])
    same += hm_eq and tr_eq
    first_div = int(np.argmax(res["accept_trace"][b] != r["accept_trace"])) if not tr_eq else -1
    print(b, "it", s[0], q[0], "acc", s[8], q[8], "k", s[5], q[5], "misfit", s[14], q[14], "hs", s[6] == q[6], "hitmap==", hm_eq, "trace==", tr_eq, "div@", first_div,
          "ncells==", np.array_equal(res["ncells_hist"][b], r["ncells_hist"]), "edges==", np.array_equal(res["edges_hist"][b], r["edges_hist"]), flush=True)
print("identical trajectories:", same, "of", B)

# throughput probe fp32 and fp64
import torch
for prec, Bb, nit in ((32, 2368, 1000), (64, 1184, 300)):
    batch = synthetic_batch(0, 64)
    d = np.tile(data, (Bb // B + 1, 1))[:Bb]; h = np.tile(batch["height"][:B], Bb // B + 1)[:Bb]
    dd = torch.tensor(d, device="cuda"); hh = torch.tensor(h, device="cuda")
    outs = ("hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "scalars")
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        r = ops.rjmcmc_run(sysc, opt, dd, hh, seed=5, max_iterations=nit, precision=prec, outputs=outs)
        torch.cuda.synchronize(); dt = time.time() - t
        its = float(r["scalars"][:, 0].sum())
        print("prec", prec, "B", Bb, "iters", its, "wall", dt, "kernel ms", ops.last_kernel_ms(), "evals/s", its / dt, "acc", float(r["scalars"][:, 8].sum()) / its, flush=True)
