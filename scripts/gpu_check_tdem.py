"""Exploratory GPU check of the time-domain path (run under gpurun): parity numbers + quick throughput."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O
from geobipy_b200 import ops

sv = ops.skytem_survey_struct(); S = O.make_tdem_system()
rng = np.random.default_rng(3)
B = 256
nl = rng.integers(1, 31, B).astype(np.int32); nl[:8] = [1, 1, 2, 2, 3, 3, 30, 30]
sig = 10.0 ** rng.uniform(-3.5, 0.5, (B, 30)); thk = rng.uniform(1.0, 40.0, (B, 30)); alt = rng.uniform(25.0, 45.0, B)
ref = np.zeros((B, 45)); refJ = np.zeros((B, 45, 30))
t = time.time()
for b in range(B):
    L = nl[b]
    ref[b] = O.tdem_forward(S, alt[b], sig[b, :L], thk[b, :L]); refJ[b, :, :L] = O.tdem_sensitivity(S, alt[b], sig[b, :L], thk[b, :L])
print("oracle fwd+J per sounding ms", (time.time() - t) / B * 1e3, flush=True)
tc = np.array(S.t_centre[:45]); floor = np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / tc)
for prec in (64, 32):
    pred, J = ops.forward(sv, nl, sig, thk, alt, precision=prec, sensitivity=True)
    pf = ops.forward(sv, nl, sig, thk, alt, precision=prec)
    e = np.abs(pred - ref) / np.abs(ref); ef = np.abs(pf - ref) / np.abs(ref)
    z = np.abs(pf - ref) / np.sqrt((0.05 * ref) ** 2 + floor ** 2)
    eJ = np.abs(J - refJ).max(axis=2) / np.abs(refJ).max(axis=2)
    print("prec", prec, "fwd(sens) max rel", e.max(), "fwd max rel", ef.max(), "median", np.median(ef), "max z", z.max(), "J max (per channel row)", eJ.max(), "median", np.median(eJ), flush=True)
    if prec == 32:
        print("  per-channel max rel", np.array2string(ef.max(0), precision=1), flush=True)
        print("  per-channel max z  ", np.array2string(z.max(0), precision=3), flush=True)
        w = np.unravel_index(ef.argmax(), ef.shape); print("  worst", w, nl[w[0]], ref[w], pf[w])

# chains: fp64 twin of the oracle
Bc = 8
oo = O.skytem_options(n_markov_chains=2000); opt = ops.make_options(n_markov_chains=2000, **ops.SKYTEM_OPTIONS)
data = np.zeros((Bc, 45)); altc = alt[:Bc]
for b in range(Bc):
    L = min(int(nl[b]), 4)
    clean = O.tdem_forward(S, altc[b], sig[b, :L], thk[b, :L])
    data[b] = clean + rng.standard_normal(45) * np.sqrt((0.05 * clean) ** 2 + floor ** 2)
NIT = 300
t = time.time()
res = ops.rjmcmc_run(sv, opt, data, altc, seed=11, max_iterations=NIT, precision=64)
print("gpu f64 chains", time.time() - t, "kernel ms", ops.last_kernel_ms(), flush=True)
same = 0
for b in range(Bc):
    r = O.run_chain(S, oo, data[b], altc[b], 11, b, max_iterations=NIT)
    s, q = res["scalars"][b], r["scalars"]
    hm = np.array_equal(res["hitmap"][b], r["hitmap"]); tr = np.array_equal(res["accept_trace"][b], r["accept_trace"])
    eh = np.array_equal(res["rel_hist"][b], r["rel_hist"]) and np.array_equal(res["add_hist"][b], r["add_hist"])
    same += hm and tr and eh
    print(b, "hitmap", hm, "trace", tr, "errh", eh, "halfspace", s[O.S_HALFSPACE], q[O.S_HALFSPACE], "acc", s[O.S_N_ACCEPT], q[O.S_N_ACCEPT],
          "misfit", s[O.S_CUR_MISFIT], q[O.S_CUR_MISFIT], "lik", s[O.S_CUR_LIKELIHOOD], q[O.S_CUR_LIKELIHOOD], "add2", s[O.S_CUR_ADD2], q[O.S_CUR_ADD2], flush=True)
print("identical chains", same, "/", Bc)
res32 = ops.rjmcmc_run(sv, opt, data, altc, seed=11, max_iterations=NIT, precision=32)
print("f32 kernel ms", ops.last_kernel_ms())
for b in range(Bc):
    s, q = res32["scalars"][b], res["scalars"][b]
    print(b, "f32 acc", s[O.S_N_ACCEPT], "f64", q[O.S_N_ACCEPT], "misfit", s[O.S_CUR_MISFIT], q[O.S_CUR_MISFIT], "lik", s[O.S_CUR_LIKELIHOOD], q[O.S_CUR_LIKELIHOOD],
          "add", s[O.S_CUR_ADD], q[O.S_CUR_ADD], "k", s[O.S_CUR_K], q[O.S_CUR_K])
# throughput
for Bt, nit in ((2368, 1000),):
    d2 = np.tile(data, (Bt // Bc + 1, 1))[:Bt]; a2 = np.tile(altc, Bt // Bc + 1)[:Bt]
    r2 = ops.rjmcmc_run(sv, opt, d2, a2, seed=5, max_iterations=nit, precision=32, outputs=("scalars", "hitmap"))
    ms = ops.last_kernel_ms()
    print("throughput fp32: B", Bt, "iters", nit, "kernel ms", ms, "evals/s", Bt * nit / ms * 1e3, "fwd/iter", r2["scalars"][:, O.S_N_FORWARD].mean() / nit)
