import os, sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import oracle_py as O
from geobipy_b200 import ops
from geobipy_b200.synthetic import synthetic_batch
sysc = ops.resolve_system_struct(); osys = O.make_system()
np.set_printoptions(linewidth=250, precision=2, suppress=True)
for sidx in (1, 5):
    b = synthetic_batch(sidx, 1); L = int(b["nlayers"][0])
    clean = O.fdem_forward(osys, b["height"][0], b["sigma"][0, :L], b["thickness"][0, :L])
    data = clean + b["noise"][0] * np.sqrt((0.05 * clean) ** 2 + 25.0)
    nrep = 384
    d = np.tile(data, (nrep, 1)); a = np.full(nrep, b["height"][0])
    opt = ops.make_options(n_markov_chains=10000)
    res = {}
    for prec, seed in ((32, 4242), (64, 4242), (640, 777)):
        r = ops.rjmcmc_run(sysc, opt, d, a, seed=seed, precision=64 if prec == 640 else prec, outputs=("hitmap", "ncells_hist", "scalars"))
        hm = r["hitmap"].sum(axis=0, dtype=np.int64); c = np.cumsum(hm, axis=0)
        med = np.array([np.searchsorted(c[:, j], 0.5 * c[-1, j]) for j in range(120)])
        mb = (hm[:, :120] * np.arange(250)[:, None]).sum(axis=0) / hm[:, :120].sum(axis=0)
        sc = r["scalars"]
        res[prec] = (med, mb, sc[:, 8].sum() / sc[:, 24].sum(), sc[:, 1].mean())
        print(sidx, prec, 'acc', res[prec][2], 'burned', res[prec][3])
    print('true edges', np.cumsum(b["thickness"][0, :L-1]), 'sig', b["sigma"][0, :L])
    print('med 32-64  ', res[32][0] - res[64][0])
    print('med 64b-64 ', res[640][0] - res[64][0])
    print('mean 32-64 ', res[32][1] - res[64][1])
    print('mean 64b-64', res[640][1] - res[64][1])
