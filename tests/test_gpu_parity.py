"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the oracle and the
committed golden vectors.

Tolerances (north_star: "stated fp64 -> fp32 tolerance"):
  fp64 instantiation : forward |d - ref| <= 5e-8 (|ref| + 1 ppm); Jacobian 1e-8 of max |J|
                       (the reference's own (H-H0)/H0 round-off floor is ~1e-8 ppm)
  fp32 instantiation : forward |d - ref| <= 2e-4 |ref| + 2e-3 ppm (the absolute floor is the fp32 round-off
                       of a 120/140-term oscillating filter sum, 1e-3 of the smallest additive error the
                       options allow, 3 ppm); Jacobian 1e-4 of max |J|
  fp64 chains        : same random stream as the oracle -> identical accept/reject trajectories
                       (>= 90 % of chains bit-identical over 400 iterations; others may flip on round-off)
  fp32 chains        : statistical agreement with the oracle ensemble
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F64_FWD, F64_J = (5e-8, 5e-8), 1e-8      # (relative, absolute ppm)
F32_FWD, F32_J = (2e-4, 2e-3), 1e-4


def fwd_ok(pred, ref, tol):
    return bool(np.all(np.abs(pred - ref) <= tol[0] * np.abs(ref) + tol[1]))


@pytest.fixture(scope="module")
def gpu(built_lib):
    from geobipy_b200 import _lib, ops
    _lib.require_cuda()
    return ops


@pytest.fixture(scope="module")
def systems(gpu, oracle):
    return gpu.resolve_system_struct(), oracle.make_system()


def _observed(oracle, osys, n, first=0):
    from geobipy_b200.synthetic import synthetic_batch
    b = synthetic_batch(first, n)
    data = np.zeros((n, 12))
    for i in range(n):
        L = int(b["nlayers"][i])
        clean = oracle.fdem_forward(osys, b["height"][i], b["sigma"][i, :L], b["thickness"][i, :L])
        data[i] = clean + b["noise"][i] * np.sqrt((0.05 * clean) ** 2 + 25.0)
    return data, b["height"]


@pytest.mark.parametrize("prec,tol", [(64, F64_FWD), (32, F32_FWD)])
def test_forward_reference_csv_goldens(gpu, systems, golden_dir, prec, tol):
    """The reference's own known-answer vectors: 6 models x 79 soundings x 12 channels."""
    g = np.load(os.path.join(golden_dir, "resolve_clean.npz"))
    sig = np.repeat(g["sigma"][:, None, :], 79, axis=1).reshape(-1, 3)
    thk = np.tile(np.stack([g["zwedge"], g["zdeep"] - g["zwedge"], np.full(79, np.inf)], axis=1), (6, 1))
    out = gpu.fdem_forward(systems[0], np.full(474, 3, np.int32), sig, thk, np.full(474, float(g["height"])), precision=prec)
    ref = g["data"].reshape(-1, 12)
    if prec == 64:
        assert np.allclose(out, ref)   # the reference's own criterion (tests/test_synthetic_data.py:30)
    assert fwd_ok(out, ref, tol)


@pytest.mark.parametrize("prec,tf,tj", [(64, F64_FWD, F64_J), (32, F32_FWD, F32_J)])
def test_forward_and_jacobian_random_models(gpu, systems, golden_dir, prec, tf, tj):
    g = np.load(os.path.join(golden_dir, "fdem_random_models.npz"))
    nl = g["nlayers"]
    sig = np.nan_to_num(g["sigma"], nan=1.0)
    thk = np.nan_to_num(g["thickness"], nan=1.0, posinf=np.inf)
    pred, J = gpu.fdem_forward(systems[0], nl, sig, thk, g["height"], precision=prec, sensitivity=True)
    pred2 = gpu.fdem_forward(systems[0], nl, sig, thk, g["height"], precision=prec)
    assert fwd_ok(pred2, pred, tf)
    assert fwd_ok(pred, g["forward"], tf)
    refJ = np.nan_to_num(g["sensitivity"], nan=0.0)
    err = np.abs(J - refJ).max(axis=(1, 2)) / np.abs(refJ).max(axis=(1, 2))
    assert err.max() < tj
    # columns beyond nlayers are zero
    for i in range(len(nl)):
        assert not J[i, :, nl[i]:].any()


@pytest.mark.parametrize("prec,tf,tj", [(64, F64_FWD, F64_J), (32, F32_FWD, (2e-4))])
def test_mixed_tensor_components_and_coil_offsets(gpu, golden_dir, prec, tf, tj):
    """Tensor ids 3 / 7 (Hxz, Hzx: fdem1d_numba.py:359-408) and non-zero vertical coil offsets (fdem1d.py:31-32)
    against the live reference (make_golden.py fdem_tensor): fp64 <= 5e-8, fp32 <= 2e-4."""
    g = np.load(os.path.join(golden_dir, "fdem_tensor_models.npz"))
    sysc = gpu.make_system_struct(*(list(g["sys_" + k]) for k in ("freq", "tor", "tmom", "tx", "ty", "tz", "ror", "rmom", "rx", "ry", "rz")))
    assert list(sysc.tid[:6]) == [3, 7, 1, 9, 3, 7]
    nl = g["nlayers"]
    sig = np.nan_to_num(g["sigma"], nan=1.0)
    thk = np.nan_to_num(g["thickness"], nan=1.0, posinf=np.inf)
    pred, J = gpu.fdem_forward(sysc, nl, sig, thk, g["height"], precision=prec, sensitivity=True)
    assert fwd_ok(pred, g["forward"], tf), np.max(np.abs(pred - g["forward"]) / (np.abs(g["forward"]) + 1.0))
    assert fwd_ok(gpu.fdem_forward(sysc, nl, sig, thk, g["height"], precision=prec), pred, tf)
    refJ = np.nan_to_num(g["sensitivity"], nan=0.0)
    err = np.abs(J - refJ).max(axis=(1, 2)) / np.abs(refJ).max(axis=(1, 2))
    assert err.max() < tj, err.max()


def test_forward_against_oracle_extremes(gpu, systems, oracle):
    """Conductive / resistive extremes, 1 and 30 layers, thin and thick layers, low and high sensors."""
    rng = np.random.default_rng(99)
    n = 200
    nl = rng.choice([1, 2, 3, 10, 30], n).astype(np.int32)
    sig = 10 ** rng.uniform(-4.5, 1.0, (n, 30))
    sig[:40] = 10 ** rng.choice([-4.0, 1.0], (40, 30))
    thk = np.exp(rng.uniform(np.log(1.0), np.log(150.0), (n, 30)))
    alt = rng.uniform(10.0, 120.0, n)
    for prec, tf, tj in ((64, F64_FWD, F64_J), (32, F32_FWD, F32_J)):
        pred, J = gpu.fdem_forward(systems[0], nl, sig, thk, alt, precision=prec, sensitivity=True)
        for i in range(n):
            L = int(nl[i])
            t = np.r_[thk[i, :L - 1], np.inf]
            ref = oracle.fdem_forward(systems[1], alt[i], sig[i, :L], t)
            refJ = oracle.fdem_sensitivity(systems[1], alt[i], sig[i, :L], t)
            assert fwd_ok(pred[i], ref, tf), (prec, i, L, np.max(np.abs(pred[i] - ref)))
            assert np.max(np.abs(J[i, :, :L] - refJ)) / np.max(np.abs(refJ)) < tj, (prec, i, L)


def test_forward_device_pointer_path(gpu, systems):
    import torch
    from geobipy_b200.synthetic import synthetic_batch
    b = synthetic_batch(0, 64)
    host = gpu.fdem_forward(systems[0], b["nlayers"], b["sigma"], b["thickness"], b["height"], precision=64)
    dev = gpu.fdem_forward(systems[0], torch.tensor(b["nlayers"], device="cuda"), torch.tensor(b["sigma"], device="cuda"),
                           torch.tensor(b["thickness"], device="cuda"), torch.tensor(b["height"], device="cuda"), precision=64)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)


def test_error_paths(gpu, systems):
    from geobipy_b200 import _lib
    s = gpu.resolve_system_struct()
    s.tid[1] = 5   # y-oriented coil: the reference has no branch for it either (fdem1d_numba.py:57-66)
    with pytest.raises(_lib.GeobipyB200Error):
        gpu.fdem_forward(s, [1], [[0.01]], [[np.inf]], [30.0])
    with pytest.raises(_lib.GeobipyB200Error):
        gpu.fdem_forward(systems[0], [31], np.ones((1, 30)), np.ones((1, 30)), [30.0])   # nlayers out of range
    with pytest.raises(_lib.GeobipyB200Error):
        gpu.rjmcmc_run(systems[0], gpu.make_options(max_layers=31), np.ones((1, 12)), np.array([30.0]))
    # empty batch is a no-op
    out = gpu.rjmcmc_run(systems[0], gpu.make_options(n_markov_chains=10), np.zeros((0, 12)), np.zeros(0))
    assert out["scalars"].shape == (0, 32)


def _merge_centre(h):
    """sigma == sigma_ref sits exactly on the edge between bins 124|125 (round-off decides, in the
    reference too): merge the two bins before comparing hitmaps."""
    h = h.copy()
    h[..., 124, :] += h[..., 125, :]
    h[..., 125, :] = 0
    return h


def test_chain_fp64_is_trajectory_twin_of_oracle(gpu, systems, oracle):
    B, nit = 24, 400
    data, alt = _observed(oracle, systems[1], B)
    opt = gpu.make_options(n_markov_chains=1000)
    oo = oracle.resolve_options(n_markov_chains=1000)
    res = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=2026, first_index=7, max_iterations=nit, precision=64)
    identical = 0
    for b in range(B):
        r = oracle.run_chain(systems[1], oo, data[b], alt[b], 2026, 7 + b, max_iterations=nit)
        s, q = res["scalars"][b], r["scalars"]
        assert s[oracle.S_ITER] == q[oracle.S_ITER] == nit
        assert abs(s[oracle.S_HALFSPACE] / q[oracle.S_HALFSPACE] - 1) < 1e-12
        same = (np.array_equal(res["accept_trace"][b], r["accept_trace"])
                and np.array_equal(_merge_centre(res["hitmap"][b]), _merge_centre(r["hitmap"]))
                and np.array_equal(res["ncells_hist"][b], r["ncells_hist"])
                and np.array_equal(res["edges_hist"][b], r["edges_hist"])
                and np.array_equal(res["rel_hist"][b], r["rel_hist"])
                and np.array_equal(res["add_hist"][b], r["add_hist"]))
        if same:
            identical += 1
            # same decisions; values agree to the round-off the Newton solves amplify (cond(H) ~ 1e4..1e6)
            assert np.allclose(res["misfit_trace"][b, :nit], r["misfit_trace"][:nit], rtol=1e-5)
            assert np.allclose(res["cur_sigma"][b], r["cur_sigma"], rtol=1e-5, equal_nan=True)
            assert np.allclose(res["cur_edges"][b], r["cur_edges"], rtol=1e-9, equal_nan=True)
            assert np.allclose(res["best_sigma"][b], r["best_sigma"], rtol=1e-5, equal_nan=True)
            for j in (oracle.S_N_ACCEPT, oracle.S_CUR_K, oracle.S_BEST_K, oracle.S_BEST_ITER, oracle.S_N_BIRTH,
                      oracle.S_N_DEATH, oracle.S_N_MOVE, oracle.S_N_NONE, oracle.S_N_FORWARD):
                assert s[j] == q[j], j
            for j in (oracle.S_CUR_MISFIT, oracle.S_CUR_PRIOR, oracle.S_CUR_LIKELIHOOD, oracle.S_BEST_POSTERIOR):
                assert abs(s[j] - q[j]) <= 1e-5 * (abs(q[j]) + 1), j
    assert identical >= 0.9 * B, identical


def test_chain_full_termination_rule_and_burn_in(gpu, systems, oracle):
    """Run to the reference's own termination rule (Inference1D.infer :650-677) with a short burn-in window so
    that burn-in, posterior reset and the N + burn + 1 stopping rule are all exercised; fp64 twin of the oracle."""
    B = 8
    data, alt = _observed(oracle, systems[1], B, first=1)
    kw = dict(n_markov_chains=600, burn_in_min_iter=150, update_plot_every=100)
    res = gpu.rjmcmc_run(systems[0], gpu.make_options(**kw), data, alt, seed=5, precision=64)
    oo = oracle.resolve_options(**kw)
    nd = res["hitmap"].shape[2]
    n_burned = 0
    for b in range(B):
        r = oracle.run_chain(systems[1], oo, data[b], alt[b], 5, b)
        s, q = res["scalars"][b], r["scalars"]
        it = int(s[oracle.S_ITER])
        if s[oracle.S_BURNED_IN]:
            n_burned += 1
            assert it == 600 + int(s[oracle.S_BURNED_IN_ITER]) + 1 and s[oracle.S_FAILED] == 0
            counted = it - int(s[oracle.S_BURNED_IN_ITER]) + 1
        else:
            assert it == 600 and s[oracle.S_FAILED] == 1
            counted = it
        if s[oracle.S_N_RESETS] > 0 and not s[oracle.S_BURNED_IN]:
            counted += 1   # Inference1D.update goes on to accumulate the re-initialised model after a reset()
        assert res["hitmap"][b].sum() == counted * nd and (res["hitmap"][b].sum(axis=0) == counted).all()
        assert res["ncells_hist"][b].sum() == counted
        if np.array_equal(res["accept_trace"][b], r["accept_trace"]):
            assert s[oracle.S_BURNED_IN_ITER] == q[oracle.S_BURNED_IN_ITER] and s[oracle.S_ITER] == q[oracle.S_ITER]
            assert np.array_equal(_merge_centre(res["hitmap"][b]), _merge_centre(r["hitmap"]))
    assert n_burned >= 1


def test_chain_edge_cases(gpu, systems, oracle):
    data, alt = _observed(oracle, systems[1], 6)
    # inactive channels (<= 0 or NaN) are dropped from misfit / likelihood (EmDataPoint.active)
    d2 = data.copy()
    d2[0, 3] = -1.0
    d2[1, 7] = np.nan
    d2[2, :] = np.nan          # no active channel: chain does not run, flagged failed
    opt = gpu.make_options(n_markov_chains=300)
    oo = oracle.resolve_options(n_markov_chains=300)
    res = gpu.rjmcmc_run(systems[0], opt, d2, alt, seed=3, max_iterations=150, precision=64)
    assert res["scalars"][2, oracle.S_ITER] == 0 and res["scalars"][2, oracle.S_FAILED] == 1
    for b in (0, 1, 3):
        r = oracle.run_chain(systems[1], oo, np.nan_to_num(d2[b], nan=-1.0), alt[b], 3, b, max_iterations=150)
        assert res["scalars"][b, oracle.S_ITER] == 150
        assert abs(res["scalars"][b, oracle.S_HALFSPACE] / r["scalars"][oracle.S_HALFSPACE] - 1) < 1e-12
        if np.array_equal(res["accept_trace"][b], r["accept_trace"]):
            assert abs(res["scalars"][b, oracle.S_CUR_MISFIT] - r["scalars"][oracle.S_CUR_MISFIT]) < 1e-6 * r["scalars"][oracle.S_CUR_MISFIT]
    # max_layers = 2: births at k = 2 are re-drawn (RectilinearMesh1D.perturb :1047)
    res = gpu.rjmcmc_run(systems[0], gpu.make_options(n_markov_chains=300, max_layers=2), data, alt, seed=4,
                         max_iterations=200, precision=32)
    assert res["ncells_hist"][:, 3:].sum() == 0 and res["ncells_hist"].sum() == 6 * 200
    assert (res["scalars"][:, oracle.S_CUR_K] <= 2).all()


def test_chain_fp32_statistics_match_oracle_ensemble(gpu, systems, oracle, golden_dir):
    """fp32 chains take different accept/reject decisions than the fp64 oracle, so parity is statistical:
    512 GPU chains vs 6 oracle chains (and the 7 reference chains) on the same observed data."""
    g = np.load(os.path.join(golden_dir, "ref_chain_1.npz"))
    B = 512
    data = np.tile(g["data"], (B, 1))
    alt = np.full(B, float(g["altitude"]))
    opt = gpu.make_options(n_markov_chains=10000)
    res = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=77, precision=32,
                         outputs=("hitmap", "ncells_hist", "scalars"))
    sc = res["scalars"]
    acc = sc[:, oracle.S_N_ACCEPT].sum() / sc[:, oracle.S_ITER].sum()
    nc = res["ncells_hist"].sum(axis=0).astype(np.float64)
    kbar = (nc * np.arange(nc.size)).sum() / nc.sum()
    assert np.all(np.abs(sc[:, oracle.S_HALFSPACE] / float(g["halfspace"]) - 1) < 1e-6)
    oo = oracle.resolve_options(n_markov_chains=10000)
    runs = [oracle.run_chain(systems[1], oo, g["data"], float(g["altitude"]), 300 + j, 1) for j in range(6)]
    o_acc = np.mean([r["scalars"][oracle.S_N_ACCEPT] / r["scalars"][oracle.S_ITER] for r in runs])
    o_nc = sum(r["ncells_hist"].astype(np.float64) for r in runs)
    o_kbar = (o_nc * np.arange(o_nc.size)).sum() / o_nc.sum()
    assert abs(acc - o_acc) < 0.05, (acc, o_acc)
    assert abs(kbar - o_kbar) < 0.5, (kbar, o_kbar)
    files = sorted(f for f in os.listdir(golden_dir) if f.startswith("ref_chain_1"))
    refs = [np.load(os.path.join(golden_dir, f)) for f in files]
    r_acc = np.mean([r["accept_trace"].mean() for r in refs])
    assert abs(acc - r_acc) < 0.05, (acc, r_acc)

    def med(h):
        c = np.cumsum(h, axis=0)
        return np.array([np.searchsorted(c[:, j], 0.5 * c[-1, j]) for j in range(h.shape[1])])
    ens = np.array([med(r["hitmap"]) for r in refs] + [med(r["hitmap"]) for r in runs])
    m = med(res["hitmap"].sum(axis=0, dtype=np.int64))[:120]
    inside = (m >= ens.min(axis=0)[:120] - 2) & (m <= ens.max(axis=0)[:120] + 2)
    assert inside.mean() >= 0.9


def _runs_from_result(res, oracle, n):
    """per-chain dicts for tests/posterior_parity.py from a batched result"""
    sc = res["scalars"]
    runs = []
    for b in range(n):
        it = int(sc[b, oracle.S_ITER])
        runs.append(dict(hitmap=res["hitmap"][b], edges_hist=res["edges_hist"][b], ncells_hist=res["ncells_hist"][b],
                         misfit_trace=res["misfit_trace"][b], iterations=it, burned_in=bool(sc[b, oracle.S_BURNED_IN]),
                         acceptance=sc[b, oracle.S_N_ACCEPT] / max(it, 1)))
    return runs


@pytest.mark.parametrize("sidx", [0, 1, 2, 3])
def test_fp32_production_kernel_matches_live_reference_ensembles(gpu, systems, oracle, golden_dir, sidx):
    """The kernel that earns the bench number - rjmcmc_kernel<float, float, 12, 16, FDEM> - against ensembles of chains of
    the LIVE REFERENCE (Inference1D.infer loop, resolve_options, n_markov_chains = 10 000; 7 chains per sounding for
    soundings 0, 2, 3 and 15 for sounding 1, tests/golden/make_golden.py chain) on the same observed data: 256 fp32 GPU chains
    per sounding, the full SURVEY.md 8(d) list with the tolerances of tests/posterior_parity.py: 5 % / 50 % / 95 %
    conductivity profiles above the depth of investigation within 2 bins of the reference envelope, interface peak within
    2 depth cells, mean layer count +-0.5, acceptance +-5 points, misfit after burn-in and burned-in fraction."""
    import posterior_parity as P
    files = sorted(f for f in os.listdir(golden_dir) if f.startswith("ref_chain_%d" % sidx))
    refs = [dict(np.load(os.path.join(golden_dir, f))) for f in files]
    assert len(refs) >= 7
    g = refs[0]
    B = 256
    opt = gpu.make_options(n_markov_chains=10000)
    res = gpu.rjmcmc_run(systems[0], opt, np.tile(g["data"], (B, 1)), np.full(B, float(g["altitude"])), seed=4100 + sidx,
                         first_index=0, precision=32, outputs=("hitmap", "edges_hist", "ncells_hist", "misfit_trace", "scalars"))
    assert np.all(np.abs(res["scalars"][:, oracle.S_HALFSPACE] / float(g["halfspace"]) - 1) < 1e-6)
    m = P.compare(refs, _runs_from_result(res, oracle, B))
    print("sounding", sidx, m)
    P.check(m)


def test_chain_fp32_matches_fp64_ensemble(gpu, systems, oracle):
    """Production path (fp32 forward AND fp32 sampler arithmetic, MUFU log/exp, fp32 Cholesky) against the
    fp64 instantiation (the oracle's trajectory twin) on the same soundings, 384 chains each.  The yardstick
    is the fp64 path's own seed-to-seed scatter: the fp32-vs-fp64 discrepancy of every statistic must not
    exceed 2.5x the discrepancy between two fp64 ensembles with different seeds (plus a small floor):
    acceptance rate, mean layer count, layer-count distribution (total variation), and the posterior-mean
    conductivity bin per depth cell over the top 60 m (RMS; bin width = 8 sigma_prior / 250 = 0.077 ln-units)."""
    nrep = 384
    for sidx in (1, 5):
        data, alt = _observed(oracle, systems[1], 1, first=sidx)
        d = np.tile(data, (nrep, 1))
        a = np.full(nrep, alt[0])
        opt = gpu.make_options(n_markov_chains=10000)
        st = {}
        for tag, prec, seed in (("f32", 32, 4242), ("f64", 64, 4242), ("f64b", 64, 777)):
            r = gpu.rjmcmc_run(systems[0], opt, d, a, seed=seed, precision=prec, outputs=("hitmap", "ncells_hist", "scalars"))
            sc = r["scalars"]
            nc = r["ncells_hist"].sum(axis=0).astype(np.float64)
            hm = r["hitmap"].sum(axis=0, dtype=np.int64)[:, :120].astype(np.float64)
            st[tag] = dict(acc=sc[:, oracle.S_N_ACCEPT].sum() / sc[:, oracle.S_TOTAL_ITER].sum(), nc=nc / nc.sum(),
                           kbar=(nc * np.arange(nc.size)).sum() / nc.sum(),
                           mean_bin=(hm * np.arange(hm.shape[0])[:, None]).sum(axis=0) / hm.sum(axis=0))

        def dist(x, y):
            return dict(acc=abs(x["acc"] - y["acc"]), kbar=abs(x["kbar"] - y["kbar"]),
                        tv=0.5 * np.abs(x["nc"] - y["nc"]).sum(),
                        rms=float(np.sqrt(np.mean((x["mean_bin"] - y["mean_bin"]) ** 2))))
        d32, d64 = dist(st["f32"], st["f64"]), dist(st["f64b"], st["f64"])
        floor = dict(acc=0.01, kbar=0.05, tv=0.02, rms=0.25)
        for key in floor:
            assert d32[key] <= 2.5 * d64[key] + floor[key], (sidx, key, d32, d64)


def test_full_size_properties(gpu, systems):
    """BASELINE config-2 batch size (4096 soundings) through size-independent invariants, device-pointer path."""
    import torch
    from geobipy_b200.synthetic import synthetic_batch
    B, nit = 4096, 60
    b = synthetic_batch(0, 256)
    sig = torch.tensor(np.tile(b["sigma"], (16, 1)), device="cuda")
    thk = torch.tensor(np.tile(b["thickness"], (16, 1)), device="cuda")
    nl = torch.tensor(np.tile(b["nlayers"], 16), device="cuda")
    alt = torch.tensor(np.tile(b["height"], 16), device="cuda")
    data = gpu.fdem_forward(systems[0], nl, sig, thk, alt, precision=64)
    opt = gpu.make_options(n_markov_chains=1000)
    res = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=1, max_iterations=nit, precision=32,
                         outputs=("hitmap", "ncells_hist", "edges_hist", "rel_hist", "add_hist", "accept_trace", "scalars"))
    torch.cuda.synchronize()
    sc = res["scalars"].cpu().numpy()
    assert (sc[:, 0] == nit).all()
    nd = res["hitmap"].shape[2]
    assert (res["hitmap"].sum(dim=1) == nit).all()                 # every depth cell visited once per iteration
    assert (res["ncells_hist"].sum(dim=1) == nit).all()
    assert (res["rel_hist"].sum(dim=1) == nit).all() and (res["add_hist"].sum(dim=1) == nit).all()
    assert np.array_equal(res["accept_trace"][:, :nit + 1].sum(dim=1).cpu().numpy(), sc[:, 8])
    assert (sc[:, 20:24].sum(axis=1) == nit).all()
    # identical soundings with different sounding indices take different random paths
    assert len(set(sc[::256, 8])) > 1
    # idempotence: same seed -> bit-identical result
    res2 = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=1, max_iterations=nit, precision=32, outputs=("hitmap", "scalars"))
    torch.cuda.synchronize()
    assert torch.equal(res["hitmap"], res2["hitmap"])
    assert nd == 440


@pytest.mark.parametrize("kind", ["fdem", "tdem", "fdem_solve_z", "tdem_solve_z"])
def test_speculative_evaluation_is_bit_identical(gpu, systems, oracle, kind, monkeypatch):
    """Idle warps evaluate future iterations of running chains speculatively (gbp_chain.cuh, "speculative evaluation").
    Per-iteration random sub-streams make that exact: every output of a batch that leaves most warps idle (so that
    speculation is active from the start) is bit-identical with speculation switched off - except the diagnostic
    count of speculated iterations."""
    from geobipy_b200 import _lib
    if kind == "fdem":
        system, opt = systems[0], gpu.make_options(n_markov_chains=3000, update_plot_every=500, burn_in_min_iter=500)
        data, alt = _observed(oracle, systems[1], 40)
    elif kind == "fdem_solve_z":   # sampled sensor height: the proposed height comes back from the speculating warp
        system, opt = systems[0], gpu.make_options(n_markov_chains=3000, update_plot_every=500, burn_in_min_iter=500,
                                                   solve_height=1, max_height_change=1.0, height_prop_var=0.01)
        data, alt = _observed(oracle, systems[1], 40)
        alt = alt + 0.4
    else:
        from geobipy_b200.synthetic import synthetic_batch, skytem_noise_std
        hk = dict(solve_height=1, max_height_change=1.0, height_prop_var=0.01) if kind == "tdem_solve_z" else {}
        system, opt = gpu.skytem_survey_struct(), gpu.make_options(n_markov_chains=3000, update_plot_every=500, burn_in_min_iter=500, **gpu.SKYTEM_OPTIONS, **hk)
        tsys = oracle.make_tdem_system()
        b = synthetic_batch(0, 40, max_depth=400.0, n_channels=45)
        data = np.zeros((40, 45))
        for i in range(40):
            L = int(b["nlayers"][i])
            clean = oracle.tdem_forward(tsys, b["height"][i], b["sigma"][i, :L], b["thickness"][i, :L])
            data[i] = clean + b["noise"][i] * skytem_noise_std(clean, np.array(tsys.t_centre[:45]), (26, 19))
        alt = b["height"] + (0.4 if kind == "tdem_solve_z" else 0.0)   # the proposed height's geometry set comes back from the speculating warp
    out, nspec = {}, {}
    for helpers, minrej in (("0", "24"), ("12", "4"), ("27", "0")):
        monkeypatch.setenv("GBP_SPEC_HELPERS", helpers)
        monkeypatch.setenv("GBP_SPEC_MIN_REJECTIONS", minrej)
        gpu.debug_counters(reset=True)
        out[helpers] = gpu.rjmcmc_run(system, opt, data, alt, seed=99, precision=32)
        nspec[helpers] = int(gpu.debug_counters()[8])
    ref = out["0"]
    assert nspec["0"] == 0
    for helpers in ("12", "27"):
        r = out[helpers]
        assert nspec[helpers] > 0.05 * r["scalars"][:, _lib.S_TOTAL_ITER].sum()
        for name in ref:
            assert np.array_equal(ref[name], r[name], equal_nan=True), (kind, helpers, name)


def test_summarise_kernel_matches_host_reference(gpu):
    """gbp_summarise_hitmap (hand-written, one thread per depth cell) against the torch/numpy statement of
    Mesh._mean / Mesh._percentile (geobipy_b200/parallel.py CPU branch; reference classes/mesh/Mesh.py:80, :173-217)."""
    import torch
    from geobipy_b200.parallel import summarise_hitmap
    rng = np.random.default_rng(4)
    B, ns, nd = 37, 250, 440
    h = rng.integers(0, 40, (B, ns, nd)).astype(np.int32)
    h[3] = 0                                   # an empty hitmap (failed chain)
    h[5, :, 7] = 0
    h[6, 100:, :] = 0
    edges = np.linspace(-9.6, 9.6, ns + 1) + 0.3
    cpu = summarise_hitmap(torch.tensor(h), torch.tensor(edges))
    dev = summarise_hitmap(torch.tensor(h, device="cuda"), torch.tensor(edges, device="cuda"))
    torch.cuda.synchronize()
    for k in cpu:
        assert torch.allclose(dev[k].cpu(), cpu[k], rtol=1e-12, atol=1e-12), k
    # per-sounding bin origins (ln sigma_ref differs per sounding)
    lo = torch.tensor(rng.uniform(-8, -3, B), device="cuda")
    mean, pct = gpu.summarise_hitmap(torch.tensor(h, device="cuda"), lo, 0.0768, (50.0,))
    mean0, pct0 = gpu.summarise_hitmap(torch.tensor(h, device="cuda"), torch.zeros(B, dtype=torch.float64, device="cuda"), 0.0768, (50.0,))
    nz = torch.tensor(h.sum(axis=1) > 0, device="cuda")
    assert torch.allclose((mean - mean0)[nz], lo.unsqueeze(1).expand(B, nd)[nz], atol=1e-9)
    assert torch.allclose(pct[0] - pct0[0], lo.unsqueeze(1).expand(B, nd), atol=1e-12)


@pytest.mark.parametrize("kind", ["fdem", "tdem"])
def test_ten_thousand_random_models_fp32_vs_fp64(gpu, systems, oracle, kind):
    """the fp32 kernels over >= 10^4 random models spanning sigma in [1e-4, 10] S/m and L in [1, 30],
    against the fp64 kernels (the oracle's twins; a 200-model subset is checked against the oracle itself)."""
    rng = np.random.default_rng(77)
    B = 10240
    nl = rng.integers(1, 31, B).astype(np.int32)
    sig = 10.0 ** rng.uniform(-4.0, 1.0, (B, 30))
    thk = 10.0 ** rng.uniform(0.0, 2.0, (B, 30))
    alt = rng.uniform(25.0, 45.0, B)
    if kind == "fdem":
        system, osys, fwd = systems[0], systems[1], oracle.fdem_forward
    else:
        system, osys, fwd = gpu.skytem_survey_struct(), oracle.make_tdem_system(), oracle.tdem_forward
    p64, J64 = gpu.forward(system, nl, sig, thk, alt, precision=64, sensitivity=True)
    p32, J32 = gpu.forward(system, nl, sig, thk, alt, precision=32, sensitivity=True)
    for b in range(0, B, B // 200):
        L = int(nl[b])
        ref = fwd(osys, alt[b], sig[b, :L], thk[b, :L])
        assert np.all(np.abs(p64[b] - ref) <= 1e-7 * np.abs(ref) + (5e-8 if kind == "fdem" else 1e-24)), b
    if kind == "fdem":
        assert fwd_ok(p32, p64, F32_FWD)
        err = np.abs(J32 - J64).max(axis=(1, 2)) / np.abs(J64).max(axis=(1, 2))
        assert err.max() < 5 * F32_J and np.median(err) < F32_J / 10
    else:
        t = np.array(osys.t_centre[:45])
        floor = np.r_[np.full(26, 2e-14), np.full(19, 2e-13)] * np.sqrt(1e-3 / t)
        sn = np.sqrt((0.05 * p64) ** 2 + floor ** 2)
        assert np.all(np.abs(p32 - p64) <= 2e-4 * np.abs(p64) + 0.01 * sn)
        assert np.median(np.abs(p32 / p64 - 1.0)) < 2e-6
        # Jacobian rows: 3 % of the row maximum, with an absolute floor of 1e-3 noise standard deviations (rows of
        # late windows over resistive ground are orders of magnitude below the noise; d/d ln(sigma) has the data's units)
        rowmax = np.abs(J64).max(axis=2)
        dJ = np.abs(J32 - J64).max(axis=2)
        ok = dJ <= 3e-2 * rowmax + 1e-3 * sn
        assert ok.all(), (float((dJ / (3e-2 * rowmax + 1e-3 * sn)).max()), int((~ok).sum()))
        assert np.median(dJ / rowmax) < 1e-4


def test_team_mode_is_bit_identical(gpu, systems, oracle, monkeypatch):
    """Teams of warps sharing their forward evaluations (gbp_chain.cuh team_round; the production default is 8 warps per
    team) must not change a single bit of the results: every frequency of a forward is still summed by one warp in the
    same order.  Team sizes 1 (off), 2, 4, 8 and both member mappings / unit granularities, on a batch that has more
    chains than resident warps (chains are claimed dynamically and teams dissolve at the end) and one that has fewer."""
    opt = gpu.make_options(n_markov_chains=10000)
    for B, nit in ((2500, 200), (37, 900)):
        data, alt = _observed(oracle, systems[1], min(B, 64))
        reps = -(-B // data.shape[0])
        data, alt = np.tile(data, (reps, 1))[:B], np.tile(alt, reps)[:B]
        ref = None
        for team, spread, fu in ((1, 1, 1), (2, 0, 0), (4, 0, 1), (4, 1, 0), (8, 1, 1)):
            monkeypatch.setenv("GBP_TEAM", str(team))
            monkeypatch.setenv("GBP_TEAM_SPREAD", str(spread))
            monkeypatch.setenv("GBP_TEAM_FREQ_UNITS", str(fu))
            r = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=5, max_iterations=nit, precision=32)
            if ref is None:
                ref = r
                continue
            for k in ref:
                assert np.array_equal(ref[k], r[k], equal_nan=True), (B, team, spread, fu, k)


def test_device_summaries_match_reference_histogram(gpu, golden_dir):
    """SURVEY.md 8(f) rank 2 on the device: gbp_summarise_posterior + gbp_opacity_doi against what the reference's OWN
    Histogram / Mesh methods return on recorded hitmaps (tests/golden/posterior_summaries.npz): mean, median, mode, 5 % /
    95 % percentiles (1e-12 relative), credible range in decades, transparency, opacity (1e-12 absolute), opacity level;
    the depth of investigation and the per-line normalisation (Inference2D.compute_opacity / compute_doi, which need an
    HDF5 file in the reference) against the host restatement dataset.opacity_and_doi."""
    import torch
    from geobipy_b200 import dataset
    g = np.load(os.path.join(golden_dir, "posterior_summaries.npz"))
    P = lambda n, k: g["s%d_%s" % (n, k)]   # noqa: E731
    hm = torch.tensor(np.stack([P(n, "hitmap") for n in range(3)]).astype(np.int32), device="cuda")
    ln_edges = np.stack([np.log(P(n, "x_edges")) for n in range(3)])
    lo = torch.tensor(ln_edges[:, 0], device="cuda")
    dx = float((ln_edges[0, -1] - ln_edges[0, 0]) / 250)
    r = gpu.summarise_posterior(hm, lo, dx, percentiles=(5.0, 50.0, 95.0), credible_percent=90.0)
    for n in range(3):
        cs = np.cumsum(P(n, "hitmap"), axis=0).astype(np.float64)
        for name, pc, mine in (("mean", None, r["mean"][n]), ("p5", 5.0, r["pct"][0, n]), ("median", None, r["pct"][1, n]),
                               ("p95", 95.0, r["pct"][2, n]), ("mode", None, r["mode"][n])):
            mine = np.exp(mine.cpu().numpy())
            # Histogram.percentile normalises by the grand total first: at an EXACT tie of a cumulative count its bin
            # depends on that round-off (tests/test_dataset.py); everything computed from the counts is exact
            tie = np.zeros(mine.shape, bool) if pc is None else (np.abs(cs - pc * 0.01 * cs[-1]) < 1e-9 * np.maximum(cs[-1], 1.0)).any(axis=0)
            assert np.allclose(mine[~tie], P(n, name)[~tie], rtol=1e-12, atol=0.0), (n, name)
            assert np.all(np.abs(np.log(mine[tie] / P(n, name)[tie])) <= dx * (1 + 1e-9)), (n, name)
        assert np.allclose(r["credible_range"][n].cpu().numpy(), P(n, "credible_range90"), rtol=0.0, atol=1e-12), n
    # per sounding normalisation = Histogram.transparency / opacity of one hitmap
    op, doi, _ = gpu.opacity_doi(r["range_bins"], doi_percent=67.0, level_percent=95.0)
    for n in range(3):
        assert np.allclose(op[n].cpu().numpy(), P(n, "opacity90"), rtol=0.0, atol=1e-12), n
        assert np.allclose(1.0 - op[n].cpu().numpy(), P(n, "transparency90"), rtol=0.0, atol=1e-12), n
    # opacity level of the 95 % credible range
    r95 = gpu.summarise_posterior(hm, lo, dx, percentiles=(50.0,), credible_percent=95.0)
    _, _, lvl = gpu.opacity_doi(r95["range_bins"], level_percent=95.0)
    for n in range(3):
        yc = 0.5 * (P(n, "y_edges")[1:] + P(n, "y_edges")[:-1])
        assert yc[int(lvl[n])] == float(P(n, "opacity_level95")), n
    # the flight line: the three soundings normalised together + depth of investigation
    grp = torch.zeros(3, dtype=torch.int32, device="cuda")
    op_l, doi_l, _ = gpu.opacity_doi(r["range_bins"], group=grp, n_groups=1, doi_percent=67.0)
    ref_op, ref_doi = dataset.opacity_and_doi_from_range(np.stack([P(n, "credible_range90") for n in range(3)]), P(0, "y_edges"),
                                                         doi_percent=67.0)
    assert np.allclose(op_l.cpu().numpy(), ref_op, rtol=0.0, atol=1e-12)
    yc = 0.5 * (P(0, "y_edges")[1:] + P(0, "y_edges")[:-1])
    assert np.array_equal(yc[doi_l.cpu().numpy()], ref_doi)
    # empty hitmap: every summary falls on the last bin, as the reference's clip
    z = gpu.summarise_posterior(torch.zeros((1, 250, 440), dtype=torch.int32, device="cuda"), lo[:1], dx)
    assert bool((z["range_bins"] == 0).all()) and np.allclose(z["pct"][1].cpu().numpy(), ln_edges[0, 0] + 249.5 * dx)
