"""GPU parity tests (-m gpu) of the time-domain (SkyTEM) path: the CUDA kernels, called through the C-ABI,
against the oracle and the reference's committed golden vectors.

Tolerances (north_star: "stated fp64 -> fp32 tolerance"):
  fp64 instantiation : forward 5e-9 relative (+1e-22 absolute); Jacobian 1e-8 of the row maximum
  fp32 instantiation : |d - ref| <= 2e-4 |ref| + 0.01 sigma_n, sigma_n = the reference's noise model for these
                       data (5 %, additive 2e-14 / 2e-13 V/Am^4 at 1 ms scaled by t^-1/2); Jacobian rows 2e-2 of
                       the row maximum (late-time rows are ~1e4 below the early ones in fp32 dynamic range)
  fp64 chains        : same Philox stream as the oracle -> identical accept/reject trajectories
  fp32 chains        : agree with the fp64 chains in acceptance and final state over a short run
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built_lib):
    from geobipy_b200 import _lib, ops
    _lib.require_cuda()
    return ops


@pytest.fixture(scope="module")
def systems(gpu, oracle):
    return gpu.skytem_survey_struct(), oracle.make_tdem_system()


def _floor(osys):
    t = np.array(osys.t_centre[:osys.C])
    return np.r_[np.full(osys.n_win[0], 2e-14), np.full(osys.n_win[1], 2e-13)] * np.sqrt(1e-3 / t)


def _random_models(B, seed=3):
    rng = np.random.default_rng(seed)
    nl = rng.integers(1, 31, B).astype(np.int32)
    nl[:8] = [1, 1, 2, 2, 3, 3, 30, 30]
    sig = 10.0 ** rng.uniform(-3.5, 0.5, (B, 30))
    thk = rng.uniform(1.0, 40.0, (B, 30))
    alt = rng.uniform(25.0, 45.0, B)
    return nl, sig, thk, alt


@pytest.mark.parametrize("prec", [64, 32])
def test_forward_reference_csv_goldens(gpu, systems, golden_dir, prec):
    """The reference's SkyTEM known-answer vectors through the GPU path: same stated tolerance as the oracle's
    own pin (tests/test_oracle_golden.py::test_tdem_forward_matches_reference_csv_goldens)."""
    g = np.load(os.path.join(golden_dir, "skytem_clean.npz"))
    sig = np.repeat(g["sigma"][:, None, :], 79, axis=1).reshape(-1, 3)
    thk = np.tile(np.stack([g["zwedge"], g["zdeep"] - g["zwedge"], np.full(79, np.inf)], axis=1), (6, 1))
    out = gpu.forward(systems[0], np.full(474, 3, np.int32), sig, thk, np.full(474, float(g["geometry"][0])), precision=prec)
    ref = g["data"].reshape(-1, 45)
    E = np.abs(out / ref - 1.0)
    Z = np.abs(out - ref) / np.sqrt((0.05 * ref) ** 2 + _floor(systems[1]) ** 2)
    assert np.median(E) < 2e-3
    assert (E < 0.01).mean() > 0.90
    assert np.all((E < 0.03) | (Z < 0.75))


@pytest.mark.parametrize("prec", [64, 32])
def test_forward_and_jacobian_random_models(gpu, systems, oracle, prec):
    nl, sig, thk, alt = _random_models(192)
    B = len(nl)
    ref = np.zeros((B, 45))
    refJ = np.zeros((B, 45, 30))
    for b in range(B):
        L = nl[b]
        ref[b] = oracle.tdem_forward(systems[1], alt[b], sig[b, :L], thk[b, :L])
        refJ[b, :, :L] = oracle.tdem_sensitivity(systems[1], alt[b], sig[b, :L], thk[b, :L])
    pred, J = gpu.forward(systems[0], nl, sig, thk, alt, precision=prec, sensitivity=True)
    pred2 = gpu.forward(systems[0], nl, sig, thk, alt, precision=prec)
    rowmax = np.abs(refJ).max(axis=2)
    if prec == 64:
        assert np.all(np.abs(pred - ref) <= 5e-9 * np.abs(ref) + 1e-22)
        assert np.all(np.abs(pred2 - ref) <= 5e-9 * np.abs(ref) + 1e-22)
        assert np.all(np.abs(J - refJ).max(axis=2) <= 1e-8 * rowmax)
    else:
        sn = np.sqrt((0.05 * ref) ** 2 + _floor(systems[1]) ** 2)
        for p in (pred, pred2):
            assert np.all(np.abs(p - ref) <= 2e-4 * np.abs(ref) + 0.01 * sn)
        assert np.median(np.abs(pred2 / ref - 1.0)) < 2e-6
        assert np.all(np.abs(J - refJ).max(axis=2) <= 2e-2 * rowmax)
        assert np.median(np.abs(J - refJ).max(axis=2) / rowmax) < 1e-5
    for b in range(B):
        assert not J[b, :, nl[b]:].any()


def test_device_pointer_entry_points_match_host_path(gpu, systems):
    import torch
    nl, sig, thk, alt = _random_models(64, seed=8)
    ph, Jh = gpu.forward(systems[0], nl, sig, thk, alt, precision=64, sensitivity=True)
    dev = torch.device("cuda:0")
    pd, Jd = gpu.forward(systems[0], torch.tensor(nl, device=dev), torch.tensor(sig, device=dev), torch.tensor(thk, device=dev),
                         torch.tensor(alt, device=dev), precision=64, sensitivity=True)
    torch.cuda.synchronize()
    assert np.array_equal(pd.cpu().numpy(), ph) and np.array_equal(Jd.cpu().numpy(), Jh)


def _observed(gpu, systems, oracle, n, first=0):
    from geobipy_b200.synthetic import synthetic_batch, skytem_noise_std
    b = synthetic_batch(first, n, max_depth=400.0, n_channels=45)
    data = np.zeros((n, 45))
    t = np.array(systems[1].t_centre[:45])
    for i in range(n):
        L = int(b["nlayers"][i])
        clean = oracle.tdem_forward(systems[1], b["height"][i], b["sigma"][i, :L], b["thickness"][i, :L])
        data[i] = clean + b["noise"][i] * skytem_noise_std(clean, t, (26, 19))
    return data, b["height"]


def test_chain_fp64_is_trajectory_twin_of_oracle(gpu, systems, oracle):
    """Same Philox stream, same arithmetic: hitmaps, traces and per-system error histograms identical."""
    B, NIT = 12, 400
    data, alt = _observed(gpu, systems, oracle, B)
    data[3, 5] = -1.0          # inactive channels (EmDataPoint.active): negative and NaN
    data[4, 40] = np.nan
    opt = gpu.make_options(n_markov_chains=2000, **gpu.SKYTEM_OPTIONS)
    oo = oracle.skytem_options(n_markov_chains=2000)
    res = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=21, max_iterations=NIT, precision=64)
    assert res["rel_hist"].shape == (B, 2, 99) and res["hitmap"].shape == (B, 250, 1209)
    same = 0
    for b in range(B):
        r = oracle.run_chain(systems[1], oo, data[b], alt[b], 21, b, max_iterations=NIT)
        s, q = res["scalars"][b], r["scalars"]
        assert abs(s[oracle.S_HALFSPACE] - q[oracle.S_HALFSPACE]) <= 1e-12 * q[oracle.S_HALFSPACE]
        assert res["hitmap"][b].sum() == r["hitmap"].sum() == NIT * 1209
        ok = (np.array_equal(res["hitmap"][b], r["hitmap"]) and np.array_equal(res["accept_trace"][b], r["accept_trace"])
              and np.array_equal(res["rel_hist"][b], r["rel_hist"]) and np.array_equal(res["add_hist"][b], r["add_hist"])
              and np.array_equal(res["ncells_hist"][b], r["ncells_hist"]) and np.array_equal(res["edges_hist"][b], r["edges_hist"]))
        if ok:
            for k in (oracle.S_CUR_REL, oracle.S_CUR_ADD, oracle.S_CUR_REL2, oracle.S_CUR_ADD2, oracle.S_CUR_MISFIT,
                      oracle.S_CUR_LIKELIHOOD, oracle.S_CUR_PRIOR, oracle.S_BEST_POSTERIOR):
                assert abs(s[k] - q[k]) <= 1e-8 * abs(q[k]) + 1e-300, (b, k)
            assert np.allclose(res["misfit_trace"][b], r["misfit_trace"], rtol=1e-8)
        same += ok
    assert same >= B - 1, same


def test_chain_fp32_agrees_with_fp64_short_run(gpu, systems, oracle):
    B, NIT = 16, 300
    data, alt = _observed(gpu, systems, oracle, B, first=100)
    opt = gpu.make_options(n_markov_chains=2000, **gpu.SKYTEM_OPTIONS)
    r64 = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=4, max_iterations=NIT, precision=64)
    r32 = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=4, max_iterations=NIT, precision=32)
    a64, a32 = r64["scalars"][:, 8], r32["scalars"][:, 8]          # S_N_ACCEPT
    assert (r32["scalars"][:, 0] == NIT).all() and not r32["scalars"][:, 7].any()
    assert abs(a64.mean() - a32.mean()) < 0.15 * a64.mean() + 5
    # the half-space search (100 forwards per chain) lands on the same conductivity
    assert np.allclose(r32["scalars"][:, 6], r64["scalars"][:, 6], rtol=1e-6)
    # errors are reported in data units whatever the internal scaling of the fp32 path
    assert np.all((r32["scalars"][:, 13] > 1e-16) & (r32["scalars"][:, 13] < 1e-10))
    assert np.all((r32["scalars"][:, 26] > 1e-16) & (r32["scalars"][:, 26] < 1e-10))
    assert r32["hitmap"].sum() == B * NIT * 1209


def test_chain_runs_to_termination_fp32(gpu, systems, oracle):
    """A full (short-option) run: termination rule, burn-in bookkeeping, posterior mass."""
    B = 32
    data, alt = _observed(gpu, systems, oracle, B, first=300)
    opt = gpu.make_options(n_markov_chains=3000, update_plot_every=500, burn_in_min_iter=500, **gpu.SKYTEM_OPTIONS)
    r = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=9, precision=32, outputs=("scalars", "hitmap", "ncells_hist", "rel_hist"))
    s = r["scalars"]
    burned = s[:, 1] > 0
    assert burned.mean() > 0.5
    # burned-in chains keep the models of iterations b .. b + N + 1 (Inference1D.infer runs while
    # iteration <= N + b, :650-677): N + 2 entries in every histogram
    assert np.all(r["ncells_hist"][burned].sum(axis=1) == 3000 + 2)
    assert np.all(r["rel_hist"][burned].sum(axis=2) == 3000 + 2)
    assert np.all(r["hitmap"][burned].sum(axis=(1, 2)) == (3000 + 2) * 1209)
    # misfit of burned-in chains is of the order of the number of active channels
    assert np.median(s[burned, 14]) < 3 * 45


def test_chain_fp32_matches_fp64_ensemble(gpu, systems, oracle):
    """Production path (fp32 forward and fp32 sampler arithmetic in 2^40-scaled data units) against the fp64
    instantiation (the oracle's trajectory twin) on the same sounding, 192 chains each run to the reference's
    termination rule.  Yardstick = the fp64 path's own seed-to-seed scatter: the fp32-vs-fp64 discrepancy of every
    statistic must not exceed 2.5x the discrepancy between two fp64 ensembles with different seeds (plus a small
    floor): acceptance rate, mean layer count, layer-count distribution (total variation), posterior-mean conductivity
    bin per depth cell over the top 150 m (RMS, bin width 0.077 ln-units), and the posterior means of the four error
    parameters (in histogram bins)."""
    nrep = 192
    data, alt = _observed(gpu, systems, oracle, 1, first=2)
    d = np.tile(data, (nrep, 1))
    a = np.full(nrep, alt[0])
    opt = gpu.make_options(n_markov_chains=4000, update_plot_every=1000, burn_in_min_iter=1000, **gpu.SKYTEM_OPTIONS)
    st = {}
    for tag, prec, seed in (("f32", 32, 4242), ("f64", 64, 4242), ("f64b", 64, 777)):
        r = gpu.rjmcmc_run(systems[0], opt, d, a, seed=seed, precision=prec, outputs=("hitmap", "ncells_hist", "rel_hist", "add_hist", "scalars"))
        sc = r["scalars"]
        nc = r["ncells_hist"].sum(axis=0).astype(np.float64)
        hm = r["hitmap"].sum(axis=0, dtype=np.int64)[:, :300].astype(np.float64)
        eh = np.concatenate([r["rel_hist"].sum(axis=0), r["add_hist"].sum(axis=0)]).astype(np.float64)   # [4, 99]
        st[tag] = dict(acc=sc[:, oracle.S_N_ACCEPT].sum() / sc[:, oracle.S_TOTAL_ITER].sum(), nc=nc / nc.sum(),
                       kbar=(nc * np.arange(nc.size)).sum() / nc.sum(),
                       mean_bin=(hm * np.arange(hm.shape[0])[:, None]).sum(axis=0) / hm.sum(axis=0),
                       err_bin=(eh * np.arange(99)).sum(axis=1) / eh.sum(axis=1), burned=sc[:, oracle.S_BURNED_IN].mean())
    assert st["f64"]["burned"] > 0.5

    def dist(x, y):
        return dict(acc=abs(x["acc"] - y["acc"]), kbar=abs(x["kbar"] - y["kbar"]), tv=0.5 * np.abs(x["nc"] - y["nc"]).sum(),
                    rms=float(np.sqrt(np.mean((x["mean_bin"] - y["mean_bin"]) ** 2))),
                    err=float(np.max(np.abs(x["err_bin"] - y["err_bin"]))), burned=abs(x["burned"] - y["burned"]))
    d32, d64 = dist(st["f32"], st["f64"]), dist(st["f64b"], st["f64"])
    floor = dict(acc=0.01, kbar=0.05, tv=0.02, rms=0.25, err=0.5, burned=0.08)
    for key in floor:
        assert d32[key] <= 2.5 * d64[key] + floor[key], (key, d32, d64)


def test_fp32_production_kernel_matches_live_reference_ensemble(gpu, systems, oracle, golden_dir):
    """rjmcmc_kernel<float, float, 48, 12, TDEM> (the production time-domain kernel) against the 6 chains of the LIVE
    REFERENCE on sounding 2 (dual-moment TdemDataPoint, skytem_options at n_markov_chains = 10 000, through
    tests/golden/fake_gatdaem1d.py): 128 fp32 GPU chains, the SURVEY.md 8(d) list of tests/posterior_parity.py (profiles
    over the top 100 m inside the reference envelope +-2 bins - tails of a 6-chain ensemble are noisy, so 0.7 / 0.9 /
    0.8 of the cells -, interface peak within 5 cells of 0.5 m, mean layer count +-0.75, acceptance +-5 points), the
    post-burn-in misfit centred on the number of active channels (0.5-1.5 x 45, as the reference's own chains:
    35-39), the same chain length (burn-in at 5001) and the posterior means of the four error parameters within 2 bins."""
    import posterior_parity as P
    files = sorted(f for f in os.listdir(golden_dir) if f.startswith("ref_tdem_chain_2"))
    refs = [dict(np.load(os.path.join(golden_dir, f))) for f in files]
    assert len(refs) >= 6
    g = refs[0]
    B = 128
    opt = gpu.make_options(n_markov_chains=10000, **gpu.SKYTEM_OPTIONS)
    res = gpu.rjmcmc_run(systems[0], opt, np.tile(g["data"], (B, 1)), np.full(B, float(g["altitude"])), seed=5200, precision=32,
                         outputs=("hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "misfit_trace", "scalars"))
    sc = res["scalars"]
    assert np.all(np.abs(sc[:, oracle.S_HALFSPACE] / float(g["halfspace"]) - 1) < 1e-6)
    runs = []
    for b in range(B):
        it = int(sc[b, oracle.S_ITER])
        runs.append(dict(hitmap=res["hitmap"][b], edges_hist=res["edges_hist"][b], ncells_hist=res["ncells_hist"][b],
                         misfit_trace=res["misfit_trace"][b], iterations=it, burned_in=bool(sc[b, oracle.S_BURNED_IN]),
                         acceptance=sc[b, oracle.S_N_ACCEPT] / max(it, 1)))
    m = P.compare(refs, runs, top=200)
    print("skytem sounding 2", m)
    P.check(m, inside=(0.7, 0.9, 0.8), peak_cells=5, layers=0.75)
    burned = sc[:, oracle.S_BURNED_IN] > 0
    assert burned.mean() >= 0.9
    assert np.median(sc[burned, oracle.S_BURNED_IN_ITER]) == int(g["burned_in_iteration"])       # 5001
    for b in np.flatnonzero(burned)[:32]:
        b0, it = int(sc[b, oracle.S_BURNED_IN_ITER]), int(sc[b, oracle.S_ITER])
        assert 0.5 * 45 < res["misfit_trace"][b, b0:it].mean() < 1.5 * 45
    k = np.arange(99)
    for name in ("rel_hist", "add_hist"):
        rr = sum(r[name].astype(np.int64) for r in refs)
        oo = res[name].sum(axis=0, dtype=np.int64)
        assert np.all(np.abs((rr * k).sum(axis=1) / rr.sum(axis=1) - (oo * k).sum(axis=1) / oo.sum(axis=1)) < 2.0), name


def test_full_size_properties(gpu, systems):
    """A 4096-sounding batch (the bench's size) through size-independent invariants, device-pointer path."""
    import torch
    from geobipy_b200.synthetic import synthetic_batch
    B, nit = 4096, 50
    b = synthetic_batch(0, 256, max_depth=400.0, n_channels=45)
    sig = torch.tensor(np.tile(b["sigma"], (16, 1)), device="cuda")
    thk = torch.tensor(np.tile(b["thickness"], (16, 1)), device="cuda")
    nl = torch.tensor(np.tile(b["nlayers"], 16), device="cuda")
    alt = torch.tensor(np.tile(b["height"], 16), device="cuda")
    data = gpu.forward(systems[0], nl, sig, thk, alt, precision=64)
    assert bool((data > 0).all())                                   # dBz/dt of a layered earth: positive windows
    # linearity of the window operator in the frequency response: halving every conductivity is NOT linear, but
    # the forward is invariant under splitting a layer in two equal-conductivity halves
    sig2 = torch.cat([sig[:, :1], sig], dim=1)[:, :30].contiguous()
    thk2 = torch.cat([0.5 * thk[:, :1], 0.5 * thk[:, :1], thk[:, 1:]], dim=1)[:, :30].contiguous()
    ok = (nl >= 2) & (nl < 30)      # the first layer of a multi-layer model has a finite thickness
    d2 = gpu.forward(systems[0], (nl + 1).to(torch.int32), sig2, thk2, alt, precision=64)
    assert torch.allclose(d2[ok], data[ok], rtol=1e-9, atol=0.0)
    opt = gpu.make_options(n_markov_chains=1000, **gpu.SKYTEM_OPTIONS)
    res = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=1, max_iterations=nit, precision=32,
                         outputs=("hitmap", "ncells_hist", "rel_hist", "add_hist", "accept_trace", "scalars"))
    torch.cuda.synchronize()
    sc = res["scalars"].cpu().numpy()
    assert (sc[:, 0] == nit).all()
    assert (res["hitmap"].sum(dim=1) == nit).all()                 # every depth cell visited once per iteration
    assert (res["ncells_hist"].sum(dim=1) == nit).all()
    assert (res["rel_hist"].sum(dim=2) == nit).all() and (res["add_hist"].sum(dim=2) == nit).all()
    assert np.array_equal(res["accept_trace"][:, :nit + 1].sum(dim=1).cpu().numpy(), sc[:, 8])
    assert (sc[:, 20:24].sum(axis=1) == nit).all()
    assert len(set(sc[::256, 8])) > 1                               # same data, different sounding index: different paths
    res2 = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=1, max_iterations=nit, precision=32, outputs=("hitmap", "scalars"))
    torch.cuda.synchronize()
    assert torch.equal(res2["hitmap"], res["hitmap"])               # idempotence: same seed -> bit-identical


def test_single_moment_datapoint_twin(gpu, oracle):
    """A time-domain datapoint with ONE system (high moment only): scalar error options, 1-D error proposals -
    fp64 chains are again trajectory twins of the oracle."""
    defs = oracle.skytem_definitions()[:1]
    osys = oracle.make_tdem_system(defs)
    sv = gpu.make_tdem_survey_struct(gpu.skytem_definitions()[:1])
    assert gpu.n_channels(sv) == 26 and osys.C == 26 and osys.n_sys == 1
    rng = np.random.default_rng(12)
    B, NIT = 6, 300
    alt = rng.uniform(28.0, 40.0, B)
    data = np.zeros((B, 26))
    t = np.array(osys.t_centre[:26])
    for b in range(B):
        sig = 10.0 ** rng.uniform(-2.5, -0.5, 3)
        clean = oracle.tdem_forward(osys, alt[b], sig, [20.0, 40.0, 1.0])
        data[b] = clean * (1.0 + 0.03 * rng.standard_normal(26)) + 1e-14 * np.sqrt(1e-3 / t) * rng.standard_normal(26)
    kw = dict(min_edge=1.0, max_edge=550.0, min_width=1.0, covariance_scaling=0.5, rel_init=0.05, rel_min=0.005, rel_max=0.5,
              rel_prop_var=1e-6, add_init=2e-14, add_min=1e-16, add_max=1e-10, add_prop_var=1e-5)
    opt = gpu.make_options(n_markov_chains=2000, **kw)
    oo = oracle.resolve_options(n_markov_chains=2000, **kw)
    assert opt.n_systems <= 1
    pred = gpu.forward(sv, np.full(B, 1, np.int32), np.full((B, 1), 0.02), np.ones((B, 1)), alt, precision=64)
    for b in range(B):
        ref = oracle.tdem_forward(osys, alt[b], [0.02], [1.0])
        assert np.all(np.abs(pred[b] - ref) <= 5e-9 * np.abs(ref))
    res = gpu.rjmcmc_run(sv, opt, data, alt, seed=5, max_iterations=NIT, precision=64)
    assert res["rel_hist"].shape == (B, 99)
    same = 0
    for b in range(B):
        r = oracle.run_chain(osys, oo, data[b], alt[b], 5, b, max_iterations=NIT)
        same += (np.array_equal(res["hitmap"][b], r["hitmap"]) and np.array_equal(res["accept_trace"][b], r["accept_trace"])
                 and np.array_equal(res["rel_hist"][b], r["rel_hist"].reshape(-1)) and np.array_equal(res["add_hist"][b], r["add_hist"].reshape(-1)))
    assert same >= B - 1, same
    r32 = gpu.rjmcmc_run(sv, opt, data, alt, seed=5, max_iterations=NIT, precision=32, outputs=("scalars",))
    assert (r32["scalars"][:, 0] == NIT).all()


# ------------------------------------------------------------------------------------------ Tempest: X + Z components, B field
TEMPEST_ADDITIVE = np.r_[0.011474, 0.012810, 0.008507, 0.005154, 0.004742, 0.004477, 0.004168, 0.003539, 0.003352, 0.003213, 0.003161,
                         0.003122, 0.002587, 0.002038, 0.002201, 0.007383, 0.005693, 0.005178, 0.003659, 0.003426, 0.003046, 0.003095,
                         0.003247, 0.002775, 0.002627, 0.002460, 0.002178, 0.001754, 0.001405, 0.001283]   # tempest_options [fT]


@pytest.fixture(scope="module")
def tempest(gpu, oracle):
    return gpu.tempest_survey_struct(), oracle.make_tdem_system([gpu.tempest_definition()], rx_offset=(-107.0, 0.0, -45.0))


@pytest.mark.parametrize("prec", [64, 32])
def test_tempest_forward_reference_csv_goldens(gpu, tempest, golden_dir, prec):
    """The reference's Tempest known-answer vectors (X and Z components of the B field of a point dipole, 474 soundings x 30
    windows + the primary field) through the forward operator: the stated tolerance of the oracle's own pin
    (tests/test_oracle_golden.py::test_tdem_forward_matches_tempest_csv_goldens)."""
    g = np.load(os.path.join(golden_dir, "tempest_clean.npz"))
    sv = gpu.make_tdem_survey_struct([gpu.tempest_definition()], tuple(g["geometry"][4:7]))
    assert gpu.n_channels(sv) == 30
    assert np.allclose(gpu.tdem_primary_field(sv), g["primary"][0, 0], rtol=1e-12)
    sig = np.repeat(g["sigma"][:, None, :], 79, axis=1).reshape(-1, 3)
    thk = np.tile(np.stack([g["zwedge"], g["zdeep"] - g["zwedge"], np.full(79, np.inf)], axis=1), (6, 1))
    out = gpu.forward(sv, np.full(474, 3, np.int32), sig, thk, np.full(474, float(g["geometry"][0])), precision=prec)
    ref = g["data"].reshape(-1, 30)
    E = np.abs(out / ref - 1.0)
    Z = np.abs(out - ref) / TEMPEST_ADDITIVE
    assert np.median(E) < 2e-3 and np.median(E[:, :15]) < 2e-3 and np.median(E[:, 15:]) < 2e-3
    assert (E < 0.01).mean() > 0.95
    assert np.all((E < 0.03) | (Z < 0.75))


@pytest.mark.parametrize("prec", [64, 32])
def test_tempest_forward_and_jacobian_against_oracle(gpu, tempest, oracle, prec):
    rng = np.random.default_rng(8)
    B = 96
    nl = rng.integers(1, 31, B).astype(np.int32)
    nl[:6] = [1, 2, 3, 5, 30, 30]
    sig = 10.0 ** rng.uniform(-3.5, 0.5, (B, 30))
    thk = rng.uniform(1.0, 40.0, (B, 30))
    alt = rng.uniform(90.0, 150.0, B)
    ref, refJ = np.zeros((B, 30)), np.zeros((B, 30, 30))
    for b in range(B):
        L = nl[b]
        ref[b] = oracle.tdem_forward(tempest[1], alt[b], sig[b, :L], thk[b, :L])
        refJ[b, :, :L] = oracle.tdem_sensitivity(tempest[1], alt[b], sig[b, :L], thk[b, :L])
    pred, J = gpu.forward(tempest[0], nl, sig, thk, alt, precision=prec, sensitivity=True)
    pred2 = gpu.forward(tempest[0], nl, sig, thk, alt, precision=prec)
    assert np.array_equal(pred, pred2) or np.allclose(pred, pred2, rtol=1e-6)
    rowmax = np.abs(refJ).max(axis=2)
    if prec == 64:
        assert np.max(np.abs(pred - ref) / (np.abs(ref) + 1e-9)) < 5e-9
        assert np.max(np.abs(J - refJ).max(axis=2) / rowmax) < 1e-8
    else:   # fp32: 2e-4 |d| + 1 % of the additive noise level of tempest_options
        assert np.all(np.abs(pred - ref) <= 2e-4 * np.abs(ref) + 0.01 * TEMPEST_ADDITIVE)
        assert np.max(np.abs(J - refJ).max(axis=2) / rowmax) < 2e-2
    for b in range(B):   # layers beyond the model's are zero columns
        assert np.all(J[b, :, nl[b]:] == 0.0)


def test_tempest_datapoint_and_sampler_refusal(gpu, tempest, golden_dir):
    """Tempest_datapoint mirror (data = secondary + primary per component, Tempest_datapoint.py:107-127); the sampler refuses an
    X / B-field system that comes without the Tempest error model."""
    from geobipy_b200 import _lib, api, tdem
    g = np.load(os.path.join(golden_dir, "tempest_clean.npz"))
    system = tdem.TdemSystem(definition=gpu.tempest_definition())
    assert system.components == ['x', 'z'] and system.nTimes == 15
    tx = tdem.TdemLoop(x=0.0, y=0.0, z=120.0)
    rx = tdem.TdemLoop(x=-107.0, y=0.0, z=75.0)
    dp = tdem.Tempest_datapoint(x=0.0, y=0.0, z=120.0, elevation=0.0, system=[system], transmitter_loop=tx, receiver_loop=rx,
                                secondary_field=g["data"][0, 10], primary_field=g["primary"][0, 10])
    assert dp.nChannels == 30 and dp.components == ['x', 'z']
    mod = api.Model(api.RectilinearMesh1D(edges=np.r_[0.0, g["zwedge"][10], g["zdeep"][10], np.inf]), g["sigma"][0])
    dp.forward(mod)
    assert np.allclose(dp.predicted_primary_field, g["primary"][0, 10], rtol=1e-12)
    assert np.median(np.abs(dp.predicted_secondary_field / g["data"][0, 10] - 1.0)) < 2e-3
    assert np.allclose(dp.predictedData[:15], dp.predicted_secondary_field[:15] + g["primary"][0, 10, 0])
    assert np.allclose(dp.data[15:], g["data"][0, 10, 15:] + g["primary"][0, 10, 1])
    J = dp.sensitivity(mod)
    assert J.shape == (30, 3)
    dp.additive_error = TEMPEST_ADDITIVE
    assert np.allclose(dp.std[:15], np.sqrt((0.01 * dp.data[:15]) ** 2 + TEMPEST_ADDITIVE[:15] ** 2))
    with pytest.raises(_lib.GeobipyB200Error, match="needs the Tempest error model"):     # X / B system without its error model
        gpu.rjmcmc_run(tempest[0], gpu.make_options(n_markov_chains=100), g["data"][0, :2], np.full(2, 120.0), max_iterations=10)


# ------------------------------------------------------------------------------------------ the sampler for Tempest datapoints
def _tempest_observed(gpu, oracle, tempest, B, first=0):
    """Synthetic Tempest soundings: the shared true-model generator at 120 m, noise = tempest_options' error model; the
    data handed to the sampler are secondary + primary field (Tempest_datapoint.py:107-118)."""
    from geobipy_b200.synthetic import synthetic_batch
    b = synthetic_batch(first, B, max_depth=400.0, n_channels=30)
    prim = np.repeat(oracle.tdem_primary_field(oracle.tempest_definition(), (-107.0, 0.0, -45.0)), 15)
    data = np.zeros((B, 30))
    for i in range(B):
        L = int(b["nlayers"][i])
        clean = oracle.tdem_forward(tempest[1], 120.0, b["sigma"][i, :L], b["thickness"][i, :L])
        data[i] = clean + prim + b["noise"][i] * np.sqrt((0.001 * (clean + prim)) ** 2 + TEMPEST_ADDITIVE ** 2)
    return data, np.full(B, 120.0)


def test_tempest_chain_fp64_is_trajectory_twin_of_oracle(gpu, tempest, oracle):
    """KIND_TEMPEST in fp64 against the oracle (pinned on 800 live-reference transitions, tests/test_oracle_golden.py::
    test_tempest_transition_terms_match_live_reference): same Philox stream, same arithmetic -> identical hitmaps, traces and
    per-component error histograms."""
    B, NIT = 10, 400
    data, alt = _tempest_observed(gpu, oracle, tempest, B)
    sv = gpu.tempest_survey_struct(additive_level=gpu.TEMPEST_ADDITIVE)
    osys = oracle.make_tempest_system()
    opt = gpu.make_options(**dict(gpu.TEMPEST_OPTIONS, n_markov_chains=2000))
    oo = oracle.tempest_options(n_markov_chains=2000)
    res = gpu.rjmcmc_run(sv, opt, data, alt, seed=33, max_iterations=NIT, precision=64)
    assert res["rel_hist"].shape == (B, 2, 99) and res["hitmap"].shape == (B, 250, 1209)
    same = 0
    for b in range(B):
        r = oracle.run_chain(osys, oo, data[b], alt[b], 33, b, max_iterations=NIT)
        s, q = res["scalars"][b], r["scalars"]
        assert abs(s[oracle.S_HALFSPACE] - q[oracle.S_HALFSPACE]) <= 1e-12 * q[oracle.S_HALFSPACE]
        assert res["hitmap"][b].sum() == r["hitmap"].sum() == NIT * 1209
        ok = (np.array_equal(res["hitmap"][b], r["hitmap"]) and np.array_equal(res["accept_trace"][b], r["accept_trace"])
              and np.array_equal(res["rel_hist"][b], r["rel_hist"]) and np.array_equal(res["add_hist"][b], r["add_hist"])
              and np.array_equal(res["ncells_hist"][b], r["ncells_hist"]) and np.array_equal(res["edges_hist"][b], r["edges_hist"]))
        if ok:
            for k in (oracle.S_CUR_REL, oracle.S_CUR_ADD, oracle.S_CUR_REL2, oracle.S_CUR_ADD2, oracle.S_CUR_MISFIT,
                      oracle.S_CUR_LIKELIHOOD, oracle.S_CUR_PRIOR, oracle.S_BEST_POSTERIOR):
                assert abs(s[k] - q[k]) <= 1e-8 * abs(q[k]) + 1e-300, (b, k)
            assert np.allclose(res["misfit_trace"][b], r["misfit_trace"], rtol=1e-8)
        same += ok
    assert same >= B - 1, same
    # the multiplier never drifts from its initial value (it is re-drawn around it every step: the reference's behaviour)
    assert np.all(np.abs(np.log(res["scalars"][:, [oracle.S_CUR_ADD, oracle.S_CUR_ADD2]])) < 6e-3)


def test_tempest_fp32_production_kernel_against_reference_chains(gpu, tempest, oracle, golden_dir):
    """The fp32 KIND_TEMPEST kernel against full chains of the live reference (Inference1D with a Tempest_datapoint,
    tempest_options at n_markov_chains = 10 000, gatdaem1d replaced by tests/golden/fake_gatdaem1d.py) on the same observed
    data: acceptance +-5 points, mean layer count +-0.75, post-burn-in misfit centred like the reference's, the pooled median
    profile inside the reference envelope +-2 bins for >= 90 % of the top 100 m."""
    files = sorted(f for f in os.listdir(golden_dir) if f.startswith("ref_tempest_chain_1"))
    refs = [np.load(os.path.join(golden_dir, f)) for f in files]
    assert len(refs) >= 5
    g = refs[0]
    sv = gpu.tempest_survey_struct(additive_level=gpu.TEMPEST_ADDITIVE)
    opt = gpu.make_options(**dict(gpu.TEMPEST_OPTIONS, n_markov_chains=10000))
    n = 128
    data = np.tile(g["data"], (n, 1))
    res = gpu.rjmcmc_run(sv, opt, data, np.full(n, 120.0), seed=77, precision=32,
                         outputs=("hitmap", "ncells_hist", "accept_trace", "misfit_trace", "rel_hist", "add_hist", "scalars"))
    s = res["scalars"]
    assert np.allclose(s[:, 6], float(g["halfspace"]), rtol=1e-5)                      # same best half-space
    ref_acc = np.mean([r["accept_trace"].mean() for r in refs])
    acc = (s[:, 8] / s[:, 24]).mean()
    kref = np.mean([(r["ncells_hist"] * np.arange(r["ncells_hist"].size)).sum() / r["ncells_hist"].sum() for r in refs])
    kk = ((res["ncells_hist"] * np.arange(res["ncells_hist"].shape[1])).sum(axis=1) / res["ncells_hist"].sum(axis=1)).mean()
    ref_burn = np.mean([bool(r["burned_in"]) for r in refs])
    burn = s[:, 1].mean()
    print("tempest sounding 1", dict(acceptance_ref=ref_acc, acceptance=acc, kbar_ref=kref, kbar=kk, burned_in_ref=ref_burn, burned_in=burn,
                                     n_ref=len(refs), n=n))
    assert abs(acc - ref_acc) < 0.05 and abs(kk - kref) < 0.75
    assert abs(burn - ref_burn) < 0.45            # 5 reference chains: 2 of them burn in
    # final misfit of the chains that burned in: around the number of active channels, as the reference's
    mis = np.array([res["misfit_trace"][b, int(s[b, 0]) - 1] for b in range(n) if s[b, 1]])
    mref = [float(r["misfit_trace"][-1]) for r in refs if bool(r["burned_in"])]
    assert mis.size > 0 and 10.0 < np.median(mis) < 60.0 and all(10.0 < m < 80.0 for m in mref)
    # pooled median profile (top 100 m = 200 depth cells) against the envelope of the reference chains
    def med(h):
        c = np.cumsum(h, axis=0)
        return (c < 0.5 * c[-1]).sum(axis=0)
    pooled = med(res["hitmap"].sum(axis=0))[:200]
    rm = np.array([med(r["hitmap"])[:200] for r in refs])
    inside = (pooled >= rm.min(axis=0) - 2) & (pooled <= rm.max(axis=0) + 2)
    assert inside.mean() > 0.9, inside.mean()


def test_tempest_speculation_and_precision_consistency(gpu, tempest, oracle, monkeypatch):
    """Speculative evaluation is bit-identical for KIND_TEMPEST too, and fp32 agrees with fp64 over a short run."""
    B, NIT = 24, 300
    data, alt = _tempest_observed(gpu, oracle, tempest, B, first=50)
    sv = gpu.tempest_survey_struct(additive_level=gpu.TEMPEST_ADDITIVE)
    opt = gpu.make_options(**dict(gpu.TEMPEST_OPTIONS, n_markov_chains=2000))
    r32 = gpu.rjmcmc_run(sv, opt, data, alt, seed=4, max_iterations=NIT, precision=32)
    monkeypatch.setenv("GBP_SPEC_MIN_REJECTIONS", "0")
    q32 = gpu.rjmcmc_run(sv, opt, data, alt, seed=4, max_iterations=NIT, precision=32)
    monkeypatch.setenv("GBP_SPEC_HELPERS", "0")
    p32 = gpu.rjmcmc_run(sv, opt, data, alt, seed=4, max_iterations=NIT, precision=32)
    for k in ("hitmap", "accept_trace", "rel_hist", "add_hist", "ncells_hist", "edges_hist"):
        assert np.array_equal(r32[k], q32[k]) and np.array_equal(r32[k], p32[k]), k
    monkeypatch.delenv("GBP_SPEC_HELPERS")
    monkeypatch.delenv("GBP_SPEC_MIN_REJECTIONS")
    r64 = gpu.rjmcmc_run(sv, opt, data, alt, seed=4, max_iterations=NIT, precision=64)
    a64, a32 = r64["scalars"][:, 8], r32["scalars"][:, 8]
    assert (r32["scalars"][:, 0] == NIT).all() and not r32["scalars"][:, 7].any()
    assert abs(a64.mean() - a32.mean()) < 0.15 * a64.mean() + 5
    assert np.allclose(r32["scalars"][:, 6], r64["scalars"][:, 6], rtol=1e-6)
    # errors come back unscaled: relative errors inside their prior, multipliers near 1
    assert np.all((r32["scalars"][:, 12] >= 1e-4) & (r32["scalars"][:, 12] <= 1e-2)) and np.all(np.abs(np.log(r32["scalars"][:, 13])) < 6e-3)
