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
