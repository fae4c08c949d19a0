"""GPU parity tests (-m gpu) of the sampled sensor height (the options file's solve_z; SURVEY.md section 8 "next" row 4:
Point.perturb pointcloud/Point.py:614-622, set_priors :959-961, set_proposals :977-979, set_z_posterior :1013-1020).

The oracle's restatement is pinned on transitions and chains recorded from the live reference
(tests/test_oracle_golden.py::test_height_*); here the CUDA path is compared with the oracle:
  fp64 chains : same random stream -> identical accept / reject trajectories, identical height histograms
  fp32 chains : the same posterior height within the seed-to-seed scatter of the fp64 chains
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HEIGHT = dict(solve_height=1, max_height_change=1.0, height_prop_var=0.01)
BIAS = 0.4   # the height handed to the inversion is this far above the one the data were simulated at


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
    return data, b["height"] + BIAS


def _mean_dz(hist, dz=1.0):
    c = -dz + (np.arange(hist.shape[-1]) + 0.5) * (2 * dz / hist.shape[-1])
    return (hist * c).sum(axis=-1) / np.maximum(hist.sum(axis=-1), 1)


def test_height_chain_fp64_is_trajectory_twin_of_oracle(gpu, systems, oracle):
    """Short burn-in so that the burn-in reset of the height histogram and the N + burn + 1 rule are exercised."""
    B = 16
    data, alt = _observed(oracle, systems[1], B)
    kw = dict(n_markov_chains=500, burn_in_min_iter=150, update_plot_every=100)
    res = gpu.rjmcmc_run(systems[0], gpu.make_options(**kw, **HEIGHT), data, alt, seed=77, first_index=3, precision=64)
    assert "height_hist" in res
    oo = oracle.resolve_options(**kw, **HEIGHT)
    identical = moved = 0
    for b in range(B):
        r = oracle.run_chain(systems[1], oo, data[b], alt[b], 77, 3 + b)
        s, q = res["scalars"][b], r["scalars"]
        assert abs(s[oracle.S_HALFSPACE] / q[oracle.S_HALFSPACE] - 1) < 1e-12
        # the histogram holds one visit per counted iteration, inside the prior
        assert res["height_hist"][b].sum() == res["ncells_hist"][b].sum() == res["rel_hist"][b].sum()
        assert abs(s[oracle.S_CUR_HEIGHT] - alt[b]) <= 1.0 and abs(s[oracle.S_BEST_HEIGHT] - alt[b]) <= 1.0
        moved += s[oracle.S_CUR_HEIGHT] != alt[b]
        if np.array_equal(res["accept_trace"][b], r["accept_trace"]):
            identical += 1
            assert s[oracle.S_ITER] == q[oracle.S_ITER] and s[oracle.S_BURNED_IN_ITER] == q[oracle.S_BURNED_IN_ITER]
            assert np.array_equal(res["height_hist"][b], r["height_hist"])
            assert np.array_equal(res["ncells_hist"][b], r["ncells_hist"])
            assert np.array_equal(res["rel_hist"][b], r["rel_hist"]) and np.array_equal(res["add_hist"][b], r["add_hist"])
            assert abs(s[oracle.S_CUR_HEIGHT] - q[oracle.S_CUR_HEIGHT]) < 1e-9
            assert abs(s[oracle.S_BEST_HEIGHT] - q[oracle.S_BEST_HEIGHT]) < 1e-9
            for j in (oracle.S_N_ACCEPT, oracle.S_CUR_K, oracle.S_BEST_ITER, oracle.S_N_FORWARD, oracle.S_N_RESETS):
                assert s[j] == q[j], j
            for j in (oracle.S_CUR_MISFIT, oracle.S_CUR_PRIOR, oracle.S_CUR_LIKELIHOOD, oracle.S_BEST_POSTERIOR):
                assert abs(s[j] - q[j]) <= 1e-5 * (abs(q[j]) + 1), j
    assert identical >= 0.8 * B, identical
    assert moved >= B - 1


def test_height_off_leaves_the_fixed_height_path_alone(gpu, systems, oracle):
    """solve_height = 0: no height histogram by default, the height scalars are the input heights, and the other
    height options are ignored."""
    data, alt = _observed(oracle, systems[1], 4)
    a = gpu.rjmcmc_run(systems[0], gpu.make_options(n_markov_chains=300), data, alt, seed=9, max_iterations=200, precision=64)
    b = gpu.rjmcmc_run(systems[0], gpu.make_options(n_markov_chains=300, max_height_change=3.0, height_prop_var=1.0),
                       data, alt, seed=9, max_iterations=200, precision=64)
    assert "height_hist" not in a
    assert np.array_equal(a["scalars"][:, oracle.S_CUR_HEIGHT], alt) and np.array_equal(a["scalars"][:, oracle.S_BEST_HEIGHT], alt)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_height_fp32_matches_fp64_posterior_height(gpu, systems, oracle):
    """The fp32 build samples the same posterior height as the fp64 one: over 64 replicas of sounding 5 (burns in at
    once, height well determined) the mean posterior height offsets agree within 3 standard errors + 0.03 m."""
    n = 64
    d1, a1 = _observed(oracle, systems[1], 1, first=5)
    data, alt = np.repeat(d1, n, axis=0), np.repeat(a1, n)
    opt = gpu.make_options(n_markov_chains=3000, burn_in_min_iter=1000, update_plot_every=500, **HEIGHT)
    m = {}
    for prec in (32, 64):
        res = gpu.rjmcmc_run(systems[0], opt, data, alt, seed=11 + prec, precision=prec,
                             outputs=("height_hist", "scalars", "ncells_hist"))
        ok = res["scalars"][:, oracle.S_BURNED_IN] == 1
        assert ok.mean() > 0.8
        assert (res["height_hist"].sum(axis=1) == res["ncells_hist"].sum(axis=1)).all()
        m[prec] = _mean_dz(res["height_hist"][ok])
    se = np.hypot(m[32].std() / np.sqrt(m[32].size), m[64].std() / np.sqrt(m[64].size))
    assert abs(m[32].mean() - m[64].mean()) <= 3 * se + 0.03, (m[32].mean(), m[64].mean(), se)
    assert abs(m[64].mean()) > 0.2   # the data did move the height away from the (biased) input


def test_height_time_domain_is_refused(gpu):
    from geobipy_b200 import _lib
    sv = gpu.skytem_survey_struct()
    opt = gpu.make_options(**dict(gpu.SKYTEM_OPTIONS, n_markov_chains=100), **HEIGHT)
    with pytest.raises(_lib.GeobipyB200Error, match="solve_height"):
        gpu.rjmcmc_run(sv, opt, np.full((1, gpu.n_channels(sv)), 1e-12), np.array([30.0]), precision=32)
