"""GPU tests of the reference-shaped object interface (geobipy_b200/api.py) - they read like the
reference's own usage (documentation_source/source/examples/Datapoints/plot_resolve_datapoint.py,
Inference_1D/plot_inference_1d_resolve.py, tests/test_synthetic_data.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(built_lib):
    from geobipy_b200 import _lib, api
    _lib.require_cuda()
    return api


def _resolve(api):
    t = api.CircularLoop(orientation=list("zzxzzz"), moment=[1, 1, -1, 1, 1, 1], x=[0] * 6, y=[0] * 6, z=[0] * 6)
    r = api.CircularLoop(orientation=list("zzxzzz"), moment=[1] * 6, x=[7.93, 7.91, 9.03, 7.91, 7.91, 7.89], y=[0] * 6, z=[0] * 6)
    return api.FdemSystem([380.0, 1776.0, 3345.0, 8171.0, 41020.0, 129550.0], t, r)


def test_system_read_and_properties(api, tmp_path):
    p = tmp_path / "resolve.stm"
    p.write_text("freq, tor, tmom, tx, ty, tzoff, ror, rmom, rx, ry, rzoff\n"
                 "380, z, 1, 0, 0, 0, z, 1, 7.93, 0, 0\n1776, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0\n"
                 "3345, x, -1, 0, 0, 0, x, 1, 9.03, 0, 0\n8171, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0\n"
                 "41020, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0\n129550, z, 1, 0, 0, 0, z, 1, 7.89, 0, 0\n")
    s = api.FdemSystem.read(str(p))
    assert list(s.tensor_id) == [9, 9, 1, 9, 9, 9]
    assert np.allclose(s.loop_separation, [7.93, 7.91, 9.03, 7.91, 7.91, 7.89])
    assert s.lamda0.shape == (6, 120) and s.lamda1.shape == (6, 140)
    assert bytes(s.c_struct) == bytes(_resolve(api).c_struct)


def test_datapoint_forward_matches_reference_golden(api, golden_dir):
    """test_resolve of the reference (tests/test_synthetic_data.py:16-30): the 'glacial' golden CSV."""
    g = np.load(os.path.join(golden_dir, "resolve_clean.npz"))
    dp = api.FdemDataPoint(x=0.0, y=0.0, z=float(g["height"]), elevation=0.0, system=_resolve(api))
    for i in (0, 40, 78):
        mod = api.Model(api.RectilinearMesh1D(edges=np.r_[0.0, g["zwedge"][i], g["zdeep"][i], np.inf]), g["sigma"][0])
        dp.forward(mod)
        assert np.allclose(dp.predictedData, g["data"][0, i])
        J = dp.sensitivity(mod)
        assert J.shape == (12, 3)
        dp.fm_dlogc(mod)
        assert np.array_equal(dp.sensitivity_matrix, J)


def test_datapoint_error_conventions(api):
    dp = api.FdemDataPoint(z=30.0, system=_resolve(api))
    with pytest.raises(AssertionError):   # FdemDataPoint.py:541: last edge must be inf
        dp.forward(api.Model(api.RectilinearMesh1D(edges=[0.0, 5.0, 10.0]), [0.01, 0.1]))
    with pytest.raises(AssertionError):   # fdem1d.py:29: sensor below the top of the model
        api.FdemDataPoint(z=-1.0, system=_resolve(api)).forward(api.Model(api.RectilinearMesh1D(edges=[0.0, np.inf]), [0.01]))
    with pytest.raises(AssertionError):
        api.FdemDataPoint(z=30.0, system="not a system object".split())


def test_inference1d_runs_like_the_reference(api, oracle):
    """BASELINE configs[0]: one RESOLVE sounding, 3-layer 'glacial' model, 1000 iterations.  As in the
    reference, 1000 iterations never burn in (> 5000 needed), so the sounding is reported failed while the
    histograms are still filled."""
    system = _resolve(api)
    true = api.Model(api.RectilinearMesh1D(edges=[0.0, 5.0, 7.5, np.inf]), [1e-2, 1e-1, 0.03333333])
    dp = api.FdemDataPoint(z=30.0, system=system)
    dp.forward(true)
    dp.data[:] = dp.predictedData
    options = dict(n_markov_chains=1000, update_plot_every=5000, solve_parameter=False, solve_gradient=True,
                   maximum_number_of_layers=30, minimum_depth=0.1, maximum_depth=200.0, minimum_thickness=1.0,
                   initial_relative_error=0.05, minimum_relative_error=0.001, maximum_relative_error=0.5,
                   initial_additive_error=5.0, minimum_additive_error=3.0, maximum_additive_error=20.0,
                   relative_error_proposal_variance=1e-6, additive_error_proposal_variance=1e-6,
                   probability_of_birth=1 / 6, probability_of_death=1 / 6, probability_of_perturb=1 / 6,
                   probability_of_no_change=0.5, covariance_scaling=1.0, solve_relative_error=True,
                   solve_additive_error=True, interactive_plot=False, save_hdf5=True)
    inf = api.Inference1D(prng=np.random.default_rng(0), precision=64, **options)
    inf.initialize(dp)
    failed = inf.infer(None)
    assert failed and not inf.burned_in and inf.iteration == 1000
    assert abs(inf.halfspace - 0.03199267) < 1e-7          # the reference's best half-space for this sounding
    assert inf.hitmap.counts.shape == (250, 440) and inf.hitmap.counts.sum() == 440 * 1000
    assert inf.n_cells_posterior.counts.sum() == 1000 and inf.data_misfit_v.shape == (2000,)
    assert 30.0 < inf.acceptance_rate < 90.0                 # reference: 60.9 % on this sounding
    assert inf.data_misfit < 124.6                           # started from the half-space misfit 124.59
    med = inf.hitmap.median()
    assert med.shape == (440,) and np.all(med > 0)
    assert abs(inf.interface_probability().sum() - 1.0) < 1e-12 or inf.edges_posterior.counts.sum() == 0
    # the stored predicted data are those of the final model
    chk = api.FdemDataPoint(z=30.0, system=system)
    chk.forward(inf.model)
    assert np.allclose(chk.predictedData, dp.predictedData)


def test_infer_batch(api, oracle):
    from geobipy_b200.synthetic import synthetic_batch
    system = _resolve(api)
    b = synthetic_batch(0, 32)
    from geobipy_b200 import ops
    clean = ops.fdem_forward(system.c_struct, b["nlayers"], b["sigma"], b["thickness"], b["height"], precision=64)
    data = clean + b["noise"] * np.sqrt((0.05 * clean) ** 2 + 25.0)
    r = api.infer_batch(system, data, b["height"], seed=9, n_markov_chains=500, max_iterations=200)
    assert (r["scalars"][:, 0] == 200).all() and r["hitmap"].shape == (32, 250, 440)
    # batch composition does not change a sounding's result (stream = (seed, sounding index))
    r2 = api.infer_batch(system, data[5:9], b["height"][5:9], seed=9, first_index=5, n_markov_chains=500, max_iterations=200)
    assert np.array_equal(r2["hitmap"], r["hitmap"][5:9]) and np.array_equal(r2["scalars"], r["scalars"][5:9])


def test_inference1d_solve_z(api, oracle):
    """The options-file keys of the sampled sensor height (solve_z / maximum_z_change / z_proposal_variance,
    Point.set_priors pointcloud/Point.py:959-961) through Inference1D: height posterior on the prior's bins, the
    datapoint left at the last sampled height, predicted data consistent with it."""
    system = _resolve(api)
    true = api.Model(api.RectilinearMesh1D(edges=[0.0, 5.0, 7.5, np.inf]), [1e-2, 1e-1, 0.03333333])
    dp = api.FdemDataPoint(z=30.0, system=system)
    dp.forward(true)
    dp.data[:] = dp.predictedData
    dp.z = 30.5   # inverted with a height 0.5 m off
    inf = api.Inference1D(prng=np.random.default_rng(0), precision=64, n_markov_chains=1000, solve_z=True,
                          maximum_z_change=1.0, z_proposal_variance=0.01)
    inf.initialize(dp)
    inf.infer(None)
    h = inf.height_posterior
    assert h.counts.shape == (99,) and h.counts.sum() == inf.n_cells_posterior.counts.sum() == 1000
    assert np.allclose(h.x_edges[[0, -1]], [29.5, 31.5])
    assert 29.5 <= dp.z <= 31.5 and dp.z != 30.5 and 29.5 <= inf.best_height <= 31.5
    chk = api.FdemDataPoint(z=dp.z, system=system)
    chk.forward(inf.model)
    assert np.allclose(chk.predictedData, dp.predictedData)


def test_inference1d_exposes_the_paths_writeHdf_serialises(api, oracle):
    """Inference1D.writeHdf (inversion/Inference1D.py:1050-1090) walks model.values.posterior, model.mesh.nCells.posterior,
    model.mesh.edges.posterior, datapoint.relative_error / additive_error / z .posterior, best_model, best_datapoint,
    multiplier: the mirror exposes exactly those paths."""
    system = _resolve(api)
    true = api.Model(api.RectilinearMesh1D(edges=[0.0, 5.0, 7.5, np.inf]), [1e-2, 1e-1, 0.03333333])
    dp = api.FdemDataPoint(z=30.0, system=system)
    dp.forward(true)
    dp.data[:] = dp.predictedData
    inf = api.Inference1D(prng=np.random.default_rng(0), n_markov_chains=600, multiplier=1.5, solve_z=True,
                          maximum_z_change=1.0, z_proposal_variance=0.01)
    inf.initialize(dp)
    inf.infer(None)
    m = inf.model
    assert m.values.hasPosterior and m.values.posterior.counts.shape == (250, 440)
    assert m.values.posterior is inf.hitmap is m.posterior
    assert int(m.mesh.nCells) == m.nCells == m.values.size
    assert m.mesh.nCells.posterior.counts.sum() == 600 and m.mesh.nCells.posterior is inf.n_cells_posterior
    assert m.mesh.edges.posterior.counts.shape == (440,)
    d = inf.datapoint
    assert d.relative_error.posterior.counts.sum() == 600 and d.additive_error.posterior.counts.sum() == 600
    assert d.z.posterior.counts.sum() == 600 and d.z.posterior is inf.height_posterior
    assert float(inf.multiplier) == 1.5
    b = inf.best_datapoint
    assert b is not d and float(b.z) == inf.best_height
    assert np.allclose(b.relative_error, inf.best_relative_error) and np.allclose(b.additive_error, inf.best_additive_error)
    chk = api.FdemDataPoint(z=float(b.z), system=system)
    chk.forward(inf.best_model)
    assert np.allclose(chk.predictedData, b.predictedData)
    # slicing / arithmetic keep plain-array behaviour
    assert np.all(np.asarray(m.values) * 2.0 == 2.0 * m.values) and (m.mesh.edges[1:] - m.mesh.edges[:-1]).shape == (m.nCells,)


def test_torch_ops_match_the_python_entry_points(api, oracle):
    """torch.ops.geobipy_b200.*  run the same C-ABI calls as geobipy_b200.ops."""
    import torch
    from geobipy_b200 import _lib, ops, torch_ops
    from geobipy_b200.synthetic import synthetic_batch
    dev = torch.device("cuda")
    sysc = ops.resolve_system_struct()
    b = synthetic_batch(0, 16)
    t = {k: torch.tensor(v, device=dev) for k, v in b.items()}
    S = torch_ops.pod(sysc)
    pred = torch.ops.geobipy_b200.fdem_forward(S, t["nlayers"], t["sigma"], t["thickness"], t["height"], 64)
    assert torch.equal(pred, ops.fdem_forward(sysc, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=64))
    p2, J = torch.ops.geobipy_b200.fdem_sensitivity(S, t["nlayers"], t["sigma"], t["thickness"], t["height"], 32)
    q2, K = ops.fdem_forward(sysc, t["nlayers"], t["sigma"], t["thickness"], t["height"], precision=32, sensitivity=True)
    assert torch.equal(p2, q2) and torch.equal(J, K) and J.shape == (16, 12, 30)
    opt = ops.make_options(n_markov_chains=400)
    data = (pred + t["noise"] * torch.sqrt((0.05 * pred) ** 2 + 25.0)).contiguous()
    out = torch.ops.geobipy_b200.rjmcmc_run(S, torch_ops.pod(opt), data, t["height"], 7, 0, 150, 32)
    ref = ops.rjmcmc_run(sysc, opt, data, t["height"], seed=7, max_iterations=150, precision=32)
    got = dict(zip(torch_ops.RJMCMC_OUTPUTS, out))
    for k, v in ref.items():
        assert torch.equal(got[k], v, ) or (v.dtype.is_floating_point and torch.allclose(got[k], v, equal_nan=True)), k
    assert got["height_hist"].numel() == 0 and got["hitmap"].shape == (16, 250, 440)
    with pytest.raises(RuntimeError):
        torch.ops.geobipy_b200.fdem_forward(S[:-1], t["nlayers"], t["sigma"], t["thickness"], t["height"], 64)


def test_integration_adapter_at_the_numba_seam(api, golden_dir):
    """`nbFdem1dfwd / nbFdem1dsen` with the reference's positional signature over the C-ABI,
    called with the arguments fdem1d.py:39-56 / :116-133 builds, against the live Numba kernels' outputs
    (tests/golden/fdem_random_models.npz)."""
    from geobipy_b200 import _lib, ops

    def _system(tid, frequencies, tHeight, rHeight, moments, rx, separation, scale, altitude):
        F = len(frequencies)
        tor, ror = (np.asarray(tid) - 1) % 3, (np.asarray(tid) - 1) // 3
        tz = np.asarray(tHeight) - altitude
        rz = np.asarray(rHeight) + np.asarray(tHeight)
        ry = np.sqrt(np.maximum(np.asarray(separation) ** 2 - np.asarray(rx) ** 2 - (rz - tz) ** 2, 0.0))
        rmom = np.asarray(scale) / np.asarray(moments)
        return ops.make_system_struct(frequencies, tor, moments, np.zeros(F), np.zeros(F), tz, ror, rmom, rx, ry, rz)

    def nbFdem1dfwd(tid, frequencies, tHeight, rHeight, moments, rx, separation, w0, lamda0, lamda02, w1, lamda1,
                    lamda12, scale, conductivity, susceptibility, permeability, thickness, altitude=None):
        altitude = tHeight[0] if altitude is None else altitude
        s = _system(tid, frequencies, tHeight, rHeight, moments, rx, separation, scale, altitude)
        L, F = len(conductivity), len(frequencies)
        out = ops.fdem_forward(s, np.int32([L]), np.asarray(conductivity, float)[None], np.asarray(thickness, float)[None],
                               np.asarray([altitude], float), precision=_lib.PRECISION_F64)[0]
        return out[:F] + 1j * out[F:]

    def nbFdem1dsen(tid, frequencies, tHeight, rHeight, moments, rx, separation, w0, lamda0, lamda02, w1, lamda1,
                    lamda12, scale, conductivity, susceptibility, permeability, thickness, altitude=None):
        altitude = tHeight[0] if altitude is None else altitude
        s = _system(tid, frequencies, tHeight, rHeight, moments, rx, separation, scale, altitude)
        L, F = len(conductivity), len(frequencies)
        _, J = ops.fdem_forward(s, np.int32([L]), np.asarray(conductivity, float)[None], np.asarray(thickness, float)[None],
                                np.asarray([altitude], float), precision=_lib.PRECISION_F64, sensitivity=True)
        J = J[0, :, :L]
        return J[:F] + 1j * J[F:]

    g = np.load(os.path.join(golden_dir, "fdem_random_models.npz"))
    sysm = _resolve(api)
    tmom, rmom = np.asarray(sysm.transmitter.moment, float), np.asarray(sysm.receiver.moment, float)
    for i in (0, 17, 101, 255):
        L = int(g["nlayers"][i])
        alt = float(g["height"][i])
        tH = alt + np.asarray(sysm.transmitter.z, float)           # fdem1d.py:31-34
        rH = -tH + np.asarray(sysm.receiver.z, float)
        args = (sysm.tensor_id, sysm.frequencies, tH, rH, tmom, np.asarray(sysm.receiver.x, float) - np.asarray(sysm.transmitter.x, float),
                sysm.loop_separation, None, None, None, None, None, None, tmom * rmom, g["sigma"][i, :L], np.zeros(L), np.zeros(L),
                g["thickness"][i, :L])
        f = nbFdem1dfwd(*args)
        ref = g["forward"][i]
        assert np.max(np.abs(np.r_[f.real, f.imag] - ref) / (np.abs(ref) + 1.0)) < 5e-8
        J = nbFdem1dsen(*args)
        refJ = g["sensitivity"][i, :, :L]
        assert J.shape == (6, L) and np.max(np.abs(np.vstack([J.real, J.imag]) - refJ)) / np.max(np.abs(refJ)) < 1e-8


def test_inference1d_writes_the_reference_hdf_layout(api, oracle, tmp_path):
    """`Inference1D.createHdf(parent, add_axis)` + `.infer(hdf_file_handle)` as the reference's survey driver calls them
    (Inference3D._create_HDF5_dataset :312-340, infer_serial :488-492): the sounding lands in its row (found by fiducial)
    of a file in the reference's layout (tests/test_hdf.py pins the layout itself on the reference's own writer)."""
    from geobipy_b200 import _lib, h5lite
    system = _resolve(api)
    true = api.Model(api.RectilinearMesh1D(edges=[0.0, 5.0, 7.5, np.inf]), [1e-2, 1e-1, 0.03333333])
    dp = api.FdemDataPoint(z=30.0, system=system, fiducial=20.0, lineNumber=7.0)
    dp.forward(true)
    dp.data[:] = dp.predictedData
    inf = api.Inference1D(prng=np.random.default_rng(0), n_markov_chains=500)
    inf.initialize(dp)
    path = str(tmp_path / "7.h5")
    with h5lite.File(path, "w") as f:
        inf.createHdf(f, add_axis=np.asarray([10.0, 20.0, 30.0]))
        f["data/fiducial/data"][:] = [10.0, 20.0, 30.0]          # Inference2D.createHdf :2011-2012
        inf.infer(f)
    f = h5lite.File(path, "r")
    assert f["iteration"][1] == 500 and f["iteration"][0] == 0 and np.isnan(f["halfspace/data"][0])
    assert np.isclose(f["halfspace/data"][1], inf.halfspace) and f["n_markov_chains"][()] == 500
    assert np.array_equal(f["model/values/posterior/values/data"][1], inf.hitmap.counts)
    k = inf.best_model.nCells
    assert f["model/mesh/nCells/data"][1] == k and np.allclose(f["model/values/data"][1, :k], inf.best_model.values)
    assert np.allclose(f["data/predicted_data/data"][1], inf.best_datapoint.predictedData)
    assert np.array_equal(f["data/relative_error/posterior/values/data"][1], inf.datapoint.relative_error.posterior.counts)
    assert f["data/line_number/data"][1] == 7.0 and f["data"].attrs["repr"] == "FdemData" and f["model"].attrs["repr"] == "Model"
