"""The reference-shaped time-domain classes (geobipy_b200/tdem.py): .stm / CSV readers on the CPU, the object
interface on the GPU.  They read like the reference's own usage (tests/test_synthetic_data.py:32-48 test_skytem,
documentation_source/source/examples/Datapoints/plot_skytem_datapoint.py)."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_stm(path, d):
    """Write a parsed description back in the GA-AEM .stm layout (the reference's SkytemHM.stm structure)."""
    wave = "\n".join("%.9E\t%.6E" % (t, c) for t, c in zip(d["waveform_time"], d["waveform_current"]))
    win = "\n".join("%.6E\t%.6E" % (a, b) for a, b in zip(d["window_start"], d["window_end"]))
    path.write_text(
        "System Begin\n\tName = %s\n\tType = Time Domain\n\tTransmitter Begin\n\t\tNumberOfTurns = 1\n\t\tPeakCurrent   = 1\n"
        "\t\tLoopArea      = 1\n\t\tBaseFrequency = %r\n\t\tWaveformDigitisingFrequency = %r\n\t\tWaveFormCurrent Begin\n%s\n\n"
        "\t\tWaveFormCurrent End\n\tTransmitter End\nReceiver Begin\n\tNumberOfWindows = %d\n\tWindowWeightingScheme = AreaUnderCurve\n"
        "\tWindowTimes Begin\n%s\n\n\t\tWindowTimes End\n\t\tLowPassFilter Begin\n\t\t\tCutOffFrequency = %s\n\t\t\tOrder           = %s\n"
        "\t\tLowPassFilter End\n\tReceiver End\nForwardModelling Begin\n\t\t//TX loop area is was 340.82 m^2 -> r = sqrt(340.82/pi)\n"
        "\t\tModellingLoopRadius = %r\n\t\tOutputType = dB/dt\n\t\tXOutputScaling = 0\n\t\tYOutputScaling = 0\n\t\tZOutputScaling = 1\n"
        "\t\tSecondaryFieldNormalisation  =  none\n\t\tFrequenciesPerDecade = 5\n\t\tNumberOfAbsiccaInHankelTransformEvaluation = %d\n"
        "\tForwardModelling End\n\nSystem End\n" % (
            d["name"], d["base_frequency"], d["digitising_frequency"], wave, len(d["window_start"]), win,
            " ".join(repr(c) for c in d["filter_cutoff"]), " ".join(str(o) for o in d["filter_order"]), d["loop_radius"],
            d["n_abscissae"]))


@pytest.fixture()
def stm_files(tmp_path):
    out = []
    for n in ("skytem_hm", "skytem_lm"):
        d = json.load(open(os.path.join(ROOT, "geobipy_b200", "data", n + ".json")))
        p = tmp_path / (n + ".stm")
        _write_stm(p, d)
        out.append((str(p), d))
    return out


def test_read_stm_round_trip(stm_files, built_lib):
    from geobipy_b200 import ops, tdem
    for path, d in stm_files:
        r = tdem.read_stm(path)
        for k in ("base_frequency", "digitising_frequency", "loop_radius", "n_abscissae", "filter_order"):
            assert r[k] == d[k], k
        for k in ("waveform_time", "waveform_current", "window_start", "window_end", "filter_cutoff"):
            assert np.allclose(r[k], d[k], rtol=1e-9), k
        s = tdem.TdemSystem.read(path)
        assert s.nTimes == len(d["window_start"]) and s.components == ['z'] and s.isGA
        assert np.allclose(s.off_time, 0.5 * (np.array(d["window_start"]) + np.array(d["window_end"])))
    with pytest.raises(AssertionError):     # TdemSystem_GAAEM.py:29: the file must exist
        tdem.TdemSystem("does_not_exist.stm")
    sv = ops.make_tdem_survey_struct([tdem.read_stm(p) for p, _ in stm_files])
    ref = ops.skytem_survey_struct()
    f1, MR1, MI1, _ = ops.tdem_window_operator(sv)
    f2, MR2, MI2, _ = ops.tdem_window_operator(ref)
    assert np.allclose(MR1, MR2, rtol=1e-6, atol=1e-9 * np.abs(MR2).max()) and np.allclose(f1, f2)


def test_datapoint_std_and_layout(stm_files, built_lib):
    from geobipy_b200 import tdem
    files = [p for p, _ in stm_files]
    dp = tdem.TdemDataPoint(z=30.0, system=files, secondary_field=np.full(45, 1e-12), relative_error=[0.05, 0.03],
                            additive_error=[2e-14, 2e-13])
    assert dp.nSystems == 2 and list(dp.nTimes) == [26, 19] and dp.nChannels == 45
    # TdemDataPoint.std :329-379
    t0, t1 = dp.off_time(0), dp.off_time(1)
    exp0 = np.sqrt((0.05 * 1e-12) ** 2 + (2e-14 * np.sqrt(1e-3 / t0)) ** 2)
    exp1 = np.sqrt((0.03 * 1e-12) ** 2 + (2e-13 * np.sqrt(1e-3 / t1)) ** 2)
    assert np.allclose(dp.std, np.r_[exp0, exp1], rtol=1e-12)
    with pytest.raises(AssertionError):
        tdem.TdemDataPoint(z=30.0, system=files, relative_error=[0.05])       # one entry per system
    with pytest.raises(AssertionError):
        tdem.TdemDataPoint(z=30.0, system=files, transmitter_loop=tdem.TdemLoop(z=30.0, pitch=2.0))
    d = dp.secondary_field.copy()
    d[3] = -1.0
    d[7] = np.nan
    dp.secondary_field[:] = d
    assert dp.n_active_channels == 43


def test_read_csv_reference_layout(stm_files, tmp_path, golden_dir, built_lib):
    from geobipy_b200 import tdem
    g = np.load(os.path.join(golden_dir, "skytem_clean.npz"))
    hdr = ("Line_number,Fiducial,Easting,Northing,Height,Elevation,tx_pitch,tx_roll,tx_yaw,txrx_dx,txrx_dy,txrx_dz,rx_pitch,rx_roll,rx_yaw,"
           + ",".join("S0Z_time_%.3e" % t for t in g["times"][:26]) + "," + ",".join("S1Z_time_%.3e" % t for t in g["times"][26:]))
    rows = [hdr]
    for i in range(5):
        rows.append(",".join(repr(float(v)) for v in [0.0, i, float(i), 0.0, 30.0, 0.0, 0, 0, 0, -13.0, 0.0, 2.0, 0, 0, 0] + list(g["data"][0, i])))
    f = tmp_path / "skytem_glacial_clean.csv"
    f.write_text("\n".join(rows) + "\n")
    ds = tdem.TdemData.read_csv(str(f), [p for p, _ in stm_files])
    # (numbers are parsed with pandas' default float parser, as the reference does - within an ulp of the printed value;
    # bit-identity with the reference's own reader on its shipped file: tests/test_readers.py)
    assert ds.nPoints == 5 and ds.nChannels == 45 and np.allclose(ds.data, g["data"][0, :5], rtol=1e-15, atol=0.0)
    assert np.all(ds.height == 30.0) and np.array_equal(ds.geometry[0, 3:6], [-13.0, 0.0, 2.0])
    sv = ds.survey_struct()
    assert (sv.rx_dx, sv.rx_dy, sv.rx_dz) == (-13.0, 0.0, 2.0) and sv.n_systems == 2
    dp = ds.datapoint(2)
    assert dp.z == 30.0 and dp.receiver.x - dp.transmitter.x == -13.0 and np.allclose(dp.data, g["data"][0, 2], rtol=1e-15, atol=0.0)


@pytest.mark.gpu
def test_datapoint_forward_matches_reference_golden(stm_files, golden_dir, built_lib):
    """test_skytem of the reference (tests/test_synthetic_data.py:32-48) at discretisation tolerance."""
    from geobipy_b200 import _lib, api, tdem
    _lib.require_cuda()
    g = np.load(os.path.join(golden_dir, "skytem_clean.npz"))
    dp = tdem.TdemDataPoint(x=0.0, y=0.0, z=30.0, elevation=0.0, system=[p for p, _ in stm_files])
    for m, i in ((0, 0), (1, 40), (4, 10)):
        mod = api.Model(api.RectilinearMesh1D(edges=np.r_[0.0, g["zwedge"][i], g["zdeep"][i], np.inf]), g["sigma"][m])
        dp.forward(mod)
        ref = g["data"][m, i]
        assert np.median(np.abs(dp.predictedData / ref - 1.0)) < 3e-3
        assert np.all(np.abs(dp.predictedData[:20] / ref[:20] - 1.0) < 0.03)
        J = dp.sensitivity(mod)
        assert J.shape == (45, 3)
        dp.fm_dlogc(mod)
        assert np.array_equal(dp.sensitivity_matrix, J)
    with pytest.raises(AssertionError):   # last edge must be inf
        dp.forward(api.Model(api.RectilinearMesh1D(edges=[0.0, 5.0, 10.0]), [0.01, 0.1]))


@pytest.mark.gpu
def test_inference1d_with_skytem_datapoint(stm_files, golden_dir, built_lib):
    """The reference's driver sequence (Inference3D.infer_serial :483-492) with a dual-moment datapoint and the
    keys of skytem_options."""
    from geobipy_b200 import _lib, api, tdem
    _lib.require_cuda()
    g = np.load(os.path.join(golden_dir, "skytem_clean.npz"))
    dp = tdem.TdemDataPoint(z=30.0, system=[p for p, _ in stm_files], secondary_field=g["data"][0, 30])
    opts = dict(n_markov_chains=3000, update_plot_every=500, maximum_number_of_layers=30, minimum_depth=1.0, maximum_depth=550.0,
                minimum_thickness=None, initial_relative_error=[0.05, 0.05], minimum_relative_error=[0.005, 0.005],
                maximum_relative_error=[0.5, 0.5], initial_additive_error=[2e-14, 2e-13], minimum_additive_error=[1e-16, 1e-16],
                maximum_additive_error=[1e-10, 1e-10], relative_error_proposal_variance=[1e-6, 1e-6],
                additive_error_proposal_variance=[1e-5, 1e-5], covariance_scaling=0.5, solve_relative_error=True,
                solve_additive_error=True, interactive_plot=False, save_hdf5=True, burn_in_min_iter=500)
    inf = api.Inference1D(prng=np.random.default_rng(0), **opts)
    inf.initialize(dp)
    failed = inf.infer(None)
    assert not failed and inf.burned_in
    assert inf.model.posterior.counts.shape == (250, 1209)
    assert len(inf.relative_error_posterior) == 2 and inf.relative_error_posterior[1].counts.sum() == 3002
    assert dp.relative_error.shape == (2,) and dp.additive_error.shape == (2,)
    # clean data, 5 % assumed error: the final model fits far inside the noise
    assert inf.data_misfit < 45.0
    # posterior median conductivity of the shallow part within a factor 2.5 of the true first layer (0.01 S/m, 31 m
    # thick; a 3000-iteration chain, so only a sanity bound)
    med = inf.model.posterior.median()
    depth = 0.5 * (inf.model.posterior.y_edges[1:] + inf.model.posterior.y_edges[:-1])
    top = med[(depth > 5.0) & (depth < 15.0)]
    assert np.all((top > 0.004) & (top < 0.025)), top
    with pytest.raises(AssertionError):   # scalar error options with a dual-moment datapoint
        api.Inference1D(prng=np.random.default_rng(0), interactive_plot=False, save_hdf5=True).initialize(dp)


@pytest.mark.gpu
def test_inference3d_time_domain_survey(stm_files, tmp_path, golden_dir, built_lib):
    """The survey driver (Inference3D.infer / save) on a time-domain CSV in the reference layout."""
    from geobipy_b200 import _lib, ops, tdem
    from geobipy_b200.dataset import Inference3D
    _lib.require_cuda()
    g = np.load(os.path.join(golden_dir, "skytem_clean.npz"))
    hdr = ("Line_number,Fiducial,Easting,Northing,Height,Elevation,tx_pitch,tx_roll,tx_yaw,txrx_dx,txrx_dy,txrx_dz,rx_pitch,rx_roll,rx_yaw,"
           + ",".join("S0Z_time_%.3e" % t for t in g["times"][:26]) + "," + ",".join("S1Z_time_%.3e" % t for t in g["times"][26:]))
    rows = [hdr]
    for i in range(4):
        rows.append(",".join(repr(float(v)) for v in [100.0 + (i // 2), i, float(i), 0.0, 30.0, 0.0, 0, 0, 0, -13.0, 0.0, 2.0, 0, 0, 0] + list(g["data"][0, 10 * i])))
    f = tmp_path / "skytem_glacial_clean.csv"
    f.write_text("\n".join(rows) + "\n")
    ds = tdem.TdemData.read_csv(str(f), [p for p, _ in stm_files])
    inv = Inference3D(ds, seed=3)
    r = inv.infer(n_markov_chains=400, max_iterations=200, **{k: v for k, v in ops.SKYTEM_OPTIONS.items()})
    assert r["hitmap"].shape == (4, 250, 1209) and (r["scalars"][:, _lib.S_ITER] == 200).all()
    assert r["rel_hist"].shape == (4, 2, 99) and r["summary_p50"].shape == (4, 1209)
    one = Inference3D(ds, seed=3).infer(index=2, n_markov_chains=400, max_iterations=200, **{k: v for k, v in ops.SKYTEM_OPTIONS.items()})
    assert np.array_equal(one["hitmap"][0], r["hitmap"][2])        # (seed, sounding index) fixes the stream
    files = inv.save(str(tmp_path / "out"), format="npz")
    assert [os.path.basename(x) for x in files] == ["100.npz", "101.npz"]
    # HDF5 in the reference's layout (TdemDataPoint.createHdf :603-626: systems as .stm text, loop pair, per-system error
    # posteriors); the reference reads the .stm text back with its own parser (TdemSystem_GAAEM.fromHdf :120-129)
    from geobipy_b200 import h5lite
    files = inv.save(str(tmp_path / "out"))
    assert [os.path.basename(x) for x in files] == ["100.h5", "101.h5"]
    h = h5lite.File(files[1], "r")
    assert h["data"].attrs["repr"] == "TdemData" and h["data/nSystems"][()] == 2 and h["nsystems"][()] == 2
    assert np.array_equal(h["model/values/posterior/values/data"][()], r["hitmap"][2:])
    assert np.array_equal(h["data/relative_error/posterior1/values/data"][()], r["rel_hist"][2:, 1])
    assert h["data/loop_pair/receiver/x/data"][0] == ds.x[2] - 13.0 and h["data/loop_pair/transmitter/z/data"][1] == 30.0
    lines = list(h["data/System0"].attrs["data"])
    p_stm = tmp_path / "back.stm"
    p_stm.write_text("".join(lines))
    assert tdem.read_stm(str(p_stm)) == ds.system[0].definition
    dp = ds.datapoint(3)
    kb = int(r["scalars"][3, _lib.S_BEST_K])
    from geobipy_b200 import api
    dp.forward(api.Model(api.RectilinearMesh1D(edges=r["best_edges"][3, :kb + 1]), r["best_sigma"][3, :kb]))
    assert np.allclose(h["data/predicted_secondary_field/data"][1], dp.predictedData, rtol=1e-10)


@pytest.mark.gpu
def test_inference1d_with_tempest_datapoint(golden_dir, built_lib, tmp_path):
    """The reference's calling sequence with a Tempest_datapoint and the keys of tempest_options: `initial_additive_error` is
    the additive level of every channel, the sampled errors are one relative error and one multiplier per component."""
    from geobipy_b200 import _lib, api, ops, tdem
    _lib.require_cuda()
    g = np.load(os.path.join(golden_dir, "ref_tempest_chain_1.npz"))
    system = tdem.TdemSystem(definition=ops.tempest_definition())
    dp = tdem.Tempest_datapoint(x=0.0, y=0.0, z=120.0, elevation=0.0, system=[system], transmitter_loop=tdem.TdemLoop(z=120.0),
                                receiver_loop=tdem.TdemLoop(x=-107.0, z=75.0), secondary_field=g["secondary"], primary_field=g["primary"])
    assert np.allclose(dp.data, g["data"], rtol=1e-13)
    options = dict(n_markov_chains=600, solve_relative_error=True, initial_relative_error=[0.001, 0.001], minimum_relative_error=[0.0001, 0.0001],
                   maximum_relative_error=[0.01, 0.01], relative_error_proposal_variance=[1e-6, 1e-6], solve_additive_error=True,
                   initial_additive_error=np.asarray(ops.TEMPEST_ADDITIVE), additive_error_proposal_variance=1e-6,
                   minimum_additive_error=[0.001, 0.001], maximum_additive_error=[100.0, 100.0], solve_transmitter_z=False,
                   solve_receiver_pitch=False, maximum_number_of_layers=30, minimum_depth=1.0, maximum_depth=550.0, minimum_thickness=None,
                   probability_of_birth=1 / 6, probability_of_death=1 / 6, probability_of_perturb=1 / 6, probability_of_no_change=0.5,
                   gradient_standard_deviation=5, covariance_scaling=0.5, interactive_plot=False, save_hdf5=True)
    inf = api.Inference1D(seed=5, precision=64, **options)
    inf.initialize(dp)
    assert inf.options.gradient_std == 5.0 and inf.options.add_init == 1.0 and inf.options.n_systems == 2
    inf.infer(None)
    assert inf.iteration == 600 and abs(inf.halfspace / float(g["halfspace"]) - 1.0) < 1e-9
    assert len(inf.relative_error_posterior) == 2 and inf.relative_error_posterior[1].counts.sum() == 600
    assert np.allclose(dp.additive_error, ops.TEMPEST_ADDITIVE) and np.all(np.abs(np.log(dp.additive_error_multiplier)) < 6e-3)
    assert dp.additive_error_multiplier.posterior[0].counts.sum() == 600
    # the same chain through the operator level
    sv = dp.survey_struct()
    r = ops.rjmcmc_run(sv, inf.options, dp.data[None], np.asarray([120.0]), seed=5, precision=64)
    assert np.array_equal(r["hitmap"][0], inf.hitmap.counts)
    # predicted data of the final model: secondary + primary
    chk = tdem.Tempest_datapoint(x=0.0, y=0.0, z=120.0, elevation=0.0, system=[system], transmitter_loop=tdem.TdemLoop(z=120.0),
                                 receiver_loop=tdem.TdemLoop(x=-107.0, z=75.0))
    chk.forward(inf.model)
    assert np.allclose(chk.predictedData, dp.predictedData) and inf.data_misfit < 18415.0
    # the result file of a Tempest line (Tempest_datapoint.createHdf / writeHdf :566-586: errors per component, the additive
    # level of every channel, the sampled multiplier with its posteriors, data = secondary + primary)
    from geobipy_b200 import h5lite
    dp.fiducial = 20.0
    with h5lite.File(str(tmp_path / "tempest.h5"), "w") as f:
        inf.createHdf(f, add_axis=np.asarray([10.0, 20.0, 30.0]))
        f["data/fiducial/data"][:] = [10.0, 20.0, 30.0]
        inf.writeHdf(f)
    h = h5lite.File(str(tmp_path / "tempest.h5"), "r")
    assert h["data"].attrs["repr"] == "TempestData" and h["data/components"][()].tolist() == [0, 2]
    assert np.allclose(h["data/additive_error/data"][1], ops.TEMPEST_ADDITIVE) and np.isnan(h["data/additive_error/data"][0]).all()
    assert h["data/additive_error_multiplier/posterior1/values/data"][1].sum() == 600
    assert np.allclose(h["data/additive_error_multiplier/data"][1], inf.best_additive_error)
    assert np.allclose(h["data/data/data"][1], g["data"], rtol=1e-13) and np.allclose(h["data/secondary_field/data"][1], g["secondary"])
    assert np.allclose(h["data/predicted_data/data"][1], inf.best_datapoint.predictedData, rtol=1e-12)
    assert np.allclose(h["data/primary_field/data"][1], g["primary"]) and np.allclose(h["data/predicted_primary_field/data"][1], g["primary"], rtol=1e-9)
    assert h["reciprocate_parameter"][()] == False and h["nsystems"][()] == 1


def _tempest_csv(path, golden_dir, n=4):
    """A Tempest survey file in the reference's layout (tests/data_checks/tempest_*_clean.csv: PX / PZ primary-field columns,
    S0X_time_* / S0Z_time_* secondary-field columns) from the reference's known-answer vectors."""
    g = np.load(os.path.join(golden_dir, "tempest_clean.npz"))
    hdr = ("Line_number,Fiducial,Easting,Northing,Height,Elevation,tx_pitch,tx_roll,tx_yaw,txrx_dx,txrx_dy,txrx_dz,rx_pitch,rx_roll,rx_yaw,PX,PZ,"
           + ",".join("S0X_time_%.3e" % t for t in g["times"]) + "," + ",".join("S0Z_time_%.3e" % t for t in g["times"]))
    rows = [hdr]
    for i in range(n):
        rows.append(",".join(repr(float(v)) for v in [100.0 + (i // 2), i, float(i), 0.0, 120.0, 0.0, 0, 0, 0, -107.0, 0.0, -45.0, 0, 0, 0]
                             + list(g["primary"][0, 10 * i]) + list(g["data"][0, 10 * i])))
    with open(path, "w") as f:
        f.write("\n".join(rows) + "\n")
    return g


def test_tempest_data_reader(tmp_path, golden_dir):
    """TempestData.read_csv (classes/data/dataset/TempestData.py:140-273): the x_time / z_time columns are the secondary field, PX /
    PZ the primary field of each component; `data` = secondary + primary (Tempest_datapoint.data :107-115)."""
    from geobipy_b200 import ops, tdem
    g = _tempest_csv(str(tmp_path / "tempest.csv"), golden_dir)
    ds = tdem.TempestData.read_csv(str(tmp_path / "tempest.csv"), [tdem.TdemSystem(definition=ops.tempest_definition())])
    assert ds.nPoints == 4 and ds.nChannels == 30 and ds.n_components == 2
    assert np.array_equal(ds.secondary_field[1], g["data"][0, 10]) and np.array_equal(ds.primary_field[1], g["primary"][0, 10])
    assert np.allclose(ds.data[1, :15], g["data"][0, 10, :15] + g["primary"][0, 10, 0], rtol=0, atol=0)
    assert np.allclose(ds.data[1, 15:], g["data"][0, 10, 15:] + g["primary"][0, 10, 1], rtol=0, atol=0)
    dp = ds.datapoint(1)
    assert isinstance(dp, tdem.Tempest_datapoint) and np.array_equal(dp.data, ds.data[1])
    sub = ds.subset(np.asarray([2, 3]))
    assert sub.nPoints == 2 and np.array_equal(sub.primary_field, ds.primary_field[2:]) and np.array_equal(sub.lineNumber, [101.0, 101.0])
    if os.path.isdir("/root/reference/tests/data_checks"):     # the reference's own file (build container)
        ref = tdem.TempestData.read_csv("/root/reference/tests/data_checks/tempest_glacial_clean.csv", [tdem.TdemSystem(definition=ops.tempest_definition())])
        assert ref.nPoints == 79 and np.array_equal(ref.secondary_field, g["data"][0]) and np.array_equal(ref.primary_field, g["primary"][0])


@pytest.mark.gpu
def test_inference3d_tempest_survey(tmp_path, golden_dir, built_lib):
    """The survey driver on a Tempest CSV with the keys of tempest_options: per-channel additive levels, errors per component;
    one `<line>.h5` per flight line in the layout of Tempest_datapoint.createHdf (:566-586)."""
    from geobipy_b200 import _lib, api, h5lite, ops, tdem
    from geobipy_b200.dataset import Inference3D
    _lib.require_cuda()
    _tempest_csv(str(tmp_path / "tempest.csv"), golden_dir)
    ds = tdem.TempestData.read_csv(str(tmp_path / "tempest.csv"), [tdem.TdemSystem(definition=ops.tempest_definition())])
    options = dict(n_markov_chains=400, solve_relative_error=True, initial_relative_error=[0.001, 0.001], minimum_relative_error=[0.0001, 0.0001],
                   maximum_relative_error=[0.01, 0.01], relative_error_proposal_variance=[1e-6, 1e-6], solve_additive_error=True,
                   initial_additive_error=np.asarray(ops.TEMPEST_ADDITIVE), additive_error_proposal_variance=1e-6,
                   minimum_additive_error=[0.001, 0.001], maximum_additive_error=[100.0, 100.0], maximum_number_of_layers=30,
                   minimum_depth=1.0, maximum_depth=550.0, minimum_thickness=None, probability_of_birth=1 / 6, probability_of_death=1 / 6,
                   probability_of_perturb=1 / 6, probability_of_no_change=0.5, gradient_standard_deviation=5, covariance_scaling=0.5)
    inv = Inference3D(ds, seed=3)
    r = inv.infer(max_iterations=200, precision=64, **options)
    assert (r["scalars"][:, _lib.S_ITER] == 200).all() and r["rel_hist"].shape == (4, 2, 99)
    # the same chains through the one-sounding interface (Inference1D with a Tempest_datapoint)
    inf = api.Inference1D(seed=3, precision=64, sounding_index=2, interactive_plot=False, save_hdf5=True, **options)
    inf.initialize(ds.datapoint(2))
    inf.infer(None, max_iterations=200)
    assert np.array_equal(inf.hitmap.counts, r["hitmap"][2])
    files = inv.save(str(tmp_path / "out"))
    assert [os.path.basename(x) for x in files] == ["100.h5", "101.h5"]
    h = h5lite.File(files[1], "r")
    assert h["data"].attrs["repr"] == "TempestData" and h["data/nSystems"][()] == 1 and h["data/components"][()].tolist() == [0, 2]
    assert np.array_equal(h["model/values/posterior/values/data"][()], r["hitmap"][2:])
    assert np.array_equal(h["data/additive_error_multiplier/posterior1/values/data"][()], r["add_hist"][2:, 1])
    assert np.allclose(h["data/additive_error/data"][()], np.tile(ops.TEMPEST_ADDITIVE, (2, 1)))
    assert np.array_equal(h["data/secondary_field/data"][()], ds.secondary_field[2:]) and np.array_equal(h["data/primary_field/data"][()], ds.primary_field[2:])
    assert np.allclose(h["data/data/data"][()], ds.data[2:], rtol=1e-15)
    # predicted data of the best model of sounding 3 = predicted secondary + primary field of the geometry
    dp = ds.datapoint(3)
    kb = int(r["scalars"][3, _lib.S_BEST_K])
    dp.forward(api.Model(api.RectilinearMesh1D(edges=r["best_edges"][3, :kb + 1]), r["best_sigma"][3, :kb]))
    assert np.allclose(h["data/predicted_data/data"][1], dp.predictedData, rtol=1e-10)
    assert np.allclose(h["data/predicted_secondary_field/data"][1], dp.predicted_secondary_field, rtol=1e-10)
    best_rel = r["scalars"][3][[_lib.S_BEST_REL, _lib.S_BEST_REL2]]
    best_mul = r["scalars"][3][[_lib.S_BEST_ADD, _lib.S_BEST_ADD2]]
    std = np.sqrt((np.repeat(best_rel, 15) * ds.data[3]) ** 2 + (np.repeat(best_mul, 15) * np.asarray(ops.TEMPEST_ADDITIVE)) ** 2)
    assert np.allclose(h["data/std/data"][1], std, rtol=1e-12)


@pytest.mark.gpu
def test_inference3d_time_domain_survey_with_sampled_transmitter_height(stm_files, tmp_path, golden_dir, built_lib):
    """solve_transmitter_z through the survey driver: the transmitter loop's z of the line file is a StatArray with its posterior
    (EmLoop.createHdf -> Point.createHdf :1403-1427), the stored receiver loop keeps the z it was given (golden:
    tests/golden/hdf_layout_tdem_height.npz, recorded from the reference's own writer)."""
    from geobipy_b200 import _lib, h5lite, ops, tdem
    from geobipy_b200.dataset import Inference3D
    _lib.require_cuda()
    g = np.load(os.path.join(golden_dir, "skytem_clean.npz"))
    hdr = ("Line_number,Fiducial,Easting,Northing,Height,Elevation,tx_pitch,tx_roll,tx_yaw,txrx_dx,txrx_dy,txrx_dz,rx_pitch,rx_roll,rx_yaw,"
           + ",".join("S0Z_time_%.3e" % t for t in g["times"][:26]) + "," + ",".join("S1Z_time_%.3e" % t for t in g["times"][26:]))
    rows = [hdr]
    for i in range(3):
        rows.append(",".join(repr(float(v)) for v in [100.0, i, float(i), 0.0, 30.0, 0.0, 0, 0, 0, -13.0, 0.0, 2.0, 0, 0, 0] + list(g["data"][0, 10 * i])))
    f = tmp_path / "skytem.csv"
    f.write_text("\n".join(rows) + "\n")
    ds = tdem.TdemData.read_csv(str(f), [p for p, _ in stm_files])
    inv = Inference3D(ds, seed=5)
    r = inv.infer(n_markov_chains=400, max_iterations=300, solve_transmitter_z=True, maximum_transmitter_z_change=1.0,
                  transmitter_z_proposal_variance=0.01, **{k: v for k, v in ops.SKYTEM_OPTIONS.items()})
    assert r["height_hist"].shape == (3, 99) and (r["height_hist"].sum(axis=1) == 300).all()
    files = inv.save(str(tmp_path / "out"))
    h = h5lite.File(files[0], "r")
    z = h["data/loop_pair/transmitter/z"]
    assert z.attrs["repr"] == "StatArray" and z["n_posteriors"][()] == 1
    assert np.array_equal(z["posterior/values/data"][()], r["height_hist"])
    assert np.allclose(z["posterior/mesh/y/relative_to/data"][()], r["scalars"][:, _lib.S_HEIGHT_REF])
    assert np.allclose(z["posterior/mesh/y/edges/data"][()], np.linspace(-1.0, 1.0, 100))
    assert np.allclose(z["data"][()], r["scalars"][:, _lib.S_BEST_HEIGHT]) and np.all(np.abs(z["data"][()] - 30.0) <= 1.0)
    assert np.allclose(h["data/loop_pair/receiver/z/data"][()], 32.0) and np.allclose(h["data/z/data"][()], 30.0)
