"""Generate the committed golden vectors under tests/golden/ (BUILD CONTAINER ONLY).

Needs the read-only reference tree at /root/reference; imports the *unmodified* reference
through oracle/ref_shims.py and records its outputs.  The GPU box never runs this.

    python tests/golden/make_golden.py fdem         # resolve_clean.npz, fdem_random_models.npz
    python tests/golden/make_golden.py fdem_tensor  # fdem_tensor_models.npz (tensor ids 3 / 7, vertical coil offsets)
    python tests/golden/make_golden.py tdem         # skytem_clean.npz
    python tests/golden/make_golden.py tempest_transitions   # tempest_transitions.npz (reference sampler with a Tempest_datapoint)
    python tests/golden/make_golden.py tempest      # tempest_clean.npz (X and Z components, B field, point dipole)
    python tests/golden/make_golden.py tdem_transitions   # tdem_transitions.npz (reference sampler + fake_gatdaem1d)
    python tests/golden/make_golden.py tdem_chain <i> [rep]   # ref_tdem_chain_<i>[_r<rep>].npz (minutes each)
    python tests/golden/make_golden.py bins         # posterior_bins.npz
    python tests/golden/make_golden.py transitions  # transitions.npz
    python tests/golden/make_golden.py chain <i> [rep]   # ref_chain_<i>[_r<rep>].npz (minutes each)
    python tests/golden/make_golden.py readers      # reader_files/ (heads of the reference's shipped data files) + readers.npz
    python tests/golden/make_golden.py hdf fdem|fdem_height|tdem|tdem_height|tempest   # hdf_layout_<kind>.npz: the reference's own HDF5 result tree

Files written
  resolve_clean.npz       the reference's own known-answer vectors
                          (tests/data_checks/resolve_*_clean.csv, 6 models x 79 soundings x 12 channels;
                          tests/test_synthetic_data.py:16-30) plus the model definitions
                          (Model.create_synthetic_model, classes/model/Model.py:885-920).
  skytem_clean.npz        the reference's only known-answer vectors for the time-domain path
                          (tests/data_checks/skytem_*_clean.csv, 6 models x 79 soundings x 45 windows,
                          tests/test_synthetic_data.py:32-48) plus model definitions and geometry columns.
  fdem_random_models.npz  nbFdem1dfwd / nbFdem1dsen outputs (fdem1d_numba.py:25,72) for random models.
  posterior_bins.npz      bin indices the reference's Histogram meshes assign to probe values.
  transitions.npz         per-term records of Inference1D.accept_reject (Inference1D.py:537-631).
  ref_chain_<i>.npz       posterior arrays of full reference chains (Inference1D.infer loop).
"""
import os
import sys
import time
import warnings
from copy import deepcopy

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import ref_shims  # noqa: E402

REF = ref_shims.REFERENCE_ROOT
SUP = os.path.join(REF, "documentation_source/source/supplementary")

MODELS = ["glacial", "saline_clay", "resistive_dolomites", "resistive_basement", "coastal_salt_water",
          "ice_over_salt_water"]


def _geobipy():
    return ref_shims.import_reference()


def _options(n_markov_chains):
    from geobipy import user_parameters
    kw = user_parameters.read(os.path.join(SUP, "options_files/resolve_options"))
    kw["interactive_plot"] = False
    kw["save_hdf5"] = True
    kw["n_markov_chains"] = n_markov_chains
    for k in ("data_type", "data_filename", "system_filename", "data_directory"):
        kw.pop(k, None)
    kw.pop("seed")
    return dict(kw)


def _system():
    from geobipy import FdemSystem
    return FdemSystem.read(os.path.join(SUP, "data/resolve.stm"))


def _datapoint(system, data, z):
    from geobipy import FdemDataPoint
    return FdemDataPoint(x=0.0, y=0.0, z=z, elevation=0.0, data=data, std=None, predictedData=None,
                         system=system, lineNumber=0.0, fiducial=0.0)


def _model(edges, sigma):
    from geobipy import Model, RectilinearMesh1D, StatArray
    mesh = RectilinearMesh1D(edges=np.asarray(edges, dtype=np.float64))
    return Model(mesh=mesh, values=StatArray(np.asarray(sigma, dtype=np.float64), "Conductivity", "S/m"))


from geobipy_b200.synthetic import synthetic_sounding  # noqa: E402


def make_fdem():
    import pandas as pd
    _geobipy()
    nb = ref_shims.load_numba_kernels()
    from geobipy import Model
    data = np.zeros((6, 79, 12))
    sig = np.zeros((6, 3))
    for m, name in enumerate(MODELS):
        df = pd.read_csv(os.path.join(REF, "tests/data_checks/resolve_%s_clean.csv" % name))
        data[m] = df.values[:, 6:18]
        assert np.all(df["Height"].values == 30.0)
        mod = Model.create_synthetic_model(name)
        sig[m] = np.asarray(mod.values[0, :])
    zwedge = np.linspace(50.0, 1.0, 79) / 10.0
    zdeep = np.linspace(75.0, 500.0, 79) / 10.0
    np.savez_compressed(os.path.join(HERE, "resolve_clean.npz"), data=data, sigma=sig, zwedge=zwedge, zdeep=zdeep,
                        height=30.0, models=np.array(MODELS))

    s = _system()
    tid = np.asarray(s.tensor_id, dtype=np.int32)
    freq = np.asarray(s.frequencies)
    sep = np.asarray(s.loop_separation)
    tmom = np.asarray(s.transmitter.moment, dtype=np.float64)
    rmom = np.asarray(s.receiver.moment, dtype=np.float64)
    xs = np.asarray(s.loop_offsets[0, :])

    def run(f, sigma, thk, z):
        tH = z + np.zeros(6)
        rH = -tH
        return f(tid, freq, tH, rH, tmom, xs, sep, s.w0, s.lamda0, s.lamda02, s.w1, s.lamda1, s.lamda12,
                 tmom * rmom, sigma, np.zeros_like(sigma), np.zeros_like(sigma), thk)

    rng = np.random.default_rng(20261017)
    n = 256
    Ls = np.r_[np.arange(1, 31), rng.integers(1, 31, n - 30)]
    sigma = np.full((n, 30), np.nan)
    thk = np.full((n, 30), np.nan)
    height = rng.uniform(25.0, 45.0, n)
    fwd = np.zeros((n, 12))
    sens = np.full((n, 12, 30), np.nan)
    for i in range(n):
        L = int(Ls[i])
        sg = 10.0 ** rng.uniform(-4.0, 1.0, L)
        t = np.r_[np.exp(rng.uniform(np.log(1.0), np.log(60.0), L - 1)), np.inf]
        sigma[i, :L] = sg
        thk[i, :L] = t
        r = run(nb.nbFdem1dfwd, sg, t, height[i])
        fwd[i] = np.r_[r.real, r.imag]
        J = run(nb.nbFdem1dsen, sg, t, height[i])
        sens[i, :, :L] = np.vstack([J.real, J.imag])
    np.savez_compressed(os.path.join(HERE, "fdem_random_models.npz"), nlayers=Ls.astype(np.int32), sigma=sigma,
                        thickness=thk, height=height, forward=fwd, sensitivity=sens)
    print("fdem goldens written")


# A frequency-domain system that exercises what RESOLVE does not: the mixed tensor components Hxz / Hzx
# (tensor ids 3 and 7, fdem1d_numba.py:359-408) and coils offset vertically from the observation point
# (fdem1d.py:31-32: transmitter_height = altitude + tz, receiver_height = -transmitter_height + rz).  rz >= 0 only: the
# reference's primary-field sums use exp(-lambda * rz) (hSum = rHeight + tHeight = rz) and overflow to inf / NaN for a
# receiver offset below the observation point.
TENSOR_STM = """freq, tor, tmom, tx, ty, tzoff, ror, rmom, rx, ry, rzoff
400, z, 1, 0, 0, 0.5, x, 1, 7.9, 0, 0.3
1800, x, 1.5, 0, 0, -0.4, z, 1, 8.1, 0, 0.6
3300, x, -1, 0, 0, 0.25, x, 2, 9.0, 0, 0.25
8200, z, 1, 0.2, 0, 0, z, 1, 7.9, 0, 1.0
40000, z, 2, 0, 0, -0.5, x, -1, 6.5, 0, 0.5
130000, x, 1, 0, 0, 0.3, z, 1, 7.5, 0, 0.2
"""


def make_fdem_tensor():
    """fdem_tensor_models.npz: fdem1dfwd / fdem1dsen of the live reference (fdem1d.py:10, :87 -> the Numba kernels)
    through its own FdemSystem class for the system above, 96 random models."""
    import tempfile
    _geobipy()
    from geobipy import FdemSystem
    from geobipy.src.classes.forwardmodelling.Electromagnetic.FD.fdem1d import fdem1dfwd, fdem1dsen
    with tempfile.NamedTemporaryFile("w", suffix=".stm", delete=False) as f:
        f.write(TENSOR_STM)
    s = FdemSystem.read(f.name)
    assert list(s.tensor_id) == [3, 7, 1, 9, 3, 7], list(s.tensor_id)
    rng = np.random.default_rng(20261018)
    n = 96
    Ls = np.r_[np.arange(1, 31), rng.integers(1, 31, n - 30)]
    sigma = np.full((n, 30), np.nan)
    thk = np.full((n, 30), np.nan)
    height = rng.uniform(25.0, 45.0, n)
    fwd = np.zeros((n, 12))
    sens = np.full((n, 12, 30), np.nan)
    for i in range(n):
        L = int(Ls[i])
        sg = 10.0 ** rng.uniform(-4.0, 1.0, L)
        t = np.r_[np.exp(rng.uniform(np.log(1.0), np.log(60.0), L - 1)), np.inf]
        sigma[i, :L] = sg
        thk[i, :L] = t
        mod = _model(np.r_[0.0, np.cumsum(t)], sg)
        r = fdem1dfwd(s, mod, height[i])
        fwd[i] = np.r_[r.real, r.imag]
        J = fdem1dsen(s, mod, height[i])
        sens[i, :, :L] = np.vstack([J.real, J.imag])
    cols = {k: np.asarray(v) for k, v in dict(
        freq=s.frequencies, tor=np.asarray(s.transmitter.orientation), tmom=s.transmitter.moment, tx=s.transmitter.x,
        ty=s.transmitter.y, tz=s.transmitter.z, ror=np.asarray(s.receiver.orientation), rmom=s.receiver.moment,
        rx=s.receiver.x, ry=s.receiver.y, rz=s.receiver.z).items()}
    np.savez_compressed(os.path.join(HERE, "fdem_tensor_models.npz"), stm=TENSOR_STM, tensor_id=np.asarray(s.tensor_id),
                        nlayers=Ls.astype(np.int32), sigma=sigma, thickness=thk, height=height, forward=fwd,
                        sensitivity=sens, **{"sys_" + k: v for k, v in cols.items()})
    print("fdem tensor goldens written; |d| range", np.abs(fwd).min(), np.abs(fwd).max())


def make_tdem():
    """The SkyTEM known-answer CSVs (plain CSV: no reference import needed)."""
    import pandas as pd
    data = np.zeros((6, 79, 45))
    sig = np.array([[1e-2, 1e-1, 0.03333333], [1e-2, 1e-1, 1.0], [2e-2, 2e-3, 2e-2], [1e-2, 1e-1, 1e-4],
                    [1.0, 1e-2, 5e-2], [1e-4, 1e-2, 1.0]])  # Model.create_synthetic_model, Model.py:902-908
    geom = None
    for m, name in enumerate(MODELS):
        df = pd.read_csv(os.path.join(REF, "tests/data_checks/skytem_%s_clean.csv" % name))
        data[m] = df.values[:, 15:60]
        g = df[["Height", "tx_pitch", "tx_roll", "tx_yaw", "txrx_dx", "txrx_dy", "txrx_dz", "rx_pitch", "rx_roll",
                "rx_yaw"]].values
        assert np.all(g == g[0])
        geom = g[0]
        times = np.array([float(c.split("_")[-1]) for c in df.columns[15:60]])
    np.savez_compressed(os.path.join(HERE, "skytem_clean.npz"), data=data, sigma=sig,
                        zwedge=np.linspace(50.0, 1.0, 79), zdeep=np.linspace(75.0, 500.0, 79), geometry=geom,
                        times=times, models=np.array(MODELS))


# solve_z of the height tests: the sensor height handed to the inversion is 0.4 m above the one the data were
# simulated at, the prior is +-1 m around it, the random walk has a 0.1 m standard deviation.
HEIGHT_KW = dict(solve_z=True, maximum_z_change=1.0, z_proposal_variance=0.01)
HEIGHT_BIAS = 0.4


def _initialised_inference(data, z, n_markov_chains, seed, **extra):
    from geobipy import Inference1D, get_prng
    kw = _options(n_markov_chains)
    kw.update(extra)
    kw["prng"] = get_prng(seed=seed)
    inf = Inference1D(**kw)
    dp = _datapoint(_system(), np.asarray(data, dtype=np.float64), z)
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        inf.initialize(dp)
    return inf


def _observed(i):
    edges, sigma, height, noise = synthetic_sounding(i)
    dp = _datapoint(_system(), None, height)
    dp.forward(_model(edges, sigma))
    clean = np.asarray(dp.predictedData).copy()
    std = np.sqrt((0.05 * clean) ** 2 + 5.0 ** 2)
    return clean + noise * std, height, edges, sigma


def make_bins():
    _geobipy()
    data, z, _, _ = _observed(0)
    inf = _initialised_inference(data, z, 1000, 7)
    rng = np.random.default_rng(3)
    hs = float(inf.halfspace.item())
    h = inf.model.values.posterior
    v = hs * np.exp(rng.uniform(-11.0, 11.0, 4000))
    sig_idx = np.asarray(h.cellIndex(v, axis=0, clip=True))
    d = rng.uniform(0.0, 219.9, 4000)
    dep_idx = np.asarray(inf.model.mesh.edges.posterior.mesh.cellIndices(d, clip=True))
    re = np.exp(rng.uniform(np.log(0.001), np.log(0.5), 2000))
    ae = np.exp(rng.uniform(np.log(3.0), np.log(20.0), 2000))
    rel_idx = np.asarray(inf.datapoint.relative_error.posterior.mesh.cellIndices(re, clip=True))
    add_idx = np.asarray(inf.datapoint.additive_error.posterior.mesh.cellIndices(ae, clip=True))
    centres = np.asarray(h.axis(1).centres)
    np.savez_compressed(os.path.join(HERE, "posterior_bins.npz"), halfspace=hs, sigma_probe=v, sigma_idx=sig_idx,
                        depth_probe=d, depth_idx=dep_idx, rel_probe=re, rel_idx=rel_idx, add_probe=ae,
                        add_idx=add_idx, hitmap_shape=np.asarray(h.counts.shape), depth_centres=centres,
                        ncells_bins=int(inf.model.mesh.nCells.posterior.counts.size),
                        err_bins=int(inf.datapoint.relative_error.posterior.counts.size))
    print("bins golden written", h.counts.shape)


ACT = {"insert": 0, "delete": 1, "perturb": 2, "none": 3}


def make_transitions(n_soundings=6, n_iter=250, height=False):
    _geobipy()
    from numpy import inf as npinf
    recs = []
    for sidx in range(n_soundings):
        data, z, _, _ = _observed(sidx)
        if height:
            z = z + HEIGHT_BIAS
            inf = _initialised_inference(data, z, 100000, 3000 + sidx, **HEIGHT_KW)
        else:
            inf = _initialised_inference(data, z, 100000, 1000 + sidx)
        hs = float(inf.halfspace.item())
        init = dict(prior=float(inf.prior), likelihood=float(inf.likelihood), misfit=float(inf.data_misfit))
        for it in range(n_iter):
            dp0, m0 = inf.datapoint, inf.model
            J_in = np.asarray(dp0.sensitivity_matrix).copy()
            pred_in = np.asarray(dp0.predictedData).copy()
            rel_cur = float(np.asarray(dp0.relative_error).item())
            add_cur = float(np.asarray(dp0.additive_error).item())
            tdp = deepcopy(dp0)
            z_cur = float(np.asarray(dp0.z).item())
            remapped, test = m0.perturb(tdp, -npinf, npinf, alpha=inf.covariance_scaling)
            action = ACT[remapped.mesh.action[0]]
            k = int(remapped.nCells.item())
            grad = np.asarray(remapped.local_gradient(observation=tdp)).copy()
            H = np.asarray(test.values.proposal.variance).copy()
            mean = np.asarray(test.values.proposal.mean).copy()
            tdp.perturb()
            tdp.forward(test)
            misfit = float(tdp.data_misfit())
            prior = float(tdp.probability) + float(test.probability(inf.solve_parameter, inf.solve_gradient))
            like = float(tdp.likelihood(log=True))
            prop, prop1 = test.proposal_probabilities(remapped, tdp, alpha=inf.covariance_scaling)
            rec = dict(sounding=sidx, altitude=z_cur, altitude_ref=z, altitude_test=float(np.asarray(tdp.z).item()),
                       sigma_ref=hs, data=np.asarray(data), k=k, action=action,
                       edges=np.asarray(remapped.mesh.edges).copy(), sigma_remap=np.asarray(remapped.values).copy(),
                       sigma_test=np.asarray(test.values).copy(), rel_cur=rel_cur, add_cur=add_cur,
                       rel_test=float(np.asarray(tdp.relative_error).item()),
                       add_test=float(np.asarray(tdp.additive_error).item()), J_in=J_in, pred_in=pred_in,
                       H=H, gradient=grad, newton_mean=mean, pred_test=np.asarray(tdp.predictedData).copy(),
                       misfit_test=misfit, prior_test=prior, likelihood_test=like, proposal=float(prop),
                       proposal1=float(prop1), init_prior=init["prior"], init_likelihood=init["likelihood"],
                       init_misfit=init["misfit"])
            recs.append(rec)
            # accept / reject exactly as Inference1D.accept_reject :604-623 does
            log_alpha = (prior - inf.prior) + (like - inf.likelihood) + (prop - prop1)
            if np.exp(log_alpha) > inf.prng.uniform():
                inf.data_misfit, inf.prior, inf.likelihood = misfit, prior, like
                inf.model, inf.datapoint = test, tdp
        print("sounding", sidx, "k now", inf.model.nCells.item(), flush=True)
    n = len(recs)
    out = {}
    for key in recs[0]:
        vals = [r[key] for r in recs]
        if np.ndim(vals[0]) == 0:
            out[key] = np.asarray(vals)
        else:
            obj = np.empty(n, dtype=object)
            for i, v in enumerate(vals):
                obj[i] = np.asarray(v)
            out[key] = obj
    np.savez_compressed(os.path.join(HERE, "transitions_height.npz" if height else "transitions.npz"), **out)
    print("transitions written:", n, "actions", np.bincount(out["action"], minlength=4))


def _tdem_setup():
    import fake_gatdaem1d
    fake_gatdaem1d.install()
    _geobipy()
    from geobipy import user_parameters
    kw = dict(user_parameters.read(os.path.join(SUP, "options_files/skytem_options")))
    kw["interactive_plot"] = False
    kw["save_hdf5"] = True
    for k in ("data_type", "data_filename", "system_filename", "data_directory", "seed"):
        kw.pop(k, None)
    return kw


def _tdem_datapoint(data, z):
    from geobipy import TdemDataPoint, CircularLoop
    tx = CircularLoop(x=0.0, y=0.0, z=z, pitch=0.0, roll=0.0, yaw=0.0, radius=10.416)
    rx = CircularLoop(x=-13.0, y=0.0, z=z + 2.0, pitch=0.0, roll=0.0, yaw=0.0, radius=10.416)
    return TdemDataPoint(x=0.0, y=0.0, z=z, elevation=0.0, secondary_field=data,
                         system=[os.path.join(SUP, "data/SkytemHM.stm"), os.path.join(SUP, "data/SkytemLM.stm")],
                         transmitter_loop=tx, receiver_loop=rx)


# solve_transmitter_z of the time-domain height tests (tempest_options keys :104-108; same prior / proposal as HEIGHT_KW)
TX_HEIGHT_KW = dict(solve_transmitter_z=True, maximum_transmitter_z_change=1.0, transmitter_z_proposal_variance=0.01)


def make_tdem_transitions(n_soundings=5, n_iter=200, height=False):
    """Per-term records of the reference's Inference1D.accept_reject with a dual-moment TdemDataPoint; the external
    gatdaem1d is replaced by tests/golden/fake_gatdaem1d.py (oracle forward), so these pin the sampler terms
    AROUND the forward: time-domain std model, per-system error priors / proposals, 45-channel Hessian."""
    kw0 = _tdem_setup()
    from geobipy import Inference1D, get_prng
    from geobipy_b200.synthetic import skytem_noise_std
    import oracle_py as O
    from numpy import inf as npinf
    import io
    import contextlib
    tsys = O.make_tdem_system()
    tc = np.array(tsys.t_centre[:45])
    recs = []
    for sidx in range(n_soundings):
        edges, sigma, z, noise = synthetic_sounding(sidx, 400.0, 45)
        clean = O.tdem_forward(tsys, z, sigma, np.r_[np.diff(edges)[:-1], 1.0])
        data = clean + noise * skytem_noise_std(clean, tc, (26, 19))
        kw = dict(kw0)
        kw["prng"] = get_prng(seed=(4000 if height else 2000) + sidx)
        if height:
            kw.update(TX_HEIGHT_KW)
            z = z + HEIGHT_BIAS
        inf = Inference1D(**kw)
        dp = _tdem_datapoint(data, z)
        with contextlib.redirect_stdout(io.StringIO()):
            inf.initialize(dp)
        hs = float(inf.halfspace.item())
        init = dict(prior=float(inf.prior), likelihood=float(inf.likelihood), misfit=float(inf.data_misfit))
        for it in range(n_iter):
            dp0, m0 = inf.datapoint, inf.model
            J_in = np.asarray(dp0.sensitivity_matrix).copy()
            pred_in = np.asarray(dp0.predictedData).copy()
            rel_cur = np.asarray(dp0.relative_error).copy()
            add_cur = np.asarray(dp0.additive_error).copy()
            z_cur = float(np.asarray(dp0.transmitter.z).item())
            tdp = deepcopy(dp0)
            remapped, test = m0.perturb(tdp, -npinf, npinf, alpha=inf.covariance_scaling)
            action = ACT[remapped.mesh.action[0]]
            k = int(remapped.nCells.item())
            grad = np.asarray(remapped.local_gradient(observation=tdp)).copy()
            H = np.asarray(test.values.proposal.variance).copy()
            mean = np.asarray(test.values.proposal.mean).copy()
            tdp.perturb()
            tdp.forward(test)
            misfit = float(tdp.data_misfit())
            prior = float(tdp.probability) + float(test.probability(inf.solve_parameter, inf.solve_gradient))
            like = float(tdp.likelihood(log=True))
            prop, prop1 = test.proposal_probabilities(remapped, tdp, alpha=inf.covariance_scaling)
            recs.append(dict(sounding=sidx, altitude=z_cur, altitude_ref=z,
                             altitude_test=float(np.asarray(tdp.transmitter.z).item()), sigma_ref=hs, data=np.asarray(data), k=k, action=action,
                             edges=np.asarray(remapped.mesh.edges).copy(), sigma_remap=np.asarray(remapped.values).copy(),
                             sigma_test=np.asarray(test.values).copy(), rel_cur=rel_cur, add_cur=add_cur,
                             rel_test=np.asarray(tdp.relative_error).copy(), add_test=np.asarray(tdp.additive_error).copy(),
                             J_in=J_in, pred_in=pred_in, H=H, gradient=grad, newton_mean=mean,
                             pred_test=np.asarray(tdp.predictedData).copy(), std_test=np.asarray(tdp.std).copy(),
                             misfit_test=misfit, prior_test=prior, likelihood_test=like, proposal=float(prop),
                             proposal1=float(prop1), init_prior=init["prior"], init_likelihood=init["likelihood"],
                             init_misfit=init["misfit"], alpha=float(inf.covariance_scaling)))
            log_alpha = (prior - inf.prior) + (like - inf.likelihood) + (prop - prop1)
            if np.exp(log_alpha) > inf.prng.uniform():
                inf.data_misfit, inf.prior, inf.likelihood = misfit, prior, like
                inf.model, inf.datapoint = test, tdp
        print("tdem sounding", sidx, "k now", inf.model.nCells.item(), "halfspace", hs, flush=True)
    n = len(recs)
    out = {}
    for key in recs[0]:
        vals = [r[key] for r in recs]
        if np.ndim(vals[0]) == 0:
            out[key] = np.asarray(vals)
        else:
            obj = np.empty(n, dtype=object)
            for i, v in enumerate(vals):
                obj[i] = np.asarray(v)
            out[key] = obj
    np.savez_compressed(os.path.join(HERE, "tdem_transitions_height.npz" if height else "tdem_transitions.npz"), **out)
    print("tdem transitions written:", n, "actions", np.bincount(out["action"], minlength=4))


def make_tdem_chain(sidx, rep=0, n_markov_chains=10000, height=False):
    """A full chain of the live reference with a dual-moment TdemDataPoint (skytem_options with n_markov_chains
    reduced), driven through fake_gatdaem1d (oracle forward).  Posterior arrays -> ref_tdem_chain_<i>[_r<rep>].npz."""
    kw = _tdem_setup()
    from geobipy import Inference1D, get_prng
    from geobipy_b200.synthetic import skytem_noise_std
    import oracle_py as O
    import io
    import contextlib
    tsys = O.make_tdem_system()
    tc = np.array(tsys.t_centre[:45])
    edges, sigma, z, noise = synthetic_sounding(sidx, 400.0, 45)
    clean = O.tdem_forward(tsys, z, sigma, np.r_[np.diff(edges)[:-1], 1.0])
    data = clean + noise * skytem_noise_std(clean, tc, (26, 19))
    kw["n_markov_chains"] = n_markov_chains
    kw["prng"] = get_prng(seed=(9000 if height else 7000) + sidx + 100 * rep)
    if height:   # transmitter height sampled, starting 0.4 m above the simulated one
        kw.update(TX_HEIGHT_KW)
        z = z + HEIGHT_BIAS
    inf = Inference1D(**kw)
    dp = _tdem_datapoint(data, z)
    t0 = time.time()
    go, failed = True, False
    with contextlib.redirect_stdout(io.StringIO()):
        inf.initialize(dp)
        while go:  # Inference1D.infer :650-677 without the HDF5 write
            failed = inf.accept_reject()
            inf.update()
            go = (not failed) and (inf.iteration <= inf.n_markov_chains + inf.burned_in_iteration)
            if (not failed) and (not inf.burned_in):
                go = inf.iteration < inf.n_markov_chains
                if not go:
                    failed = True
    dt = time.time() - t0
    it = int(inf.iteration)
    rel = np.stack([np.asarray(p.counts, dtype=np.int32) for p in inf.datapoint.relative_error.posterior])
    add = np.stack([np.asarray(p.counts, dtype=np.int32) for p in inf.datapoint.additive_error.posterior])
    extra, stem = {}, "ref_tdem_chain_"
    if height:
        stem = "ref_tdem_height_chain_"
        extra = dict(height_hist=np.asarray(inf.datapoint.transmitter.z.posterior.counts, dtype=np.int32),
                     height_cur=float(np.asarray(inf.datapoint.transmitter.z).item()),
                     height_best=float(np.asarray(inf.best_datapoint.transmitter.z).item()))
    np.savez_compressed(
        os.path.join(HERE, stem + ("%d.npz" % sidx if rep == 0 else "%d_r%d.npz" % (sidx, rep))), **extra,
        sounding=sidx, data=data, altitude=z, true_edges=edges, true_sigma=sigma, halfspace=float(inf.halfspace.item()),
        iterations=it, failed=bool(failed), burned_in=bool(inf.burned_in), burned_in_iteration=int(inf.burned_in_iteration),
        hitmap=np.asarray(inf.model.values.posterior.counts, dtype=np.int32),
        edges_hist=np.asarray(inf.model.mesh.edges.posterior.counts, dtype=np.int32),
        ncells_hist=np.asarray(inf.model.mesh.nCells.posterior.counts, dtype=np.int32), rel_hist=rel, add_hist=add,
        misfit_trace=np.asarray(inf.data_misfit_v[:it], dtype=np.float32),
        accept_trace=np.asarray(inf.acceptance_v[:it + 1], dtype=np.uint8), seconds=dt, n_markov_chains=n_markov_chains)
    print("tdem chain", sidx, rep, "iterations", it, "burned in", inf.burned_in, inf.burned_in_iteration, "s/it", dt / it,
          "acc", np.asarray(inf.acceptance_v[:it + 1]).mean(), flush=True)


def make_chain(sidx, rep=0, n_markov_chains=10000, height=False):
    _geobipy()
    data, z, edges, sigma = _observed(sidx)
    if height:
        z = z + HEIGHT_BIAS
        inf = _initialised_inference(data, z, n_markov_chains, 7000 + sidx + 100 * rep, **HEIGHT_KW)
    else:
        inf = _initialised_inference(data, z, n_markov_chains, 5000 + sidx + 100 * rep)
    t0 = time.time()
    import io
    import contextlib
    go, failed = True, False
    with contextlib.redirect_stdout(io.StringIO()):
        while go:  # Inference1D.infer :650-677 without the HDF5 write
            failed = inf.accept_reject()
            inf.update()
            go = (not failed) and (inf.iteration <= inf.n_markov_chains + inf.burned_in_iteration)
            if (not failed) and (not inf.burned_in):
                go = inf.iteration < inf.n_markov_chains
                if not go:
                    failed = True
    dt = time.time() - t0
    it = int(inf.iteration)
    extra = {}
    stem = "ref_chain_"
    if height:
        stem = "ref_height_chain_"
        extra = dict(height_hist=np.asarray(inf.datapoint.z.posterior.counts, dtype=np.int32),
                     height_cur=float(np.asarray(inf.datapoint.z).item()),
                     height_best=float(np.asarray(inf.best_datapoint.z).item()))
    np.savez_compressed(
        os.path.join(HERE, stem + ("%d.npz" % sidx if rep == 0 else "%d_r%d.npz" % (sidx, rep))), **extra,
        sounding=sidx, data=data, altitude=z, true_edges=edges,
        true_sigma=sigma, halfspace=float(inf.halfspace.item()), iterations=it, failed=bool(failed),
        burned_in=bool(inf.burned_in), burned_in_iteration=int(inf.burned_in_iteration),
        hitmap=np.asarray(inf.model.values.posterior.counts, dtype=np.int32),
        edges_hist=np.asarray(inf.model.mesh.edges.posterior.counts, dtype=np.int32),
        ncells_hist=np.asarray(inf.model.mesh.nCells.posterior.counts, dtype=np.int32),
        rel_hist=np.asarray(inf.datapoint.relative_error.posterior.counts, dtype=np.int32),
        add_hist=np.asarray(inf.datapoint.additive_error.posterior.counts, dtype=np.int32),
        misfit_trace=np.asarray(inf.data_misfit_v[:it], dtype=np.float32),
        accept_trace=np.asarray(inf.acceptance_v[:it + 1], dtype=np.uint8),
        seconds=dt, n_markov_chains=n_markov_chains)
    print("chain", sidx, "iterations", it, "burned in", inf.burned_in, inf.burned_in_iteration, "s/it", dt / it)


def make_summaries():
    """posterior_summaries.npz: what the reference's OWN Histogram / Mesh methods return for recorded hitmaps (the 2-D
    conductivity-depth posteriors of ref_chain_{1,2,3}): Histogram.mean / median / mode / percentile / credible_range /
    transparency / opacity / opacity_level (classes/statistics/Histogram.py:262-401, :509-542 over Mesh._mean /
    _percentile / _mode / _credible_range, classes/mesh/Mesh.py:30-217).  The Histogram object is the live one of an
    initialised Inference1D (Model.set_posteriors, Model.py:666-684: conductivity axis log = 10, relative to the
    half-space), its counts replaced by the recorded ones."""
    _geobipy()
    out = {}
    for n, sidx in enumerate((1, 2, 3)):
        g = np.load(os.path.join(HERE, "ref_chain_%d.npz" % sidx))
        inf = _initialised_inference(g["data"], float(g["altitude"]), 100, 1)
        H = inf.model.values.posterior
        assert abs(float(inf.halfspace.item()) / float(g["halfspace"]) - 1.0) < 1e-12
        assert tuple(H.counts.shape) == tuple(g["hitmap"].shape)
        H.values = np.asarray(g["hitmap"], dtype=np.float64).copy()   # Histogram.values setter -> counts
        assert np.array_equal(np.asarray(H.counts), g["hitmap"])
        rec = dict(hitmap=g["hitmap"], halfspace=float(g["halfspace"]),
                   x_edges=np.asarray(H.mesh.x.edges_absolute, dtype=np.float64), y_edges=np.asarray(H.mesh.y.edges, dtype=np.float64),
                   mean=np.asarray(H.mean(axis=0).values), median=np.asarray(H.median(axis=0).values),
                   mode=np.asarray(H.mode(axis=0).values), p5=np.asarray(H.percentile(5.0, axis=0).values),
                   p95=np.asarray(H.percentile(95.0, axis=0).values),
                   credible_range90=np.asarray(H.credible_range(percent=90.0, log=10, axis=0)),
                   transparency90=np.asarray(H.transparency(percent=90.0, log=10, axis=0).values),
                   opacity90=np.asarray(H.opacity(percent=90.0, log=10, axis=0).values),
                   transparency95=np.asarray(H.transparency(percent=95.0, axis=0).values),
                   opacity_level95=float(H.opacity_level(percent=95.0, axis=0)))
        for k, v in rec.items():
            out["s%d_%s" % (n, k)] = np.asarray(v, dtype=np.float64) if k != "hitmap" else v
    np.savez_compressed(os.path.join(HERE, "posterior_summaries.npz"), soundings=np.array([1, 2, 3]), **out)
    print("posterior summaries written", {k: np.shape(v) for k, v in out.items() if k.startswith("s0_")})


def make_height_reset():
    """What Inference1D.reset() (:984-994) does to a sampled height: recorded from the live reference."""
    import io
    import contextlib
    import json
    _geobipy()
    data, z, _, _ = _observed(1)
    z = z + HEIGHT_BIAS
    inf = _initialised_inference(data, z, 2000, 5, **HEIGHT_KW)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(200):
            inf.accept_reject()
            inf.update()
    d = inf.datapoint
    rec = dict(z_input=float(z), z_before_reset=float(d.z.item()),
               prior_before=[float(d.z.prior._min.item()), float(d.z.prior._max.item())],
               relative_error_before=float(np.asarray(d.relative_error).item()))
    with contextlib.redirect_stdout(io.StringIO()):
        inf.reset()
    d = inf.datapoint
    rec.update(z_after_reset=float(d.z.item()), prior_after=[float(d.z.prior._min.item()), float(d.z.prior._max.item())],
               proposal_mean_after=float(np.asarray(d.z.proposal.mean).item()),
               posterior_relative_to_after=float(np.asarray(d.z.posterior.mesh.relative_to).item()),
               posterior_edges_after=[float(np.asarray(d.z.posterior.mesh.edges)[0]), float(np.asarray(d.z.posterior.mesh.edges)[-1])],
               relative_error_after=float(np.asarray(d.relative_error).item()), iteration_after=int(inf.iteration))
    json.dump(rec, open(os.path.join(HERE, "height_reset.json"), "w"), indent=1)
    print(rec)


def _h5lite_as_h5py():
    """The reference writes its result files through h5py, which this image lacks: hand it geobipy_b200.h5lite (the same
    calls over an in-memory tree) under that name, BEFORE the reference is imported."""
    import types
    from geobipy_b200 import h5lite
    m = types.ModuleType("h5py")
    m.File, m.Group, m.Dataset, m._hl = h5lite.File, h5lite.Group, h5lite.Dataset, h5lite._hl
    sys.modules["h5py"] = m
    return h5lite


def _dump_tree(f, h5lite):
    """{path: array} + {path: {kind, dtype, shape, attrs}} of an h5lite tree."""
    arrays, meta = {}, {}

    def visit(name, obj):
        attrs = {k: (v if isinstance(v, str) else np.asarray(v).tolist()) for k, v in obj.attrs.items()}
        if isinstance(obj, h5lite.Dataset):
            a = np.asarray(obj)
            meta[name] = dict(kind="dataset", dtype=a.dtype.str, shape=list(a.shape), attrs=attrs)
            arrays[name] = a
        else:
            meta[name] = dict(kind="group", attrs=attrs)
    f.visititems(visit)
    return arrays, meta


def make_hdf(kind="fdem", n_iter=1200, n_points=3, index=1):
    """hdf_layout_<kind>.npz: the tree the reference's OWN Inference2D.createHdf / Inference1D.writeHdf code builds
    (inversion/Inference2D.py:2001-2016, Inference1D.py:1002-1090) for a line of `n_points` soundings with one sounding
    written at `index`, recorded through h5lite, together with the state of that reference chain in the form of the
    product's result arrays - what geobipy_b200.hdf must turn into the same tree."""
    import io
    import json
    import contextlib
    h5lite = _h5lite_as_h5py()
    height = kind == "fdem_height"
    if kind == "tempest":
        kw = _tempest_setup()
        from geobipy import Inference1D, get_prng
        import oracle_py as O
        tsys = O.make_tdem_system([O.tempest_definition()], rx_offset=(-107.0, 0.0, -45.0))
        sec, prim = _tempest_observed(1, O, tsys)
        z = 120.0
        kw["n_markov_chains"] = 10000
        kw["prng"] = get_prng(seed=4242)
        inf = Inference1D(**kw)
        dp = _tempest_datapoint(sec, prim)
        with contextlib.redirect_stdout(io.StringIO()):
            inf.initialize(dp)
    elif kind in ("tdem", "tdem_height"):
        kw = _tdem_setup()
        if kind == "tdem_height":
            kw.update(TX_HEIGHT_KW)
        from geobipy import Inference1D, get_prng
        from geobipy_b200.synthetic import skytem_noise_std
        import oracle_py as O
        tsys = O.make_tdem_system()
        tc = np.array(tsys.t_centre[:45])
        edges, sigma, z, noise = synthetic_sounding(2, 400.0, 45)
        clean = O.tdem_forward(tsys, z, sigma, np.r_[np.diff(edges)[:-1], 1.0])
        data = clean + noise * skytem_noise_std(clean, tc, (26, 19))
        kw["n_markov_chains"] = 10000
        kw["prng"] = get_prng(seed=4242)
        inf = Inference1D(**kw)
        dp = _tdem_datapoint(data, z)
        with contextlib.redirect_stdout(io.StringIO()):
            inf.initialize(dp)
    else:
        _geobipy()
        data, z, edges, sigma = _observed(1)
        if height:
            z = z + HEIGHT_BIAS
        inf = _initialised_inference(data, z, 10000, 4242, **(HEIGHT_KW if height else {}))
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(n_iter):
            inf.accept_reject()
            inf.update()
    from geobipy import StatArray
    fid = np.arange(n_points, dtype=np.float64) * 10.0 + 10.0
    inf.datapoint.fiducial = StatArray(fid[index], "fiducial")
    inf.datapoint.lineNumber = StatArray(100.0, "Line number")
    inf.best_datapoint.fiducial = StatArray(fid[index], "fiducial")
    inf.best_datapoint.lineNumber = StatArray(100.0, "Line number")
    f = h5lite.File(os.path.join("/tmp", "hdf_layout_%s.h5" % kind), "w")
    with contextlib.redirect_stdout(io.StringIO()):
        inf.createHdf(f, add_axis=StatArray(fid, "fiducial"))        # what Inference2D.createHdf does (:2006-2014)
        StatArray(np.full(n_points, 100.0), "Line number").writeHdf(f, "data/line_number")
        StatArray(fid, "fiducial").writeHdf(f, "data/fiducial")
        inf.writeHdf(f, index=index)
    arrays, meta = _dump_tree(f, h5lite)
    f.close()

    # the same chain as the product's result arrays (include/geobipy_b200.h gbp_chain_buffers)
    m, bm, d, bd = inf.model, inf.best_model, inf.datapoint, inf.best_datapoint
    ml = 30
    N2 = 2 * inf.n_markov_chains

    def padded(a, n, fill=np.nan):
        out = np.full(n, fill)
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        out[:a.size] = a
        return out
    ns = 2 if kind in ("tdem", "tdem_height", "tempest") else 1
    # a Tempest datapoint samples the multiplier of fixed per-channel additive levels (Tempest_datapoint.py:85-105)
    add_of = (lambda q: q.additive_error_multiplier) if kind == "tempest" else (lambda q: q.additive_error)
    rel_hist = np.stack([np.asarray(d.relative_error.posterior[i].counts if ns > 1 else d.relative_error.posterior.counts, dtype=np.int32) for i in range(ns)])
    add_hist = np.stack([np.asarray(add_of(d).posterior[i].counts if ns > 1 else add_of(d).posterior.counts, dtype=np.int32) for i in range(ns)])
    state = dict(
        hitmap=np.asarray(m.values.posterior.counts, dtype=np.int32), edges_hist=np.asarray(m.mesh.edges.posterior.counts, dtype=np.int32),
        ncells_hist=np.asarray(m.mesh.nCells.posterior.counts, dtype=np.int32), rel_hist=rel_hist if ns > 1 else rel_hist[0],
        add_hist=add_hist if ns > 1 else add_hist[0],
        misfit_trace=padded(np.asarray(inf.data_misfit_v), N2, 0.0), accept_trace=padded(np.asarray(inf.acceptance_v), N2, 0.0).astype(np.uint8),
        best_sigma=padded(bm.values, ml), best_edges=padded(bm.mesh.edges, ml + 1),
        cur_sigma=padded(m.values, ml), cur_edges=padded(m.mesh.edges, ml + 1),
        iteration=int(inf.iteration), burned_in=bool(inf.burned_in), burned_in_iteration=int(inf.burned_in_iteration),
        best_iteration=int(inf.best_iteration), best_k=int(bm.nCells.item()), cur_k=int(m.nCells.item()),
        halfspace=float(inf.halfspace.item()), multiplier=float(inf.multiplier),
        cur_rel=np.asarray(d.relative_error, dtype=np.float64), cur_add=np.asarray(add_of(d), dtype=np.float64),
        best_rel=np.asarray(bd.relative_error, dtype=np.float64), best_add=np.asarray(add_of(bd), dtype=np.float64),
        data=np.asarray(d.data, dtype=np.float64), std_best=np.asarray(bd.std, dtype=np.float64),
        predicted_best=np.asarray(bd.predictedData, dtype=np.float64), z_input=float(z),
        x=float(d.x), y=float(d.y), elevation=float(d.elevation), fiducial=fid, line_number=100.0, index=index, n_points=n_points)
    if kind in ("tdem", "tdem_height", "tempest"):
        state.update(tx_z=float(np.asarray(bd.transmitter.z).item()), rx_z=float(np.asarray(bd.receiver.z).item()))
    if kind == "tempest":
        state.update(secondary=sec, primary=prim, additive_levels=np.asarray(bd.additive_error, dtype=np.float64),
                     predicted_primary_best=np.asarray(bd.predicted_primary_field, dtype=np.float64),
                     predicted_secondary_best=np.asarray(bd.predicted_secondary_field, dtype=np.float64))
    if kind == "tdem_height":
        tz = d.transmitter.z
        state.update(height_hist=np.asarray(tz.posterior.counts, dtype=np.int32), cur_height=float(np.asarray(tz).item()),
                     best_height=float(np.asarray(bd.transmitter.z).item()),
                     height_edges=np.asarray(tz.posterior.mesh.edges, dtype=np.float64),
                     height_relative_to=float(np.asarray(tz.posterior.mesh.relative_to).item()))
    if height:
        state.update(height_hist=np.asarray(d.z.posterior.counts, dtype=np.int32), cur_height=float(np.asarray(d.z).item()),
                     best_height=float(np.asarray(bd.z).item()),
                     height_edges=np.asarray(d.z.posterior.mesh.edges, dtype=np.float64))
    out = {"tree/" + k: v for k, v in arrays.items()}
    out.update({"state/" + k: np.asarray(v) for k, v in state.items()})
    np.savez_compressed(os.path.join(HERE, "hdf_layout_%s.npz" % kind), meta=json.dumps(meta), **out)
    print(kind, "tree entries", len(meta), "datasets", len(arrays), "iteration", inf.iteration, "k", state["cur_k"], state["best_k"])


def make_readers():
    """reader_files/ + readers.npz: heads of the data files the reference ships (documentation_source/source/
    supplementary/data: a comma-separated RESOLVE file, a whitespace-separated field file with other column names, a
    SkyTEM dual-moment file) and what the reference's OWN readers make of exactly those heads
    (FdemData.read_csv classes/data/dataset/FdemData.py:520-610, TdemData.read_csv classes/data/dataset/TdemData.py:
    401-560)."""
    import fake_gatdaem1d
    fake_gatdaem1d.install()
    _geobipy()
    from geobipy import FdemData, TdemData
    out_dir = os.path.join(HERE, "reader_files")
    os.makedirs(out_dir, exist_ok=True)
    data_dir = os.path.join(SUP, "data")
    rec = {}
    for name, rows in (("resolve_glacial.csv", 12), ("Resolve_small.txt", 40), ("Resolve_single.txt", 2), ("skytem_glacial.csv", 7)):
        with open(os.path.join(data_dir, name)) as f:
            head = [next(f) for _ in range(rows)] if rows else f.readlines()
        with open(os.path.join(out_dir, name), "w") as f:
            f.writelines(head)
    for name in ("resolve.stm", "FdemSystem1.stm", "SkytemHM.stm", "SkytemLM.stm"):
        with open(os.path.join(data_dir, name)) as f, open(os.path.join(out_dir, name), "w") as g:
            g.write(f.read())
    for name, stm in (("resolve_glacial.csv", "resolve.stm"), ("Resolve_small.txt", "FdemSystem1.stm"), ("Resolve_single.txt", "FdemSystem1.stm")):
        d = FdemData.read_csv(os.path.join(out_dir, name), os.path.join(out_dir, stm))
        key = name.split(".")[0]
        for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data", "std"):
            rec[key + "/" + k] = np.asarray(getattr(d, k), dtype=np.float64)
        rec[key + "/frequencies"] = np.asarray(d.system[0].frequencies, dtype=np.float64)
        rec[key + "/loop_separation"] = np.asarray(d.system[0].loop_separation, dtype=np.float64)
        print(name, d.nPoints, "points", rec[key + "/data"].shape)
    d = TdemData.read_csv(os.path.join(out_dir, "skytem_glacial.csv"), [os.path.join(out_dir, "SkytemHM.stm"), os.path.join(out_dir, "SkytemLM.stm")])
    for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data", "std"):
        rec["skytem_glacial/" + k] = np.asarray(getattr(d, k), dtype=np.float64)
    for who in ("transmitter", "receiver"):
        lp = getattr(d.loop_pair, who)
        for k in ("x", "y", "z", "pitch", "roll", "yaw", "radius"):
            rec["skytem_glacial/%s_%s" % (who, k)] = np.asarray(getattr(lp, k), dtype=np.float64)
    rec["skytem_glacial/off_time0"] = np.asarray(d.system[0].off_time, dtype=np.float64)
    rec["skytem_glacial/off_time1"] = np.asarray(d.system[1].off_time, dtype=np.float64)
    print("skytem", d.nPoints, "points", rec["skytem_glacial/data"].shape)
    np.savez_compressed(os.path.join(HERE, "readers.npz"), **rec)


def make_readers_tempest():
    """reader_files/tempest_glacial.csv + tempest.stm and readers_tempest.npz: the head of the Tempest file the reference
    ships and what the reference's OWN TempestData.read_csv (classes/data/dataset/TempestData.py:140-273) makes of it."""
    import fake_gatdaem1d
    fake_gatdaem1d.install()
    _geobipy()
    from geobipy import TempestData
    out_dir = os.path.join(HERE, "reader_files")
    data_dir = os.path.join(SUP, "data")
    with open(os.path.join(data_dir, "tempest_glacial.csv")) as f:
        head = [next(f) for _ in range(7)]
    with open(os.path.join(out_dir, "tempest_glacial.csv"), "w") as f:
        f.writelines(head)
    with open(os.path.join(data_dir, "tempest.stm")) as f, open(os.path.join(out_dir, "tempest.stm"), "w") as g:
        g.write(f.read())
    d = TempestData.read_csv(os.path.join(out_dir, "tempest_glacial.csv"), os.path.join(out_dir, "tempest.stm"))
    rec = {}
    for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data", "std", "primary_field", "secondary_field", "relative_error",
              "additive_error", "additive_error_multiplier"):
        rec[k] = np.array(getattr(d, k), dtype=np.float64, copy=True)
    for who in ("transmitter", "receiver"):
        lp = getattr(d.loop_pair, who)
        for k in ("x", "y", "z", "pitch", "roll", "yaw", "radius"):
            rec["%s_%s" % (who, k)] = np.asarray(getattr(lp, k), dtype=np.float64)
    rec["off_time0"] = np.asarray(d.system[0].off_time, dtype=np.float64)
    rec["components"] = np.asarray(d.components)
    d.relative_error = np.tile([0.001, 0.002], (d.nPoints, 1))     # (the file gives none: 0, which a datapoint refuses)
    d.additive_error = np.tile(np.linspace(0.01, 0.02, 30), (d.nPoints, 1))
    dp = d.datapoint(2)
    rec["dp2_data"], rec["dp2_std"] = np.asarray(dp.data, dtype=np.float64), np.asarray(dp.std, dtype=np.float64)
    rec["std_with_errors"] = np.asarray(d.std, dtype=np.float64)
    print("tempest", d.nPoints, "points", rec["data"].shape, rec["primary_field"].shape, rec["additive_error"].shape, rec["components"])
    np.savez_compressed(os.path.join(HERE, "readers_tempest.npz"), **rec)


def make_tempest():
    """tempest_clean.npz: the reference's Tempest known-answer CSVs (tests/data_checks/tempest_*_clean.csv, tests/
    test_synthetic_data.py:51-66; plain CSV, no reference import needed): 6 models x 79 soundings x (15 X + 15 Z windows) of
    B-field in fT plus the primary field PX, PZ - a second system (point dipole, square wave, OutputType = B), a second
    geometry (120 m, receiver 107 m behind and 45 m below) and a second component for the time-domain arithmetic."""
    import pandas as pd
    data = np.zeros((6, 79, 30))
    prim = np.zeros((6, 79, 2))
    sig = np.array([[1e-2, 1e-1, 0.03333333], [1e-2, 1e-1, 1.0], [2e-2, 2e-3, 2e-2], [1e-2, 1e-1, 1e-4],
                    [1.0, 1e-2, 5e-2], [1e-4, 1e-2, 1.0]])  # Model.create_synthetic_model, Model.py:902-908
    geom = None
    for m, name in enumerate(MODELS):
        df = pd.read_csv(os.path.join(REF, "tests/data_checks/tempest_%s_clean.csv" % name))
        xc = [c for c in df.columns if c.startswith("S0X")]
        zc = [c for c in df.columns if c.startswith("S0Z")]
        data[m] = df[xc + zc].values
        prim[m] = df[["PX", "PZ"]].values
        g = df[["Height", "tx_pitch", "tx_roll", "tx_yaw", "txrx_dx", "txrx_dy", "txrx_dz", "rx_pitch", "rx_roll", "rx_yaw"]].values
        assert np.all(g == g[0])
        geom = g[0]
        times = np.array([float(c.split("_")[-1]) for c in xc])
    np.savez_compressed(os.path.join(HERE, "tempest_clean.npz"), data=data, primary=prim, sigma=sig,
                        zwedge=np.linspace(50.0, 1.0, 79), zdeep=np.linspace(75.0, 500.0, 79), geometry=geom,
                        times=times, models=np.array(MODELS))
    print("tempest", data.shape, geom)


def _tempest_setup():
    import fake_gatdaem1d
    fake_gatdaem1d.install()
    _geobipy()
    from geobipy import user_parameters
    kw = dict(user_parameters.read(os.path.join(SUP, "options_files/tempest_options")))
    kw["interactive_plot"] = False
    kw["save_hdf5"] = True
    for k in ("data_type", "data_filename", "system_filename", "data_directory", "seed"):
        kw.pop(k, None)
    return kw


def _tempest_datapoint(secondary, primary, z=120.0):
    from geobipy import Tempest_datapoint, CircularLoop
    tx = CircularLoop(x=0.0, y=0.0, z=z, pitch=0.0, roll=0.0, yaw=0.0, radius=1.0)
    rx = CircularLoop(x=-107.0, y=0.0, z=z - 45.0, pitch=0.0, roll=0.0, yaw=0.0, radius=1.0)
    return Tempest_datapoint(x=0.0, y=0.0, z=z, elevation=0.0, secondary_field=secondary, primary_field=primary,
                             system=[os.path.join(SUP, "data/tempest.stm")], transmitter_loop=tx, receiver_loop=rx)


def _tempest_observed(sidx, O, tsys):
    """Synthetic Tempest sounding `sidx`: the shared true-model generator, 120 m, noise = tempest_options' error model."""
    edges, sigma, _, noise = synthetic_sounding(sidx, 400.0, 30)
    clean = O.tdem_forward(tsys, 120.0, sigma, np.r_[np.diff(edges)[:-1], 1.0])
    prim = np.repeat(O.tdem_primary_field(O.tempest_definition(), (-107.0, 0.0, -45.0)), 15)
    std = np.sqrt((0.001 * (clean + prim)) ** 2 + O.TEMPEST_ADDITIVE ** 2)
    return clean + noise * std, prim[[0, 15]]


def make_tempest_transitions(n_soundings=4, n_iter=200):
    """tempest_transitions.npz: per-term records of the reference's Inference1D.accept_reject with a Tempest_datapoint and
    tempest_options (errors per component, additive-error multiplier, data = secondary + primary field), the external
    gatdaem1d replaced by tests/golden/fake_gatdaem1d.py over the oracle's forward."""
    kw0 = _tempest_setup()
    from geobipy import Inference1D, get_prng
    import oracle_py as O
    from numpy import inf as npinf
    import io
    import contextlib
    tsys = O.make_tdem_system([O.tempest_definition()], rx_offset=(-107.0, 0.0, -45.0))
    recs = []
    for sidx in range(n_soundings):
        sec, prim = _tempest_observed(sidx, O, tsys)
        kw = dict(kw0)
        kw["prng"] = get_prng(seed=6000 + sidx)
        inf = Inference1D(**kw)
        dp = _tempest_datapoint(sec, prim)
        with contextlib.redirect_stdout(io.StringIO()):
            inf.initialize(dp)
        hs = float(inf.halfspace.item())
        init = dict(prior=float(inf.prior), likelihood=float(inf.likelihood), misfit=float(inf.data_misfit))
        for it in range(n_iter):
            dp0, m0 = inf.datapoint, inf.model
            J_in = np.asarray(dp0.sensitivity_matrix).copy()
            pred_in = np.asarray(dp0.predictedData).copy()
            rel_cur = np.asarray(dp0.relative_error).copy()
            add_cur = np.asarray(dp0.additive_error_multiplier).copy()
            tdp = deepcopy(dp0)
            remapped, test = m0.perturb(tdp, -npinf, npinf, alpha=inf.covariance_scaling)
            action = ACT[remapped.mesh.action[0]]
            k = int(remapped.nCells.item())
            grad = np.asarray(remapped.local_gradient(observation=tdp)).copy()
            H = np.asarray(test.values.proposal.variance).copy()
            mean = np.asarray(test.values.proposal.mean).copy()
            tdp.perturb()
            tdp.forward(test)
            misfit = float(tdp.data_misfit())
            prior = float(tdp.probability) + float(test.probability(inf.solve_parameter, inf.solve_gradient))
            like = float(tdp.likelihood(log=True))
            prop, prop1 = test.proposal_probabilities(remapped, tdp, alpha=inf.covariance_scaling)
            recs.append(dict(sounding=sidx, altitude=120.0, sigma_ref=hs, data=np.asarray(dp0.data).copy(), k=k, action=action,
                             edges=np.asarray(remapped.mesh.edges).copy(), sigma_remap=np.asarray(remapped.values).copy(),
                             sigma_test=np.asarray(test.values).copy(), rel_cur=rel_cur, add_cur=add_cur,
                             rel_test=np.asarray(tdp.relative_error).copy(), add_test=np.asarray(tdp.additive_error_multiplier).copy(),
                             J_in=J_in, pred_in=pred_in, H=H, gradient=grad, newton_mean=mean,
                             pred_test=np.asarray(tdp.predictedData).copy(), std_test=np.asarray(tdp.std).copy(),
                             misfit_test=misfit, prior_test=prior, likelihood_test=like, proposal=float(prop),
                             proposal1=float(prop1), init_prior=init["prior"], init_likelihood=init["likelihood"],
                             init_misfit=init["misfit"], alpha=float(inf.covariance_scaling)))
            log_alpha = (prior - inf.prior) + (like - inf.likelihood) + (prop - prop1)
            if np.exp(log_alpha) > inf.prng.uniform():
                inf.data_misfit, inf.prior, inf.likelihood = misfit, prior, like
                inf.model, inf.datapoint = test, tdp
        print("tempest sounding", sidx, "k now", inf.model.nCells.item(), "halfspace", hs, flush=True)
    n = len(recs)
    out = {}
    for key in recs[0]:
        vals = [r[key] for r in recs]
        if np.ndim(vals[0]) == 0:
            out[key] = np.asarray(vals)
        else:
            obj = np.empty(n, dtype=object)
            for i, v in enumerate(vals):
                obj[i] = np.asarray(v)
            out[key] = obj
    np.savez_compressed(os.path.join(HERE, "tempest_transitions.npz"), **out)
    print("tempest transitions written:", n, "actions", np.bincount(out["action"], minlength=4))


def make_tempest_chain(sidx, rep=0, n_markov_chains=10000):
    """A full chain of the live reference with a Tempest_datapoint (tempest_options with n_markov_chains raised to 10 000),
    driven through fake_gatdaem1d (oracle forward).  Posterior arrays -> ref_tempest_chain_<i>[_r<rep>].npz."""
    kw = _tempest_setup()
    from geobipy import Inference1D, get_prng
    import oracle_py as O
    import io
    import contextlib
    tsys = O.make_tdem_system([O.tempest_definition()], rx_offset=(-107.0, 0.0, -45.0))
    sec, prim = _tempest_observed(sidx, O, tsys)
    kw["n_markov_chains"] = n_markov_chains
    kw["prng"] = get_prng(seed=8000 + sidx + 100 * rep)
    inf = Inference1D(**kw)
    dp = _tempest_datapoint(sec, prim)
    t0 = time.time()
    go, failed = True, False
    with contextlib.redirect_stdout(io.StringIO()):
        inf.initialize(dp)
        while go:  # Inference1D.infer :650-677 without the HDF5 write
            failed = inf.accept_reject()
            inf.update()
            go = (not failed) and (inf.iteration <= inf.n_markov_chains + inf.burned_in_iteration)
            if (not failed) and (not inf.burned_in):
                go = inf.iteration < inf.n_markov_chains
                if not go:
                    failed = True
    dt = time.time() - t0
    it = int(inf.iteration)
    d = inf.datapoint
    np.savez_compressed(
        os.path.join(HERE, "ref_tempest_chain_%d%s.npz" % (sidx, "" if rep == 0 else "_r%d" % rep)),
        sounding=sidx, data=np.asarray(d.data, dtype=np.float64), secondary=sec, primary=prim, altitude=120.0,
        halfspace=float(inf.halfspace.item()), iterations=it, failed=bool(failed), burned_in=bool(inf.burned_in),
        burned_in_iteration=int(inf.burned_in_iteration),
        hitmap=np.asarray(inf.model.values.posterior.counts, dtype=np.int32),
        edges_hist=np.asarray(inf.model.mesh.edges.posterior.counts, dtype=np.int32),
        ncells_hist=np.asarray(inf.model.mesh.nCells.posterior.counts, dtype=np.int32),
        rel_hist=np.stack([np.asarray(p.counts, dtype=np.int32) for p in d.relative_error.posterior]),
        add_hist=np.stack([np.asarray(p.counts, dtype=np.int32) for p in d.additive_error_multiplier.posterior]),
        misfit_trace=np.asarray(inf.data_misfit_v[:it], dtype=np.float32),
        accept_trace=np.asarray(inf.acceptance_v[:it + 1], dtype=np.uint8), seconds=dt, n_markov_chains=n_markov_chains)
    print("tempest chain", sidx, rep, "iterations", it, "burned in", inf.burned_in, inf.burned_in_iteration, "s/it", dt / it)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "tempest_chain":
        make_tempest_chain(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    if what == "tempest_transitions":
        make_tempest_transitions()
    if what == "tempest":
        make_tempest()
    if what == "readers":
        make_readers()
    if what == "readers_tempest":
        make_readers_tempest()
    if what == "hdf":
        make_hdf(sys.argv[2] if len(sys.argv) > 2 else "fdem")
    if what == "tdem":
        make_tdem()
    if what == "tdem_transitions":
        make_tdem_transitions()
    if what == "tdem_transitions_height":
        make_tdem_transitions(n_soundings=3, n_iter=200, height=True)
    if what == "tdem_chain":
        make_tdem_chain(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    if what == "tdem_height_chain":
        make_tdem_chain(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0, height=True)
    if what == "fdem_tensor":
        make_fdem_tensor()
    if what == "fdem":
        make_fdem()
    elif what == "bins":
        make_bins()
    elif what == "summaries":
        make_summaries()
    elif what == "transitions":
        make_transitions()
    elif what == "chain":
        make_chain(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    elif what == "transitions_height":
        make_transitions(n_soundings=3, n_iter=250, height=True)
    elif what == "height_reset":
        make_height_reset()
    elif what == "height_chain":
        make_chain(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0, height=True)
