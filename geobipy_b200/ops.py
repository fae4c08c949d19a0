"""Batched operators over the C-ABI: FDEM forward / Jacobian and the rjMCMC sampler.

Two calling conventions, as in include/geobipy_b200.h:
  * numpy arrays  -> the ``*_host`` entry points (copies inside the call);
  * torch CUDA tensors -> the device entry points, stream-ordered on torch's current stream.
PyTorch is used for device memory / streams only.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import (BUFFER_FIELDS, NSCALARS, PRECISION_F32, PRECISION_F64, ChainBuffersC, FdemSystemC, OptionsC,
                   TdemSurveyC, TdemSystemC)

_ORI = {"x": 0, "y": 1, "z": 2}


def make_system_struct(freq, tor, tmom, tx, ty, tz, ror, rmom, rx, ry, rz):
    """gbp_fdem_system from the columns of an .stm file (FdemSystem.read, FdemSystem.py:146-183)."""
    n = len(freq)
    if n < 1 or n > _lib.MAXF:
        raise ValueError("a system must have 1..%d frequencies" % _lib.MAXF)
    s = FdemSystemC()
    s.n_freq = n
    for i in range(n):
        to = tor[i] if isinstance(tor[i], (int, np.integer)) else _ORI[str(tor[i]).strip()]
        ro = ror[i] if isinstance(ror[i], (int, np.integer)) else _ORI[str(ror[i]).strip()]
        s.tid[i] = 1 + 3 * int(ro) + int(to)  # FdemSystem.tensor_id, FdemSystem.py:199-203
        s.freq[i], s.tmom[i], s.tx[i], s.ty[i], s.tz[i] = float(freq[i]), float(tmom[i]), float(tx[i]), float(ty[i]), float(tz[i])
        s.rmom[i], s.rx[i], s.ry[i], s.rz[i] = float(rmom[i]), float(rx[i]), float(ry[i]), float(rz[i])
    return s


def resolve_system_struct():
    """RESOLVE (documentation_source/source/supplementary/data/resolve.stm)."""
    return make_system_struct(
        [380.0, 1776.0, 3345.0, 8171.0, 41020.0, 129550.0], list("zzxzzz"), [1, 1, -1, 1, 1, 1], [0] * 6, [0] * 6,
        [0] * 6, list("zzxzzz"), [1] * 6, [7.93, 7.91, 9.03, 7.91, 7.91, 7.89], [0] * 6, [0] * 6)


def make_tdem_system_struct(d):
    """gbp_tdem_system from a parsed .stm description (tdem.read_stm / geobipy_b200/data/*.json)."""
    s = TdemSystemC()
    nw, nwin, nf = len(d["waveform_time"]), len(d["window_start"]), len(d["filter_cutoff"])
    if nw > _lib.TD_MAXWAVE or nwin > _lib.TD_MAXWIN or nf > _lib.TD_MAXFILT:
        raise ValueError("time-domain system too large for the C-ABI limits")
    if d.get("y_scaling", 0.0) != 0.0:
        raise ValueError("the Y component is not supported (X and Z are)")
    ot = str(d.get("output_type", "dB/dt")).strip().lower()
    if ot not in ("db/dt", "b"):
        raise ValueError("OutputType must be dB/dt or B")
    s.output_type = 1 if ot == "b" else 0
    s.peak_current = float(d.get("peak_current", 1.0))
    s.x_scaling, s.z_scaling = float(d.get("x_scaling", 0.0)), float(d.get("z_scaling", 1.0))
    if s.x_scaling == 0.0 and s.z_scaling == 0.0:
        raise ValueError("a system must measure X or Z")
    s.n_wave, s.n_windows, s.n_filters, s.n_abscissae = nw, nwin, nf, int(d["n_abscissae"])
    s.base_frequency, s.digitising_frequency = float(d["base_frequency"]), float(d["digitising_frequency"])
    s.loop_radius = float(d.get("loop_radius", 0.0))
    for i in range(nw):
        s.wave_time[i], s.wave_current[i] = float(d["waveform_time"][i]), float(d["waveform_current"][i])
    for i in range(nwin):
        s.window_start[i], s.window_end[i] = float(d["window_start"][i]), float(d["window_end"][i])
    for i in range(nf):
        s.filter_cutoff[i], s.filter_order[i] = float(d["filter_cutoff"][i]), int(d["filter_order"][i])
    return s


def make_tdem_survey_struct(definitions, rx_offset=(-13.0, 0.0, 2.0), additive_level=None):
    """gbp_tdem_survey: the systems of one time-domain datapoint type + the transmitter->receiver offset.
    `additive_level` [C] selects the Tempest datapoint's error model for the sampler (the options file's
    initial_additive_error vector: a fixed additive error per channel, the unknowns being one multiplier per component)."""
    if not 1 <= len(definitions) <= _lib.TD_MAXSYS:
        raise ValueError("a time-domain datapoint has 1..%d systems" % _lib.TD_MAXSYS)
    sv = TdemSurveyC()
    sv.n_systems = len(definitions)
    for i, d in enumerate(definitions):
        sv.sys[i] = make_tdem_system_struct(d)
    sv.rx_dx, sv.rx_dy, sv.rx_dz = (float(v) for v in rx_offset)
    if additive_level is not None:
        a = np.asarray(additive_level, dtype=np.float64).reshape(-1)
        if len(definitions) != 1 or a.size != n_channels(sv) or not np.all(a > 0.0):
            raise ValueError("the Tempest error model takes one system and one positive additive level per channel")
        sv.error_model = 1
        for i, v in enumerate(a):
            sv.additive_level[i] = float(v)
    return sv


def skytem_definitions():
    """SkyTEM high / low moment (documentation_source/source/supplementary/data/SkytemHM.stm, SkytemLM.stm)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
    return [json.load(open(os.path.join(d, n))) for n in ("skytem_hm.json", "skytem_lm.json")]


def tempest_definition():
    """The Tempest fixed-wing system of the reference (documentation_source/source/supplementary/data/tempest.stm): point
    dipole, square wave, 15 windows of X and Z B-field in fT."""
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tempest.json")))


# tempest_options (documentation_source/source/supplementary/options_files/tempest_options): the additive error of every
# channel [fT] ...
TEMPEST_ADDITIVE = (0.011474, 0.012810, 0.008507, 0.005154, 0.004742, 0.004477, 0.004168, 0.003539, 0.003352, 0.003213, 0.003161,
                    0.003122, 0.002587, 0.002038, 0.002201, 0.007383, 0.005693, 0.005178, 0.003659, 0.003426, 0.003046, 0.003095,
                    0.003247, 0.002775, 0.002627, 0.002460, 0.002178, 0.001754, 0.001405, 0.001283)
# ... and the sampler options: errors per component (x, z); the "additive error" unknown is the multiplier of those levels
# (initially 1; minimum / maximum_additive_error bound its histogram); minimum_thickness None -> 1.0
TEMPEST_OPTIONS = dict(
    n_markov_chains=1000, min_edge=1.0, max_edge=550.0, min_width=1.0, covariance_scaling=0.5, gradient_std=5.0, n_systems=2,
    rel_init=0.001, rel_min=0.0001, rel_max=0.01, rel_prop_var=1e-6, add_init=1.0, add_min=0.001, add_max=100.0, add_prop_var=1e-6,
    rel_init2=0.001, rel_min2=0.0001, rel_max2=0.01, rel_prop_var2=1e-6, add_init2=1.0, add_min2=0.001, add_max2=100.0,
    add_prop_var2=1e-6)


def tempest_survey_struct(rx_offset=(-107.0, 0.0, -45.0), additive_level=None):
    """The Tempest datapoint type; with `additive_level` (e.g. TEMPEST_ADDITIVE) also for the sampler."""
    return make_tdem_survey_struct([tempest_definition()], rx_offset, additive_level)


def tdem_primary_field(survey):
    """Primary field per system and measured component, in channel order (gbp_tdem_primary_field)."""
    out = np.zeros(2 * _lib.TD_MAXSYS)
    n = _lib.load().gbp_tdem_primary_field(ctypes.addressof(survey), out.ctypes.data)
    if n < 0:
        raise _lib.GeobipyB200Error(_lib.load().gbp_last_error().decode())
    return out[:n]


def skytem_survey_struct(rx_offset=(-13.0, 0.0, 2.0)):
    return make_tdem_survey_struct(skytem_definitions(), rx_offset)


def is_tdem(system):
    return isinstance(system, TdemSurveyC)


def n_channels(system):
    if is_tdem(system):
        return int(_lib.load().gbp_tdem_n_channels(ctypes.addressof(system)))
    return 2 * system.n_freq


def tdem_window_operator(survey):
    """(freq [32], MR [C, 32], MI [C, 32], t_centre [C]) - the model-independent tables of the time-domain path."""
    C = n_channels(survey)
    f, MR, MI, t = np.zeros(_lib.TD_NFREQ), np.zeros((C, _lib.TD_NFREQ)), np.zeros((C, _lib.TD_NFREQ)), np.zeros(C)
    _lib.check(_lib.load().gbp_tdem_window_operator(ctypes.addressof(survey), f.ctypes.data, MR.ctypes.data,
                                                    MI.ctypes.data, t.ctypes.data))
    return f, MR, MI, t


# skytem_options (documentation_source/source/supplementary/options_files/skytem_options); minimum_thickness None -> 1.0
# (RectilinearMesh1D.py:358)
SKYTEM_OPTIONS = dict(
    min_edge=1.0, max_edge=550.0, min_width=1.0, covariance_scaling=0.5, n_systems=2,
    rel_init=0.05, rel_min=0.005, rel_max=0.5, rel_prop_var=1e-6, add_init=2e-14, add_min=1e-16, add_max=1e-10,
    add_prop_var=1e-5, rel_init2=0.05, rel_min2=0.005, rel_max2=0.5, rel_prop_var2=1e-6, add_init2=2e-13,
    add_min2=1e-16, add_max2=1e-10, add_prop_var2=1e-5)

_OPTION_DEFAULTS = dict(
    n_markov_chains=100000, update_plot_every=5000, max_layers=30, solve_parameter=0, solve_gradient=1,
    solve_relative_error=1, solve_additive_error=1, reset_limit=1, min_edge=0.1, max_edge=200.0, min_width=1.0,
    p_birth=1.0 / 6.0, p_death=1.0 / 6.0, p_move=1.0 / 6.0, p_none=0.5, factor=10.0, gradient_std=1.5,
    covariance_scaling=1.0, rel_init=0.05, rel_min=0.001, rel_max=0.5, rel_prop_var=1e-6, add_init=5.0,
    add_min=3.0, add_max=20.0, add_prop_var=1e-6, n_sigma_bins=250, n_err_bins=99, sigma_bins_nstd=4.0,
    burn_in_min_iter=5000, n_systems=1,
    rel_init2=0.05, rel_min2=0.001, rel_max2=0.5, rel_prop_var2=1e-6, add_init2=5.0, add_min2=3.0, add_max2=20.0,
    add_prop_var2=1e-6, solve_height=0, max_height_change=1.0, height_prop_var=0.01)
_PER_SYSTEM = ("rel_init", "rel_min", "rel_max", "rel_prop_var", "add_init", "add_min", "add_max", "add_prop_var")

# reference option-file keys -> gbp_options fields (resolve_options; user_parameters.py:40-44)
_REFERENCE_KEYS = dict(
    n_markov_chains="n_markov_chains", update_plot_every="update_plot_every",
    maximum_number_of_layers="max_layers", solve_parameter="solve_parameter", solve_gradient="solve_gradient",
    solve_relative_error="solve_relative_error", solve_additive_error="solve_additive_error",
    reset_limit="reset_limit", minimum_depth="min_edge", maximum_depth="max_edge", minimum_thickness="min_width",
    probability_of_birth="p_birth", probability_of_death="p_death", probability_of_perturb="p_move",
    probability_of_no_change="p_none", factor="factor", gradient_standard_deviation="gradient_std",
    covariance_scaling="covariance_scaling", initial_relative_error="rel_init",
    minimum_relative_error="rel_min", maximum_relative_error="rel_max",
    relative_error_proposal_variance="rel_prop_var", initial_additive_error="add_init",
    minimum_additive_error="add_min", maximum_additive_error="add_max",
    additive_error_proposal_variance="add_prop_var",
    solve_z="solve_height", maximum_z_change="max_height_change", z_proposal_variance="height_prop_var",
    # time-domain datapoints: the sampled height is the transmitter loop's (EmLoop.set_priors via Loop_pair, option keys
    # prefixed "transmitter_"; TdemDataPoint.perturb :681-683)
    solve_transmitter_z="solve_height", maximum_transmitter_z_change="max_height_change",
    transmitter_z_proposal_variance="height_prop_var")


def make_options(**kw):
    """gbp_options.  Accepts gbp_options field names or the reference's option-file keys.

    Defaults = documentation_source/.../options_files/resolve_options with the ``None`` entries
    replaced as user_parameters.__init__ does (factor 10, gradient std 1.5, covariance scaling 1).
    """
    vals = dict(_OPTION_DEFAULTS)
    for k, v in kw.items():
        if v is None:
            continue
        k = _REFERENCE_KEYS.get(k, k)
        if k not in vals:
            continue  # keys of the reference options file that do not concern the sampler kernel
        if k in _PER_SYSTEM and np.size(v) > 1:  # per-system lists of a dual-moment options file (skytem_options)
            v = np.asarray(v, dtype=np.float64).reshape(-1)
            vals[k], vals[k + "2"] = float(v[0]), float(v[1])
            vals["n_systems"] = 2
            continue
        vals[k] = v
    o = OptionsC()
    for k, v in vals.items():
        if isinstance(getattr(o, k), int):
            v = int(v)
        else:
            v = float(np.asarray(v).reshape(-1)[0])
        setattr(o, k, v)
    # which unknown the height options were spelled for (Inference1D checks it against the datapoint type)
    o.height_key = "solve_transmitter_z" if kw.get("solve_transmitter_z") else ("solve_z" if kw.get("solve_z") else None)
    return o


# Keys the shipped option files carry (resolve_options :64, :113, :129; skytem_options :65) that the reference's classes
# never read: Point.set_priors / set_proposals (pointcloud/Point.py:959-961, :977-979) look for solve_z,
# maximum_z_change and z_proposal_variance, so `solve_height = True` in an options file samples nothing there.
_DEAD_REFERENCE_KEYS = ("solve_height", "maximum_height_change", "height_proposal_variance")
# unknowns of the reference that are not built here: asking for one is an error, not a silent no-op
_UNBUILT_PREFIXES = ("solve_transmitter_", "solve_receiver_")
_UNBUILT_KEYS = ("solve_x", "solve_y", "solve_calibration")
_BUILT_KEYS = ("solve_transmitter_z",)   # time-domain datapoints: the transmitter height (KIND_TDEM_Z)


def options_from_reference(**kw):
    """`make_options` for keyword arguments that come from a reference options file (`Inference1D(**options)`,
    `Inference3D.infer(**options)`): the keys the reference itself ignores are ignored here too (with a warning when
    they ask for something), and unknowns this library does not sample raise instead of being dropped."""
    import warnings
    kw = dict(kw)
    if kw.get("solve_height"):
        warnings.warn("solve_height / maximum_height_change / height_proposal_variance are not read by the reference "
                      "(Point.set_priors reads solve_z, maximum_z_change, z_proposal_variance): ignored, as there",
                      stacklevel=3)
    for k in _DEAD_REFERENCE_KEYS:
        kw.pop(k, None)
    for k, v in kw.items():
        if (k in _UNBUILT_KEYS or k.startswith(_UNBUILT_PREFIXES)) and k not in _BUILT_KEYS and np.any(v):
            raise NotImplementedError("%s: this unknown of the reference is not built in geobipy_b200 (DESIGN.md section 7)" % k)
    return make_options(**kw)


def n_depth(opt):
    return int(_lib.load().gbp_n_depth(ctypes.addressof(opt)))


def posterior_grids(opt, halfspace):
    """Bin edges of the posterior arrays, as the reference builds them.

    sigma edges: Model.set_posteriors (Model.py:666-684); depth edges: RectilinearMesh1D.set_posteriors
    (RectilinearMesh1D.py:1438-1455); error edges: DataPoint.set_*_error_posterior (DataPoint.py:668-695).
    """
    s = np.log(1.0 + opt.factor)
    t = opt.sigma_bins_nstd * s
    sig = np.exp(np.linspace(-t, t, opt.n_sigma_bins + 1) + np.log(halfspace))
    depth = np.arange(0.0, 1.1 * opt.max_edge, 0.5 * opt.min_width)
    rel = np.exp(np.linspace(np.log(opt.rel_min), np.log(opt.rel_max), opt.n_err_bins + 1))
    add = np.exp(np.linspace(np.log(opt.add_min), np.log(opt.add_max), opt.n_err_bins + 1))
    if opt.n_systems > 1:
        rel = np.stack([rel, np.exp(np.linspace(np.log(opt.rel_min2), np.log(opt.rel_max2), opt.n_err_bins + 1))])
        add = np.stack([add, np.exp(np.linspace(np.log(opt.add_min2), np.log(opt.add_max2), opt.n_err_bins + 1))])
    return dict(sigma_edges=sig, depth_edges=depth, rel_edges=rel, add_edges=add,
                ncells_centres=np.arange(0.0, opt.max_layers + 1.0))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def fdem_forward(system, nlayers, sigma, thickness, altitude, precision=PRECISION_F32, device=0, sensitivity=False):
    """Predicted data [B, 2F] (and Jacobian [B, 2F, Lstride] w.r.t. ln sigma) for B soundings.

    sigma / thickness are [B, Lstride] (Lstride <= 30); entries beyond nlayers[b] are ignored, the last
    layer is an infinite half-space.  numpy in -> numpy out (host path); torch CUDA in -> torch out.
    """
    lib = _lib.require_cuda()
    td = is_tdem(system)
    f_fwd, f_sens = (lib.gbp_tdem_forward, lib.gbp_tdem_sensitivity) if td else (lib.gbp_fdem_forward, lib.gbp_fdem_sensitivity)
    f_fwd_h, f_sens_h = ((lib.gbp_tdem_forward_host, lib.gbp_tdem_sensitivity_host) if td
                         else (lib.gbp_fdem_forward_host, lib.gbp_fdem_sensitivity_host))
    C = n_channels(system)
    if _is_torch(sigma):
        import torch
        assert sigma.is_cuda and sigma.dtype == torch.float64 and sigma.is_contiguous()
        B, ls = sigma.shape
        thickness = thickness.contiguous()
        altitude = altitude.contiguous()
        nlayers = nlayers.to(torch.int32).contiguous()
        out = torch.empty((B, C), dtype=torch.float64, device=sigma.device)
        st = torch.cuda.current_stream(sigma.device).cuda_stream
        with torch.cuda.device(sigma.device):
            if sensitivity:
                J = torch.empty((B, C, ls), dtype=torch.float64, device=sigma.device)
                _lib.check(f_sens(ctypes.addressof(system), B, ls, nlayers.data_ptr(), sigma.data_ptr(),
                                                    thickness.data_ptr(), altitude.data_ptr(), out.data_ptr(),
                                                    J.data_ptr(), precision, st))
                return out, J
            _lib.check(f_fwd(ctypes.addressof(system), B, ls, nlayers.data_ptr(), sigma.data_ptr(),
                                            thickness.data_ptr(), altitude.data_ptr(), out.data_ptr(), precision, st))
            return out
    sigma = _np(np.atleast_2d(sigma), np.float64)
    thickness = _np(np.atleast_2d(thickness), np.float64)
    B, ls = sigma.shape
    nlayers = _np(np.atleast_1d(nlayers), np.int32)
    altitude = _np(np.atleast_1d(altitude), np.float64)
    assert thickness.shape == sigma.shape and nlayers.shape == (B,) and altitude.shape == (B,)
    out = np.empty((B, C))
    if sensitivity:
        J = np.empty((B, C, ls))
        _lib.check(f_sens_h(ctypes.addressof(system), B, ls, nlayers.ctypes.data, sigma.ctypes.data,
                                                 thickness.ctypes.data, altitude.ctypes.data, out.ctypes.data,
                                                 J.ctypes.data, precision, device))
        return out, J
    _lib.check(f_fwd_h(ctypes.addressof(system), B, ls, nlayers.ctypes.data, sigma.ctypes.data,
                                         thickness.ctypes.data, altitude.ctypes.data, out.ctypes.data, precision, device))
    return out


tdem_forward = None  # bound below: same operator, dispatched on the system type


def chain_buffer_shapes(opt, B):
    nd = n_depth(opt)
    N2 = 2 * opt.n_markov_chains
    ml = opt.max_layers
    eshape = (B, 2, opt.n_err_bins) if opt.n_systems > 1 else (B, opt.n_err_bins)
    return dict(
        hitmap=((B, opt.n_sigma_bins, nd), np.int32), edges_hist=((B, nd), np.int32),
        ncells_hist=((B, ml + 1), np.int32), rel_hist=(eshape, np.int32),
        add_hist=(eshape, np.int32), misfit_trace=((B, N2), np.float64),
        accept_trace=((B, N2), np.uint8), best_sigma=((B, ml), np.float64), best_edges=((B, ml + 1), np.float64),
        cur_sigma=((B, ml), np.float64), cur_edges=((B, ml + 1), np.float64), scalars=((B, NSCALARS), np.float64),
        height_hist=((B, opt.n_err_bins), np.int32))


# height_hist (datapoint.z.posterior) joins the default outputs only when the options sample the height (solve_z)
DEFAULT_OUTPUTS = tuple(f for f in BUFFER_FIELDS if f != "height_hist")


def rjmcmc_run(system, opt, data, altitude, seed=0, first_index=0, max_iterations=0, precision=PRECISION_F32,
               device=0, outputs=DEFAULT_OUTPUTS, buffers=None):
    """Run B independent rjMCMC chains (one per sounding) on the GPU.

    data [B, 2F] observed ppm, altitude [B].  numpy in -> dict of numpy arrays (host path, copies inside the
    C call); torch CUDA in -> dict of torch tensors left on the device (``buffers`` may supply pre-zeroed
    result tensors to reuse).
    """
    lib = _lib.require_cuda()
    td = is_tdem(system)
    f_run, f_run_h = (lib.gbp_tdem_rjmcmc_run, lib.gbp_tdem_rjmcmc_run_host) if td else (lib.gbp_rjmcmc_run, lib.gbp_rjmcmc_run_host)
    if opt.solve_height and outputs is DEFAULT_OUTPUTS:
        outputs = DEFAULT_OUTPUTS + ("height_hist",)
    outputs = tuple(outputs)
    if "scalars" not in outputs:
        outputs = outputs + ("scalars",)
    if _is_torch(data):
        import torch
        assert data.is_cuda and data.dtype == torch.float64
        data = data.contiguous()
        altitude = altitude.contiguous()
        B = data.shape[0]
        shapes = chain_buffer_shapes(opt, B)
        tdt = {np.int32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}
        res = {}
        cb = ChainBuffersC()
        for name in outputs:
            shp, dt = shapes[name]
            if buffers is not None and name in buffers:
                t = buffers[name]
                t.zero_()
            else:
                t = torch.zeros(shp, dtype=tdt[dt], device=data.device)
            res[name] = t
            setattr(cb, name, t.data_ptr())
        st = torch.cuda.current_stream(data.device).cuda_stream
        with torch.cuda.device(data.device):
            _lib.check(f_run(ctypes.addressof(system), ctypes.addressof(opt), B, data.data_ptr(),
                                          altitude.data_ptr(), int(seed), int(first_index), int(max_iterations),
                                          ctypes.addressof(cb), precision, st))
        return res
    data = _np(np.atleast_2d(data), np.float64)
    altitude = _np(np.atleast_1d(altitude), np.float64)
    B = data.shape[0]
    assert data.shape[1] == n_channels(system) and altitude.shape == (B,)
    shapes = chain_buffer_shapes(opt, B)
    res = {}
    cb = ChainBuffersC()
    for name in outputs:
        shp, dt = shapes[name]
        a = buffers[name] if (buffers is not None and name in buffers) else np.zeros(shp, dtype=dt)
        res[name] = a
        setattr(cb, name, a.ctypes.data)
    _lib.check(f_run_h(ctypes.addressof(system), ctypes.addressof(opt), B, data.ctypes.data,
                                       altitude.ctypes.data, int(seed), int(first_index), int(max_iterations),
                                       ctypes.addressof(cb), precision, device))
    return res


def last_kernel_ms():
    ms = ctypes.c_float(0.0)
    _lib.check(_lib.load().gbp_last_kernel_ms(ctypes.byref(ms)))
    return float(ms.value)


def kernel_ms_stats(last_n):
    """(mean duration [ms], number averaged) of the last `last_n` (<= 32) kernels launched on the current device."""
    ms, n = ctypes.c_float(0.0), ctypes.c_int(0)
    _lib.check(_lib.load().gbp_kernel_ms_stats(int(last_n), ctypes.byref(ms), ctypes.byref(n)))
    return float(ms.value), int(n.value)


def launch_count():
    return int(_lib.load().gbp_launch_count())


def flops_per_forward(system, n_layers):
    if is_tdem(system):
        return float(_lib.load().gbp_tdem_flops_per_forward(ctypes.addressof(system), int(n_layers)))
    return float(_lib.load().gbp_flops_per_forward(ctypes.addressof(system), int(n_layers)))


def filter_points(system):
    return int(_lib.load().gbp_filter_points(ctypes.addressof(system)))


def forward(system, nlayers, sigma, thickness, altitude, **kw):
    """Forward / Jacobian operator of either datapoint type (FDEM system struct or time-domain survey struct)."""
    return fdem_forward(system, nlayers, sigma, thickness, altitude, **kw)


tdem_forward = forward


def mufu_per_forward(system, n_layers):
    """Special-function-unit operations of one forward (rcp, sqrt, ex2, sin, cos)."""
    lib = _lib.load()
    f = lib.gbp_tdem_mufu_per_forward if is_tdem(system) else lib.gbp_mufu_per_forward
    return float(f(ctypes.addressof(system), int(n_layers)))


def measure_peaks():
    """(fp32 TFLOP/s, MUFU Gop/s) measured on the current device with two microbenchmark kernels."""
    lib = _lib.require_cuda()
    a, b = ctypes.c_double(0.0), ctypes.c_double(0.0)
    _lib.check(lib.gbp_measure_peaks(ctypes.byref(a), ctypes.byref(b)))
    return float(a.value), float(b.value)


def debug_counters(reset=False):
    """The 16 speculation diagnostics of include/geobipy_b200.h gbp_debug_counters (numpy uint64)."""
    out = np.zeros(16, dtype=np.uint64)
    _lib.check(_lib.load().gbp_debug_counters(out.ctypes.data, 1 if reset else 0))
    return out


def release_host_buffers():
    """Free the device buffers the numpy (host-pointer) path of rjmcmc_run keeps between calls."""
    _lib.check(_lib.load().gbp_release_host_buffers())


def summarise_hitmap(hitmap, sig_lo, dx, percentiles=(5.0, 50.0, 95.0)):
    """Per-depth-cell mean and percentiles of ln(sigma) of hitmaps [B, n_sig, n_depth] (torch CUDA int32) with the
    hand-written kernel behind gbp_summarise_hitmap.  sig_lo [B] = ln(sigma) at the lower edge of bin 0 (torch CUDA
    float64), dx = bin width in ln(sigma).  Returns (mean [B, n_depth], pct [n_pct, B, n_depth])."""
    import torch
    lib = _lib.require_cuda()
    assert hitmap.is_cuda and hitmap.dtype == torch.int32 and hitmap.is_contiguous()
    B, ns, nd = hitmap.shape
    sig_lo = sig_lo.to(torch.float64).contiguous()
    p = np.ascontiguousarray(percentiles, dtype=np.float64)
    mean = torch.empty((B, nd), dtype=torch.float64, device=hitmap.device)
    pct = torch.empty((p.size, B, nd), dtype=torch.float64, device=hitmap.device)
    st = torch.cuda.current_stream(hitmap.device).cuda_stream
    with torch.cuda.device(hitmap.device):
        _lib.check(lib.gbp_summarise_hitmap(hitmap.data_ptr(), B, ns, nd, sig_lo.data_ptr(), float(dx), p.ctypes.data, int(p.size),
                                            mean.data_ptr(), pct.data_ptr(), st))
    return mean, pct


def summarise_posterior(hitmap, sig_lo, dx, percentiles=(5.0, 50.0, 95.0), credible_percent=90.0):
    """All per-depth-cell summaries of hitmaps [B, n_sig, n_depth] (torch CUDA int32) in one pass of the hand-written
    kernel (gbp_summarise_posterior): mean, percentiles and mode of ln(sigma), and the credible range of
    `credible_percent` in bins.  Returns dict(mean [B, nd], pct [n_pct, B, nd], mode [B, nd], range_bins [B, nd] int32,
    credible_range [B, nd] in decades = Mesh._credible_range with the hitmap's log = 10 axis)."""
    import torch
    lib = _lib.require_cuda()
    assert hitmap.is_cuda and hitmap.dtype == torch.int32 and hitmap.is_contiguous()
    B, ns, nd = hitmap.shape
    sig_lo = sig_lo.to(torch.float64).contiguous()
    p = np.ascontiguousarray(percentiles, dtype=np.float64)
    dev = hitmap.device
    mean = torch.empty((B, nd), dtype=torch.float64, device=dev)
    pct = torch.empty((p.size, B, nd), dtype=torch.float64, device=dev)
    mode = torch.empty((B, nd), dtype=torch.float64, device=dev)
    rng = torch.empty((B, nd), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.gbp_summarise_posterior(hitmap.data_ptr(), B, ns, nd, sig_lo.data_ptr(), float(dx), p.ctypes.data, int(p.size),
                                               float(credible_percent), mean.data_ptr(), pct.data_ptr(), mode.data_ptr(),
                                               rng.data_ptr(), st))
    return dict(mean=mean, pct=pct, mode=mode, range_bins=rng, credible_range=rng.to(torch.float64) * (float(dx) / np.log(10.0)))


def opacity_doi(range_bins, group=None, n_groups=0, doi_percent=67.0, level_percent=95.0):
    """Opacity [B, nd], depth-of-investigation cell [B] and opacity-level cell [B] from credible ranges in bins (torch CUDA
    int32 [B, nd]) with gbp_opacity_doi.  group [B] int32 (torch CUDA) = the flight line of each sounding: the range is
    normalised over the line (Inference2D.compute_opacity); None = per sounding (Histogram.opacity)."""
    import torch
    lib = _lib.require_cuda()
    assert range_bins.is_cuda and range_bins.dtype == torch.int32 and range_bins.is_contiguous()
    B, nd = range_bins.shape
    dev = range_bins.device
    ng = int(n_groups) if group is not None else B
    if group is not None:
        group = group.to(torch.int32).contiguous()
    scratch = torch.empty((2, max(ng, 1)), dtype=torch.int32, device=dev)
    opacity = torch.empty((B, nd), dtype=torch.float64, device=dev)
    doi = torch.empty((B,), dtype=torch.int32, device=dev)
    level = torch.empty((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.gbp_opacity_doi(range_bins.data_ptr(), B, nd, group.data_ptr() if group is not None else None, ng,
                                       float(doi_percent), float(level_percent), scratch.data_ptr(), opacity.data_ptr(),
                                       doi.data_ptr(), level.data_ptr(), st))
    return opacity, doi, level
