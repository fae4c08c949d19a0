"""TEST INFRASTRUCTURE ONLY - ctypes front end of the C oracle (oracle/*.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (geobipy_b200) never imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgbp_oracle.so")

MAXL, MAXF = 64, 16
MAXC = 64
MAXSYS, TD_NFREQ, TD_MAXLAM = 2, 32, 32
NSCALARS = 32

S_ITER, S_BURNED_IN, S_BURNED_IN_ITER, S_BEST_ITER, S_BEST_K, S_CUR_K, S_HALFSPACE, S_FAILED, S_N_ACCEPT, \
    S_N_FORWARD, S_N_SENS, S_BEST_POSTERIOR, S_CUR_REL, S_CUR_ADD, S_CUR_MISFIT, S_CUR_PRIOR, S_CUR_LIKELIHOOD, \
    S_BEST_REL, S_BEST_ADD, S_N_RESETS, S_N_BIRTH, S_N_DEATH, S_N_MOVE, S_N_NONE, S_TOTAL_ITER, \
    S_CUR_REL2, S_CUR_ADD2, S_BEST_REL2, S_BEST_ADD2, S_CUR_HEIGHT, S_BEST_HEIGHT, S_HEIGHT_REF = range(32)


class FdemSystemC(ctypes.Structure):
    _fields_ = [("n_freq", ctypes.c_int32), ("tid", ctypes.c_int32 * MAXF)] + [
        (n, ctypes.c_double * MAXF) for n in ("freq", "tmom", "tx", "ty", "tz", "rmom", "rx", "ry", "rz")]


class OptionsC(ctypes.Structure):
    _fields_ = [
        ("n_markov_chains", ctypes.c_int32), ("update_plot_every", ctypes.c_int32), ("max_layers", ctypes.c_int32),
        ("solve_parameter", ctypes.c_int32), ("solve_gradient", ctypes.c_int32),
        ("solve_relative_error", ctypes.c_int32), ("solve_additive_error", ctypes.c_int32),
        ("reset_limit", ctypes.c_int32),
        ("min_edge", ctypes.c_double), ("max_edge", ctypes.c_double), ("min_width", ctypes.c_double),
        ("p_birth", ctypes.c_double), ("p_death", ctypes.c_double), ("p_move", ctypes.c_double),
        ("p_none", ctypes.c_double),
        ("factor", ctypes.c_double), ("gradient_std", ctypes.c_double), ("covariance_scaling", ctypes.c_double),
        ("rel_init", ctypes.c_double), ("rel_min", ctypes.c_double), ("rel_max", ctypes.c_double),
        ("rel_prop_var", ctypes.c_double),
        ("add_init", ctypes.c_double), ("add_min", ctypes.c_double), ("add_max", ctypes.c_double),
        ("add_prop_var", ctypes.c_double),
        ("n_sigma_bins", ctypes.c_int32), ("n_err_bins", ctypes.c_int32), ("sigma_bins_nstd", ctypes.c_double),
        ("burn_in_min_iter", ctypes.c_int32), ("n_systems", ctypes.c_int32),
        ("rel_init2", ctypes.c_double), ("rel_min2", ctypes.c_double), ("rel_max2", ctypes.c_double),
        ("rel_prop_var2", ctypes.c_double),
        ("add_init2", ctypes.c_double), ("add_min2", ctypes.c_double), ("add_max2", ctypes.c_double),
        ("add_prop_var2", ctypes.c_double),
        ("solve_height", ctypes.c_int32), ("pad_h_", ctypes.c_int32),
        ("max_height_change", ctypes.c_double), ("height_prop_var", ctypes.c_double),
    ]


class TdemSystemC(ctypes.Structure):
    """gbo_tdem_system"""
    _fields_ = [
        ("n_sys", ctypes.c_int32), ("n_freq", ctypes.c_int32), ("n_lam", ctypes.c_int32), ("C", ctypes.c_int32),
        ("n_win", ctypes.c_int32 * MAXSYS), ("pad_", ctypes.c_int32 * 2),
        ("freq", ctypes.c_double * TD_NFREQ), ("xi", ctypes.c_double * TD_MAXLAM),
        ("rx_dx", ctypes.c_double), ("rx_dy", ctypes.c_double), ("rx_dz", ctypes.c_double),
        ("loop_radius", ctypes.c_double),
        ("MR", ctypes.c_double * (MAXC * TD_NFREQ)), ("MI", ctypes.c_double * (MAXC * TD_NFREQ)),
        ("t_centre", ctypes.c_double * MAXC),
        ("comp", ctypes.c_int32 * MAXC), ("rx_cx", ctypes.c_double),
        ("tempest", ctypes.c_int32), ("pad_t", ctypes.c_int32),
        ("add_level", ctypes.c_double * MAXC), ("primary", ctypes.c_double * MAXC),
    ]


class ChainOutC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "misfit_trace", "accept_trace",
        "best_sigma", "best_edges", "cur_sigma", "cur_edges", "scalars", "height_hist")]


class TransitionC(ctypes.Structure):
    _fields_ = [
        ("k", ctypes.c_int32), ("action", ctypes.c_int32), ("altitude", ctypes.c_double),
        ("sigma_ref", ctypes.c_double),
        ("edges", ctypes.c_double * (MAXL + 1)), ("sigma_remap", ctypes.c_double * MAXL),
        ("sigma_test", ctypes.c_double * MAXL),
        ("rel_cur", ctypes.c_double * MAXSYS), ("add_cur", ctypes.c_double * MAXSYS),
        ("rel_test", ctypes.c_double * MAXSYS), ("add_test", ctypes.c_double * MAXSYS),
        ("data", ctypes.c_double * MAXC), ("J_in", ctypes.c_double * (MAXC * MAXL)),
        ("pred_in", ctypes.c_double * MAXC),
        ("hessian", ctypes.c_double * (MAXL * MAXL)), ("gradient", ctypes.c_double * MAXL),
        ("newton_mean", ctypes.c_double * MAXL), ("pred_test", ctypes.c_double * MAXC),
        ("misfit_test", ctypes.c_double), ("prior_test", ctypes.c_double), ("likelihood_test", ctypes.c_double),
        ("proposal", ctypes.c_double), ("proposal1", ctypes.c_double),
        ("altitude_test", ctypes.c_double), ("altitude_ref", ctypes.c_double),
    ]


def build(force=False):
    """Compile the oracle with gcc (a few seconds)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in ("fdem1d_oracle.c", "tdem1d_oracle.c", "rjmcmc_oracle.c", "oracle.h")):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.gbo_fdem_forward.restype = ctypes.c_int
        _lib.gbo_fdem_sensitivity.restype = ctypes.c_int
        _lib.gbo_run_chain.restype = ctypes.c_int
        _lib.gbo_run_chain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                       ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p]
        _lib.gbo_n_depth.restype = ctypes.c_int
        _lib.gbo_run_chain_tdem.restype = ctypes.c_int
        _lib.gbo_run_chain_tdem.argtypes = _lib.gbo_run_chain.argtypes
        for n in ("gbo_tdem_forward", "gbo_tdem_sensitivity", "gbo_tdem_frequency_response"):
            getattr(_lib, n).restype = ctypes.c_int
    return _lib


# RESOLVE system (documentation_source/source/supplementary/data/resolve.stm)
RESOLVE = dict(
    freq=[380.0, 1776.0, 3345.0, 8171.0, 41020.0, 129550.0],
    tor=["z", "z", "x", "z", "z", "z"], tmom=[1, 1, -1, 1, 1, 1], tx=[0] * 6, ty=[0] * 6, tz=[0] * 6,
    ror=["z", "z", "x", "z", "z", "z"], rmom=[1, 1, 1, 1, 1, 1], rx=[7.93, 7.91, 9.03, 7.91, 7.91, 7.89],
    ry=[0] * 6, rz=[0] * 6)

_ORI = {"x": 0, "y": 1, "z": 2}


def make_system(d=None):
    d = RESOLVE if d is None else d
    s = FdemSystemC()
    n = len(d["freq"])
    s.n_freq = n
    for i in range(n):
        s.tid[i] = 1 + 3 * _ORI[d["ror"][i]] + _ORI[d["tor"][i]]
        for f in ("freq", "tmom", "tx", "ty", "tz", "rmom", "rx", "ry", "rz"):
            getattr(s, f)[i] = float(d[f][i])
    return s


def resolve_options(**over):
    """resolve_options + user_parameters defaults (user_parameters.py:40-44)."""
    o = OptionsC()
    o.n_markov_chains = 100000
    o.update_plot_every = 5000
    o.max_layers = 30
    o.solve_parameter = 0
    o.solve_gradient = 1
    o.solve_relative_error = 1
    o.solve_additive_error = 1
    o.reset_limit = 1
    o.min_edge, o.max_edge, o.min_width = 0.1, 200.0, 1.0
    o.p_birth = o.p_death = o.p_move = 1.0 / 6.0
    o.p_none = 0.5
    o.factor, o.gradient_std, o.covariance_scaling = 10.0, 1.5, 1.0
    o.rel_init, o.rel_min, o.rel_max, o.rel_prop_var = 0.05, 0.001, 0.5, 1e-6
    o.add_init, o.add_min, o.add_max, o.add_prop_var = 5.0, 3.0, 20.0, 1e-6
    o.n_sigma_bins, o.n_err_bins, o.sigma_bins_nstd = 250, 99, 4.0
    o.burn_in_min_iter = 5000
    o.solve_height, o.max_height_change, o.height_prop_var = 0, 1.0, 0.01   # solve_z is off in the shipped options
    for k, v in over.items():
        setattr(o, k, v)
    return o


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fdem_forward(sys, altitude, sigma, thickness):
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    thickness = np.ascontiguousarray(thickness, dtype=np.float64)
    out = np.zeros(2 * sys.n_freq)
    rc = lib().gbo_fdem_forward(ctypes.byref(sys), ctypes.c_double(altitude), ctypes.c_int(sigma.size), _p(sigma),
                                _p(thickness), _p(out))
    assert rc == 0
    return out


def fdem_sensitivity(sys, altitude, sigma, thickness):
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    thickness = np.ascontiguousarray(thickness, dtype=np.float64)
    J = np.zeros((2 * sys.n_freq, sigma.size))
    rc = lib().gbo_fdem_sensitivity(ctypes.byref(sys), ctypes.c_double(altitude), ctypes.c_int(sigma.size),
                                    _p(sigma), _p(thickness), _p(J))
    assert rc == 0
    return J


def philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().gbo_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def run_chain(sys, opt, data, altitude, seed, sounding_index, max_iterations=0):
    """Run one chain (FDEM system or TdemSystemC); returns a dict of numpy arrays."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    nd = lib().gbo_n_depth(ctypes.byref(opt))
    N2 = 2 * opt.n_markov_chains
    tdem = isinstance(sys, TdemSystemC)
    eshape = (max(int(opt.n_systems), 1), opt.n_err_bins) if tdem else (opt.n_err_bins,)   # (Tempest: one entry per component)
    r = dict(
        hitmap=np.zeros((opt.n_sigma_bins, nd), np.int32), edges_hist=np.zeros(nd, np.int32),
        ncells_hist=np.zeros(opt.max_layers + 1, np.int32), rel_hist=np.zeros(eshape, np.int32),
        add_hist=np.zeros(eshape, np.int32), misfit_trace=np.zeros(N2), accept_trace=np.zeros(N2, np.uint8),
        best_sigma=np.zeros(opt.max_layers), best_edges=np.zeros(opt.max_layers + 1),
        cur_sigma=np.zeros(opt.max_layers), cur_edges=np.zeros(opt.max_layers + 1), scalars=np.zeros(NSCALARS),
        height_hist=np.zeros(opt.n_err_bins, np.int32))
    co = ChainOutC()
    for k, v in r.items():
        setattr(co, k, v.ctypes.data)
    fn = lib().gbo_run_chain_tdem if tdem else lib().gbo_run_chain
    rc = fn(ctypes.addressof(sys), ctypes.addressof(opt), data.ctypes.data, float(altitude),
            int(seed), int(sounding_index), int(max_iterations), ctypes.addressof(co))
    assert rc == 0, rc
    return r


def eval_transition(sys, opt, **kw):
    t = TransitionC()
    k = int(kw["k"])
    t.k, t.action, t.altitude, t.sigma_ref = k, int(kw["action"]), float(kw["altitude"]), float(kw["sigma_ref"])
    t.altitude_test = float(kw.get("altitude_test", kw["altitude"]))
    t.altitude_ref = float(kw.get("altitude_ref", kw["altitude"]))
    tdem = isinstance(sys, TdemSystemC)
    C = sys.C if tdem else 2 * sys.n_freq
    for i in range(k + 1):
        t.edges[i] = float(kw["edges"][i])
    for i in range(k):
        t.sigma_remap[i] = float(kw["sigma_remap"][i])
        t.sigma_test[i] = float(kw["sigma_test"][i])
    for n in ("rel_cur", "add_cur", "rel_test", "add_test"):
        for i, v in enumerate(np.atleast_1d(np.asarray(kw[n], dtype=np.float64))):
            getattr(t, n)[i] = float(v)
    for i in range(C):
        t.data[i] = float(kw["data"][i])
        t.pred_in[i] = float(kw["pred_in"][i])
    Jin = np.asarray(kw["J_in"], dtype=np.float64).reshape(C, -1)
    if Jin.shape[1] == k:
        for c in range(C):
            for i in range(k):
                t.J_in[c * k + i] = Jin[c, i]
    fn = lib().gbo_eval_transition_tdem if tdem else lib().gbo_eval_transition
    rc = fn(ctypes.byref(sys), ctypes.byref(opt), ctypes.byref(t))
    return rc, dict(
        hessian=np.array(t.hessian[:k * k]).reshape(k, k), gradient=np.array(t.gradient[:k]),
        newton_mean=np.array(t.newton_mean[:k]), pred_test=np.array(t.pred_test[:C]),
        misfit_test=t.misfit_test, prior_test=t.prior_test, likelihood_test=t.likelihood_test,
        proposal=t.proposal, proposal1=t.proposal1)


# ---------------------------------------------------------------------------------- time domain (SkyTEM)
# Instrument descriptions = the reference's .stm files, parsed into JSON (geobipy_b200/data/*.json hold the
# numbers of documentation_source/source/supplementary/data/SkytemHM.stm / SkytemLM.stm).
def skytem_definitions():
    import json
    d = os.path.join(HERE, "..", "geobipy_b200", "data")
    return [json.load(open(os.path.join(d, n))) for n in ("skytem_hm.json", "skytem_lm.json")]


TD_XI_LO, TD_XI_HI = -4.6, 2.3   # ln(lambda ZH / 2) range of the Hankel trapezoid rule


def tdem_nodes(defs):
    """Spline-node frequencies shared by the systems: TD_NFREQ log-spaced values from the lowest base
    frequency to the highest digitising Nyquist frequency."""
    lo = min(d["base_frequency"] for d in defs)
    hi = max(0.5 * d["digitising_frequency"] for d in defs)
    return lo * (hi / lo) ** (np.arange(TD_NFREQ) / (TD_NFREQ - 1.0))


def tdem_components(d):
    """Components a system measures, in the reference's channel order (TdemDataPoint.forward :1008-1016: x, then z):
    those with a non-zero Output Scaling in the .stm (SkyTEM: z; Tempest: x and z)."""
    return [c for c, k in (("x", "x_scaling"), ("z", "z_scaling")) if float(d.get(k, 1.0 if c == "z" else 0.0)) != 0.0]


def tdem_window_operator(d, fnodes, component="z"):
    """MR, MI [n_windows, n_nodes]: window averages of the measured quantity from the node values of S(f) (see
    tdem1d_oracle.c step 3).  Independent numpy/scipy statement of what the product builds in C++.

    The physical time signal of harmonic coefficients A is 2 Re(A S e^{iwt}) -> (2 Re A, -2 Im A).  OutputType dB/dt: the
    data are receiver voltages, -dB/dt (SkyTEM goldens); OutputType B: A / (i w) and the field itself (Tempest goldens).
    The x component comes with the opposite sign of z (Tempest goldens; unpinned for a dB/dt system).  PeakCurrent scales
    the normalised waveform, X / ZOutputScaling the output (Tempest: 0.5 A, 1e15 = fT)."""
    from scipy.interpolate import CubicSpline
    base, nyq = d["base_frequency"], 0.5 * d["digitising_frequency"]
    T = 1.0 / base
    f = np.arange(1, int(nyq / base) + 1, 2) * base          # odd harmonics of the bipolar waveform
    w = 2.0 * np.pi * f
    wt, wa = np.asarray(d["waveform_time"]), np.asarray(d["waveform_current"])
    slope = np.diff(wa) / np.diff(wt)
    dn = np.zeros(f.size, complex)                            # Fourier coefficients of dI/dt
    for j, s in enumerate(slope):
        if s != 0.0:
            dn += s * (np.exp(-1j * w * wt[j]) - np.exp(-1j * w * wt[j + 1])) / (1j * w)
    full = abs((wt[-1] - wt[0]) * base - 1.0) < 1e-3          # the .stm gives the whole period (Tempest) or half of it
    dn *= (1.0 if full else 2.0) / T                          # (second half period = minus the first)
    F = np.ones(f.size, complex)
    for fc, order in zip(d["filter_cutoff"], d["filter_order"]):
        F *= (1.0 / (1.0 + 1j * f / fc)) ** order             # cascaded first-order stages
    ta, tb = np.asarray(d["window_start"])[:, None], np.asarray(d["window_end"])[:, None]
    A = dn * F * (np.exp(1j * w * tb) - np.exp(1j * w * ta)) / (1j * w * (tb - ta))
    b_field = str(d.get("output_type", "dB/dt")).strip().lower() == "b"
    if b_field:
        A = A / (1j * w)
    sign = (1.0 if b_field else -1.0) * (1.0 if component == "z" else -1.0)
    scale = sign * float(d.get("peak_current", 1.0)) * float(d.get("z_scaling" if component == "z" else "x_scaling", 1.0))
    lf = np.log10(fnodes)
    G = np.stack([CubicSpline(lf, e)(np.log10(f)) for e in np.eye(fnodes.size)], axis=1)  # harmonics x nodes
    return scale * 2.0 * A.real @ G, -scale * 2.0 * A.imag @ G


def tdem_primary_field(d, rx_offset):
    """Primary field [per component, reference order] of the transmitter dipole at the receiver during the windows
    (Tempest: PX, PZ; TdemDataPoint.forward :1008-1016 stacks PX, -PZ): B = mu0 m / (4 pi) (3 x z / R^5, 3 z^2 / R^5 - 1 / R^3),
    m = PeakCurrent x NumberOfTurns x LoopArea, z sign flipped, times the Output Scaling."""
    x, y, z = rx_offset
    R = np.sqrt(x * x + y * y + z * z)
    m = float(d.get("peak_current", 1.0)) * float(d.get("n_turns", 1.0)) * float(d.get("loop_area", 1.0)) * 1e-7
    out = []
    for c in tdem_components(d):
        if c == "x":
            out.append(m * 3.0 * x * z / R ** 5 * float(d.get("x_scaling", 1.0)))
        else:
            out.append(-m * (3.0 * z * z / R ** 5 - 1.0 / R ** 3) * float(d.get("z_scaling", 1.0)))
    return np.asarray(out)


def make_tdem_system(defs=None, rx_offset=(-13.0, 0.0, 2.0)):
    defs = skytem_definitions() if defs is None else defs
    s = TdemSystemC()
    fn = tdem_nodes(defs)
    n_lam = 2 * ((max(d["n_abscissae"] for d in defs) + 1) // 2)
    s.n_sys, s.n_freq, s.n_lam = len(defs), TD_NFREQ, n_lam
    for i, v in enumerate(fn):
        s.freq[i] = v
    for i, v in enumerate(np.linspace(TD_XI_LO, TD_XI_HI, n_lam)):
        s.xi[i] = v
    s.rx_dx, s.rx_dy, s.rx_dz = rx_offset
    r = float(np.hypot(rx_offset[0], rx_offset[1]))
    s.rx_cx = rx_offset[0] / r if r > 0.0 else 0.0
    s.loop_radius = defs[0]["loop_radius"]
    c = 0
    for k, d in enumerate(defs):
        s.n_win[k] = len(d["window_start"])
        for comp in tdem_components(d):        # channel order: system, then component (x, z), then window
            MR, MI = tdem_window_operator(d, fn, comp)
            for i in range(MR.shape[0]):
                for j in range(TD_NFREQ):
                    s.MR[c * TD_NFREQ + j] = MR[i, j]
                    s.MI[c * TD_NFREQ + j] = MI[i, j]
                s.t_centre[c] = 0.5 * (d["window_start"][i] + d["window_end"][i])
                s.comp[c] = 1 if comp == "x" else 0
                c += 1
    s.C = c
    return s


TEMPEST_ADDITIVE = np.r_[0.011474, 0.012810, 0.008507, 0.005154, 0.004742, 0.004477, 0.004168, 0.003539, 0.003352, 0.003213, 0.003161,
                         0.003122, 0.002587, 0.002038, 0.002201, 0.007383, 0.005693, 0.005178, 0.003659, 0.003426, 0.003046, 0.003095,
                         0.003247, 0.002775, 0.002627, 0.002460, 0.002178, 0.001754, 0.001405, 0.001283]   # tempest_options: initial_additive_error [fT]


def tempest_definition():
    import json
    return json.load(open(os.path.join(HERE, "..", "geobipy_b200", "data", "tempest.json")))


def make_tempest_system(d=None, rx_offset=(-107.0, 0.0, -45.0), additive_level=TEMPEST_ADDITIVE):
    """A Tempest datapoint type for the SAMPLER oracle: the forward tables of make_tdem_system plus the fixed additive level
    of every channel and the predicted primary field of its component (Tempest_datapoint.py:107-127, :141-176)."""
    d = tempest_definition() if d is None else d
    s = make_tdem_system([d], rx_offset=rx_offset)
    s.tempest = 1
    comps = tdem_components(d)
    prim = tdem_primary_field(d, rx_offset)
    nw = len(d["window_start"])
    for q in range(len(comps)):
        for i in range(nw):
            s.primary[q * nw + i] = prim[q]
            s.add_level[q * nw + i] = float(additive_level[q * nw + i])
    return s


def tempest_options(**over):
    """tempest_options (documentation_source/source/supplementary/options_files/tempest_options): errors per component (x,
    z); the "additive error" unknown is the multiplier of the fixed additive levels (initially 1, prior bounds
    minimum / maximum_additive_error)."""
    o = resolve_options()
    o.min_edge, o.max_edge, o.min_width = 1.0, 550.0, 1.0   # minimum_thickness None -> 1.0 (RectilinearMesh1D.py:358)
    o.covariance_scaling, o.gradient_std = 0.5, 5.0
    o.n_markov_chains = 1000
    o.n_systems = 2
    o.rel_init, o.rel_min, o.rel_max, o.rel_prop_var = 0.001, 0.0001, 0.01, 1e-6
    o.rel_init2, o.rel_min2, o.rel_max2, o.rel_prop_var2 = 0.001, 0.0001, 0.01, 1e-6
    o.add_init, o.add_min, o.add_max, o.add_prop_var = 1.0, 0.001, 100.0, 1e-6
    o.add_init2, o.add_min2, o.add_max2, o.add_prop_var2 = 1.0, 0.001, 100.0, 1e-6
    for k, v in over.items():
        setattr(o, k, v)
    return o


def tdem_forward(sys, altitude, sigma, thickness):
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    thickness = np.ascontiguousarray(thickness, dtype=np.float64)
    out = np.zeros(sys.C)
    rc = lib().gbo_tdem_forward(ctypes.byref(sys), ctypes.c_double(altitude), ctypes.c_int(sigma.size), _p(sigma),
                                _p(thickness), _p(out))
    assert rc == 0
    return out


def tdem_sensitivity(sys, altitude, sigma, thickness):
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    thickness = np.ascontiguousarray(thickness, dtype=np.float64)
    J = np.zeros((sys.C, sigma.size))
    rc = lib().gbo_tdem_sensitivity(ctypes.byref(sys), ctypes.c_double(altitude), ctypes.c_int(sigma.size),
                                    _p(sigma), _p(thickness), _p(J))
    assert rc == 0
    return J


def skytem_options(**over):
    """skytem_options (documentation_source/source/supplementary/options_files/skytem_options)."""
    o = resolve_options()
    o.min_edge, o.max_edge, o.min_width = 1.0, 550.0, 1.0   # minimum_thickness None -> 1.0 (RectilinearMesh1D.py:358)
    o.covariance_scaling = 0.5
    o.n_systems = 2
    o.rel_init, o.rel_min, o.rel_max, o.rel_prop_var = 0.05, 0.005, 0.5, 1e-6
    o.rel_init2, o.rel_min2, o.rel_max2, o.rel_prop_var2 = 0.05, 0.005, 0.5, 1e-6
    o.add_init, o.add_min, o.add_max, o.add_prop_var = 2e-14, 1e-16, 1e-10, 1e-5
    o.add_init2, o.add_min2, o.add_max2, o.add_prop_var2 = 2e-13, 1e-16, 1e-10, 1e-5
    for k, v in over.items():
        setattr(o, k, v)
    return o
