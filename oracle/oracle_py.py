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
MAXC = 2 * MAXF
NSCALARS = 32

S_ITER, S_BURNED_IN, S_BURNED_IN_ITER, S_BEST_ITER, S_BEST_K, S_CUR_K, S_HALFSPACE, S_FAILED, S_N_ACCEPT, \
    S_N_FORWARD, S_N_SENS, S_BEST_POSTERIOR, S_CUR_REL, S_CUR_ADD, S_CUR_MISFIT, S_CUR_PRIOR, S_CUR_LIKELIHOOD, \
    S_BEST_REL, S_BEST_ADD, S_N_RESETS, S_N_BIRTH, S_N_DEATH, S_N_MOVE, S_N_NONE, S_TOTAL_ITER = range(25)


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
        ("burn_in_min_iter", ctypes.c_int32), ("pad_", ctypes.c_int32),
    ]


class ChainOutC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "misfit_trace", "accept_trace",
        "best_sigma", "best_edges", "cur_sigma", "cur_edges", "scalars")]


class TransitionC(ctypes.Structure):
    _fields_ = [
        ("k", ctypes.c_int32), ("action", ctypes.c_int32), ("altitude", ctypes.c_double),
        ("sigma_ref", ctypes.c_double),
        ("edges", ctypes.c_double * (MAXL + 1)), ("sigma_remap", ctypes.c_double * MAXL),
        ("sigma_test", ctypes.c_double * MAXL),
        ("rel_cur", ctypes.c_double), ("add_cur", ctypes.c_double),
        ("rel_test", ctypes.c_double), ("add_test", ctypes.c_double),
        ("data", ctypes.c_double * MAXC), ("J_in", ctypes.c_double * (MAXC * MAXL)),
        ("pred_in", ctypes.c_double * MAXC),
        ("hessian", ctypes.c_double * (MAXL * MAXL)), ("gradient", ctypes.c_double * MAXL),
        ("newton_mean", ctypes.c_double * MAXL), ("pred_test", ctypes.c_double * MAXC),
        ("misfit_test", ctypes.c_double), ("prior_test", ctypes.c_double), ("likelihood_test", ctypes.c_double),
        ("proposal", ctypes.c_double), ("proposal1", ctypes.c_double),
    ]


def build(force=False):
    """Compile the oracle with gcc (a few seconds)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in ("fdem1d_oracle.c", "rjmcmc_oracle.c", "oracle.h")):
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
    """Run one chain; returns a dict of numpy arrays."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    nd = lib().gbo_n_depth(ctypes.byref(opt))
    N2 = 2 * opt.n_markov_chains
    r = dict(
        hitmap=np.zeros((opt.n_sigma_bins, nd), np.int32), edges_hist=np.zeros(nd, np.int32),
        ncells_hist=np.zeros(opt.max_layers + 1, np.int32), rel_hist=np.zeros(opt.n_err_bins, np.int32),
        add_hist=np.zeros(opt.n_err_bins, np.int32), misfit_trace=np.zeros(N2), accept_trace=np.zeros(N2, np.uint8),
        best_sigma=np.zeros(opt.max_layers), best_edges=np.zeros(opt.max_layers + 1),
        cur_sigma=np.zeros(opt.max_layers), cur_edges=np.zeros(opt.max_layers + 1), scalars=np.zeros(NSCALARS))
    co = ChainOutC()
    for k, v in r.items():
        setattr(co, k, v.ctypes.data)
    rc = lib().gbo_run_chain(ctypes.addressof(sys), ctypes.addressof(opt), data.ctypes.data, float(altitude),
                             int(seed), int(sounding_index), int(max_iterations), ctypes.addressof(co))
    assert rc == 0
    return r


def eval_transition(sys, opt, **kw):
    t = TransitionC()
    k = int(kw["k"])
    t.k, t.action, t.altitude, t.sigma_ref = k, int(kw["action"]), float(kw["altitude"]), float(kw["sigma_ref"])
    C = 2 * sys.n_freq
    for i in range(k + 1):
        t.edges[i] = float(kw["edges"][i])
    for i in range(k):
        t.sigma_remap[i] = float(kw["sigma_remap"][i])
        t.sigma_test[i] = float(kw["sigma_test"][i])
    t.rel_cur, t.add_cur, t.rel_test, t.add_test = (float(kw[n]) for n in ("rel_cur", "add_cur", "rel_test", "add_test"))
    for i in range(C):
        t.data[i] = float(kw["data"][i])
        t.pred_in[i] = float(kw["pred_in"][i])
    Jin = np.asarray(kw["J_in"], dtype=np.float64).reshape(C, -1)
    if Jin.shape[1] == k:
        for c in range(C):
            for i in range(k):
                t.J_in[c * k + i] = Jin[c, i]
    rc = lib().gbo_eval_transition(ctypes.byref(sys), ctypes.byref(opt), ctypes.byref(t))
    return rc, dict(
        hessian=np.array(t.hessian[:k * k]).reshape(k, k), gradient=np.array(t.gradient[:k]),
        newton_mean=np.array(t.newton_mean[:k]), pred_test=np.array(t.pred_test[:C]),
        misfit_test=t.misfit_test, prior_test=t.prior_test, likelihood_test=t.likelihood_test,
        proposal=t.proposal, proposal1=t.proposal1)
