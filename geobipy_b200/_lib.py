"""ctypes binding of the C-ABI library (include/geobipy_b200.h).

The shared library is built in-tree by ``geobipy_b200.build.build()`` (nvcc, sm_100a).  There is
no CPU fallback: if the library is missing, or no CUDA device is present, the compute entry points
raise.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GBP_LIB_PATH") or os.path.join(HERE, "libgeobipy_b200.so")   # GBP_LIB_PATH: A/B builds (scripts/)

MAXF, MAXL = 16, 30
MAXC = 2 * MAXF
TD_MAXSYS, TD_NFREQ, TD_MAXLAM, TD_MAXWIN, TD_MAXC, TD_MAXWAVE, TD_MAXFILT = 2, 32, 32, 32, 64, 64, 4
TD_SAMPLER_MAXC = 48   # channels the time-domain SAMPLER holds per chain (GBP_TD_SAMPLER_MAXC)
NSCALARS = 32
PRECISION_F32, PRECISION_F64 = 32, 64

(S_ITER, S_BURNED_IN, S_BURNED_IN_ITER, S_BEST_ITER, S_BEST_K, S_CUR_K, S_HALFSPACE, S_FAILED, S_N_ACCEPT,
 S_N_FORWARD, S_N_SENS, S_BEST_POSTERIOR, S_CUR_REL, S_CUR_ADD, S_CUR_MISFIT, S_CUR_PRIOR, S_CUR_LIKELIHOOD,
 S_BEST_REL, S_BEST_ADD, S_N_RESETS, S_N_BIRTH, S_N_DEATH, S_N_MOVE, S_N_NONE, S_TOTAL_ITER,
 S_CUR_REL2, S_CUR_ADD2, S_BEST_REL2, S_BEST_ADD2, S_CUR_HEIGHT, S_BEST_HEIGHT, S_HEIGHT_REF) = range(32)


class FdemSystemC(ctypes.Structure):
    """gbp_fdem_system"""
    _fields_ = [("n_freq", ctypes.c_int32), ("tid", ctypes.c_int32 * MAXF)] + [
        (n, ctypes.c_double * MAXF) for n in ("freq", "tmom", "tx", "ty", "tz", "rmom", "rx", "ry", "rz")]


class OptionsC(ctypes.Structure):
    """gbp_options"""
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
    """gbp_tdem_system"""
    _fields_ = [
        ("n_wave", ctypes.c_int32), ("n_windows", ctypes.c_int32), ("n_filters", ctypes.c_int32),
        ("n_abscissae", ctypes.c_int32),
        ("base_frequency", ctypes.c_double), ("digitising_frequency", ctypes.c_double),
        ("loop_radius", ctypes.c_double),
        ("wave_time", ctypes.c_double * TD_MAXWAVE), ("wave_current", ctypes.c_double * TD_MAXWAVE),
        ("window_start", ctypes.c_double * TD_MAXWIN), ("window_end", ctypes.c_double * TD_MAXWIN),
        ("filter_cutoff", ctypes.c_double * TD_MAXFILT), ("filter_order", ctypes.c_int32 * TD_MAXFILT),
        ("output_type", ctypes.c_int32), ("pad2_", ctypes.c_int32), ("peak_current", ctypes.c_double),
        ("x_scaling", ctypes.c_double), ("z_scaling", ctypes.c_double),
    ]


class TdemSurveyC(ctypes.Structure):
    """gbp_tdem_survey"""
    _fields_ = [("n_systems", ctypes.c_int32), ("error_model", ctypes.c_int32), ("sys", TdemSystemC * TD_MAXSYS),
                ("rx_dx", ctypes.c_double), ("rx_dy", ctypes.c_double), ("rx_dz", ctypes.c_double),
                ("additive_level", ctypes.c_double * TD_MAXC)]


BUFFER_FIELDS = ("hitmap", "edges_hist", "ncells_hist", "rel_hist", "add_hist", "misfit_trace", "accept_trace",
                 "best_sigma", "best_edges", "cur_sigma", "cur_edges", "scalars", "height_hist")


class ChainBuffersC(ctypes.Structure):
    """gbp_chain_buffers"""
    _fields_ = [(n, ctypes.c_void_p) for n in BUFFER_FIELDS]


# every symbol include/geobipy_b200.h declares
EXPORTS = (
    "gbp_version", "gbp_last_error", "gbp_device_count", "gbp_n_depth", "gbp_flops_per_forward",
    "gbp_filter_points", "gbp_launch_count", "gbp_last_kernel_ms", "gbp_kernel_ms_stats", "gbp_mufu_per_forward", "gbp_measure_peaks", "gbp_debug_counters", "gbp_debug_finish_times", "gbp_debug_progress_times", "gbp_tdem_primary_field",
    "gbp_tdem_mufu_per_forward",
    "gbp_fdem_forward", "gbp_fdem_sensitivity", "gbp_fdem_forward_host", "gbp_fdem_sensitivity_host",
    "gbp_rjmcmc_run", "gbp_rjmcmc_run_host", "gbp_release_host_buffers", "gbp_summarise_hitmap",
    "gbp_summarise_posterior", "gbp_opacity_doi",
    "gbp_tdem_n_channels", "gbp_tdem_window_operator", "gbp_tdem_flops_per_forward",
    "gbp_tdem_forward", "gbp_tdem_sensitivity", "gbp_tdem_forward_host", "gbp_tdem_sensitivity_host",
    "gbp_tdem_rjmcmc_run", "gbp_tdem_rjmcmc_run_host",
)

_lib = None


class GeobipyB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA extension; raises (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GeobipyB200Error(
            "CUDA extension %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    lib.gbp_version.restype = ctypes.c_char_p
    lib.gbp_last_error.restype = ctypes.c_char_p
    lib.gbp_device_count.restype = i32
    lib.gbp_n_depth.restype = i32
    lib.gbp_n_depth.argtypes = [vp]
    lib.gbp_flops_per_forward.restype = dbl
    lib.gbp_flops_per_forward.argtypes = [vp, i32]
    lib.gbp_filter_points.restype = i32
    lib.gbp_filter_points.argtypes = [vp]
    lib.gbp_launch_count.restype = i64
    lib.gbp_last_kernel_ms.restype = i32
    lib.gbp_last_kernel_ms.argtypes = [vp]
    lib.gbp_kernel_ms_stats.restype = i32
    lib.gbp_kernel_ms_stats.argtypes = [i32, vp, vp]
    lib.gbp_fdem_forward.restype = i32
    lib.gbp_fdem_forward.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    lib.gbp_fdem_sensitivity.restype = i32
    lib.gbp_fdem_sensitivity.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.gbp_fdem_forward_host.restype = i32
    lib.gbp_fdem_forward_host.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, i32]
    lib.gbp_fdem_sensitivity_host.restype = i32
    lib.gbp_fdem_sensitivity_host.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.gbp_rjmcmc_run.restype = i32
    lib.gbp_rjmcmc_run.argtypes = [vp, vp, i32, vp, vp, u64, u64, i64, vp, i32, vp]
    lib.gbp_rjmcmc_run_host.restype = i32
    lib.gbp_rjmcmc_run_host.argtypes = [vp, vp, i32, vp, vp, u64, u64, i64, vp, i32, i32]
    lib.gbp_mufu_per_forward.restype = dbl
    lib.gbp_mufu_per_forward.argtypes = [vp, i32]
    lib.gbp_tdem_mufu_per_forward.restype = dbl
    lib.gbp_tdem_mufu_per_forward.argtypes = [vp, i32]
    lib.gbp_summarise_hitmap.restype = i32
    lib.gbp_summarise_hitmap.argtypes = [vp, i32, i32, i32, vp, dbl, vp, i32, vp, vp, vp]
    lib.gbp_summarise_posterior.restype = i32
    lib.gbp_summarise_posterior.argtypes = [vp, i32, i32, i32, vp, dbl, vp, i32, dbl, vp, vp, vp, vp, vp]
    lib.gbp_opacity_doi.restype = i32
    lib.gbp_opacity_doi.argtypes = [vp, i32, i32, vp, i32, dbl, dbl, vp, vp, vp, vp, vp]
    lib.gbp_release_host_buffers.restype = i32
    lib.gbp_debug_finish_times.restype = i32
    lib.gbp_debug_finish_times.argtypes = [vp, i32]
    lib.gbp_debug_progress_times.restype = i32
    lib.gbp_debug_progress_times.argtypes = [vp, i32]
    lib.gbp_tdem_primary_field.restype = i32
    lib.gbp_tdem_primary_field.argtypes = [vp, vp]
    lib.gbp_debug_counters.restype = i32
    lib.gbp_debug_counters.argtypes = [vp, i32]
    lib.gbp_measure_peaks.restype = i32
    lib.gbp_measure_peaks.argtypes = [vp, vp]
    lib.gbp_tdem_n_channels.restype = i32
    lib.gbp_tdem_n_channels.argtypes = [vp]
    lib.gbp_tdem_window_operator.restype = i32
    lib.gbp_tdem_window_operator.argtypes = [vp, vp, vp, vp, vp]
    lib.gbp_tdem_flops_per_forward.restype = dbl
    lib.gbp_tdem_flops_per_forward.argtypes = [vp, i32]
    for name, ref in (("gbp_tdem_forward", lib.gbp_fdem_forward), ("gbp_tdem_sensitivity", lib.gbp_fdem_sensitivity),
                      ("gbp_tdem_forward_host", lib.gbp_fdem_forward_host),
                      ("gbp_tdem_sensitivity_host", lib.gbp_fdem_sensitivity_host),
                      ("gbp_tdem_rjmcmc_run", lib.gbp_rjmcmc_run), ("gbp_tdem_rjmcmc_run_host", lib.gbp_rjmcmc_run_host)):
        getattr(lib, name).restype = i32
        getattr(lib, name).argtypes = ref.argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GeobipyB200Error(load().gbp_last_error().decode() or "geobipy_b200 call failed (rc=%d)" % rc)


def require_cuda():
    lib = load()
    if lib.gbp_device_count() < 1:
        raise GeobipyB200Error("no CUDA device visible - geobipy_b200 has no CPU fallback")
    return lib


def ptr(a):
    """Pointer of a C-contiguous numpy array (host entry points)."""
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
