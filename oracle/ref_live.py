"""TEST / BENCH INFRASTRUCTURE ONLY - never imported by the product path.

Drives the UNMODIFIED reference package (DOI-USGS/geobipy 2.3.1) where it can be imported: from /root/reference in the
build container, or from `baseline/_ref` (the one offline `pip install --no-deps --target baseline/_ref` of the
reference; git-ignored, travels to the GPU box) - through the import stubs of `ref_shims.py` for the plotting / HDF5 /
MPI modules this image lacks.  numba IS in the image, so the reference's own forward kernels run as shipped.

Used by `bench.py` (`cpu_baseline_reference`: the reference's own `Inference1D.accept_reject() + update()` loop,
Inference1D.py:537, :705, timed on the host cores beside the C port) and by nothing else.
"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# documentation_source/source/supplementary/data/resolve.stm (the data files are not part of the installed package)
RESOLVE_STM = """freq, tor, tmom, tx, ty, tzoff, ror, rmom, rx, ry, rzoff
380, z, 1, 0, 0, 0, z, 1, 7.93, 0, 0
1776, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0
3345, x, -1, 0, 0, 0, x, 1, 9.03, 0, 0
8171, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0
41020, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0
129550, z, 1, 0, 0, 0, z, 1, 7.89, 0, 0
"""


def reference_root():
    """Where an importable copy of the reference lives on this machine, or None."""
    for p in (os.environ.get("GEOBIPY_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "geobipy", "src")):
            return p
    return None


def import_reference():
    root = reference_root()
    if root is None:
        raise RuntimeError("no importable copy of the reference (baseline/_ref or /root/reference)")
    sys.path.insert(0, HERE)
    import ref_shims
    ref_shims.REFERENCE_ROOT = root
    return ref_shims.import_reference()


def resolve_options(n_markov_chains):
    """The keys `user_parameters.read(resolve_options)` hands to Inference1D (documentation_source/.../resolve_options
    with the None entries replaced as user_parameters.py:40-44 does)."""
    return dict(
        n_markov_chains=n_markov_chains, interactive_plot=False, update_plot_every=5000, save_png=False, save_hdf5=True,
        solve_parameter=False, solve_gradient=True, solve_relative_error=True, solve_additive_error=True,
        solve_height=False, solve_calibration=False, maximum_number_of_layers=30, minimum_depth=0.1, maximum_depth=200.0,
        minimum_thickness=1.0, initial_relative_error=0.05, minimum_relative_error=0.001, maximum_relative_error=0.5,
        initial_additive_error=5.0, minimum_additive_error=3.0, maximum_additive_error=20.0, maximum_height_change=1.0,
        relative_error_proposal_variance=1e-6, additive_error_proposal_variance=1e-6, height_proposal_variance=0.01,
        probability_of_birth=1.0 / 6.0, probability_of_death=1.0 / 6.0, probability_of_perturb=1.0 / 6.0,
        probability_of_no_change=0.5, factor=np.float64(10.0), gradient_standard_deviation=1.5,
        covariance_scaling=np.float64(1.0), multiplier=np.float64(1.0), clip_ratio=None, ignore_likelihood=False,
        parameter_limits=None, reciprocate_parameters=True, verbose=False, stochastic_newton=True)


_SYSTEM = None


def _system():
    global _SYSTEM
    if _SYSTEM is None:
        from geobipy import FdemSystem
        with tempfile.NamedTemporaryFile("w", suffix=".stm", delete=False) as f:
            f.write(RESOLVE_STM)
        _SYSTEM = FdemSystem.read(f.name)
        os.unlink(f.name)
    return _SYSTEM


def time_chain(job):
    """Run the reference's own sampler loop (Inference1D.infer :650-677 without the HDF5 write) on one sounding of
    the bench workload for at most `max_iterations` iterations.  job = (sounding index, data [12], height, n_markov_chains,
    max_iterations, seed).  Returns (iterations done, seconds) - numba compilation happens in `initialize`, outside."""
    idx, data, z, n_markov_chains, max_iterations, seed = job
    import warnings
    warnings.filterwarnings("ignore")
    import_reference()
    from geobipy import FdemDataPoint, Inference1D, get_prng
    kw = resolve_options(n_markov_chains)
    kw["prng"] = get_prng(seed=seed + idx)
    inf = Inference1D(**kw)
    dp = FdemDataPoint(x=0.0, y=0.0, z=float(z), elevation=0.0, data=np.asarray(data, dtype=np.float64), std=None,
                       predictedData=None, system=_system(), lineNumber=0.0, fiducial=0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        inf.initialize(dp)     # the half-space search: compiles the numba kernels
        t0 = time.perf_counter()
        go, n = True, 0
        while go and n < max_iterations:
            failed = inf.accept_reject()
            inf.update()
            n += 1
            go = (not failed) and (inf.iteration <= inf.n_markov_chains + inf.burned_in_iteration)
            if (not failed) and (not inf.burned_in):
                go = inf.iteration < inf.n_markov_chains
        dt = time.perf_counter() - t0
    return n, dt
