"""CPU tests: the C-ABI library loads, exports every symbol include/geobipy_b200.h declares, and the
host-side logic (options mapping, grids, synthetic inputs) behaves.  No compute call is made."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "geobipy_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gbp_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built_lib):
    from geobipy_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(built_lib, name), name
    assert sorted(_lib.EXPORTS) == declared
    assert built_lib.gbp_version().startswith(b"geobipy_b200")


def test_struct_layouts_match_oracle(built_lib, oracle):
    from geobipy_b200 import _lib
    assert ctypes.sizeof(_lib.FdemSystemC) == ctypes.sizeof(oracle.FdemSystemC)
    assert ctypes.sizeof(_lib.OptionsC) == ctypes.sizeof(oracle.OptionsC)
    assert [f[0] for f in _lib.OptionsC._fields_] == [f[0] for f in oracle.OptionsC._fields_]


def test_host_helpers(built_lib):
    from geobipy_b200 import ops
    s = ops.resolve_system_struct()
    assert list(s.tid[:6]) == [9, 9, 1, 9, 9, 9]
    assert ops.filter_points(s) == 5 * 120 + 120 + 140 == 860
    assert ops.flops_per_forward(s, 10) == 860 * (75 * 10 + 39)
    o = ops.make_options(maximum_number_of_layers=12, minimum_depth=1.0, maximum_depth=550.0, minimum_thickness=2.0,
                         covariance_scaling=None, factor=None, n_markov_chains=1234)
    assert (o.max_layers, o.min_edge, o.max_edge, o.min_width, o.n_markov_chains) == (12, 1.0, 550.0, 2.0, 1234)
    assert o.covariance_scaling == 1.0 and o.factor == 10.0          # user_parameters.py:40-44 defaults
    assert ops.n_depth(ops.make_options()) == 440                      # resolve_options -> hitmap [250, 440]
    assert ops.n_depth(o) == len(np.arange(0.0, 1.1 * 550.0, 1.0)) - 1
    shapes = ops.chain_buffer_shapes(ops.make_options(n_markov_chains=10), 3)
    assert shapes["hitmap"][0] == (3, 250, 440) and shapes["misfit_trace"][0] == (3, 20)


def test_no_cpu_fallback_without_gpu(built_lib):
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    from geobipy_b200 import _lib, ops
    if built_lib.gbp_device_count() > 0:
        pytest.skip("a GPU is present")
    s = ops.resolve_system_struct()
    with pytest.raises(_lib.GeobipyB200Error):
        ops.fdem_forward(s, [1], [[0.01]], [[np.inf]], [30.0])
    with pytest.raises(_lib.GeobipyB200Error):
        ops.rjmcmc_run(s, ops.make_options(), np.ones((1, 12)), np.array([30.0]))


def test_time_domain_sampler_rejects_more_channels_than_it_holds(built_lib):
    """the sampler kernels hold GBP_TD_SAMPLER_MAXC = 48 channels per chain in shared memory while the
    forward operators take up to GBP_TD_MAXC = 64 (2 x 32 windows).  A 2 x 26-window datapoint type must be refused
    loudly - before any device work, so this runs without a GPU - instead of writing past the per-chain arrays."""
    from geobipy_b200 import _lib, ops
    defs = ops.skytem_definitions()
    hm = dict(defs[0])
    assert len(hm["window_start"]) == 26
    sv = ops.make_tdem_survey_struct([hm, dict(hm)])
    assert ops.n_channels(sv) == 52 and 48 == _lib.TD_SAMPLER_MAXC < ops.n_channels(sv) <= _lib.TD_MAXC
    opt = ops.make_options(n_markov_chains=100, **ops.SKYTEM_OPTIONS)
    cb = _lib.ChainBuffersC()
    cb.scalars = 1   # non-NULL; never dereferenced: the channel check comes first
    for fn in (built_lib.gbp_tdem_rjmcmc_run, ):
        rc = fn(ctypes.addressof(sv), ctypes.addressof(opt), 1, None, None, 0, 0, 0, ctypes.addressof(cb), 32, None)
        assert rc != 0 and b"48 data channels" in built_lib.gbp_last_error()
    data, alt = np.full((1, 52), 1e-12), np.array([30.0])
    h = _lib.ChainBuffersC()
    sc = np.zeros((1, _lib.NSCALARS))
    h.scalars = sc.ctypes.data
    if built_lib.gbp_device_count() > 0:
        rc = built_lib.gbp_tdem_rjmcmc_run_host(ctypes.addressof(sv), ctypes.addressof(opt), 1, data.ctypes.data, alt.ctypes.data,
                                                0, 0, 0, ctypes.addressof(h), 32, 0)
        assert rc != 0 and b"48 data channels" in built_lib.gbp_last_error()
    # the SkyTEM dual-moment type (26 + 19 = 45 channels) is inside the limit
    assert ops.n_channels(ops.skytem_survey_struct()) == 45


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "geobipy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in src and "ref_shims" not in src and "oracle/" not in src, f


def test_synthetic_inputs():
    from geobipy_b200.synthetic import synthetic_batch, synthetic_sounding
    e, s, h, n = synthetic_sounding(7)
    e2, s2, h2, n2 = synthetic_sounding(7)
    assert np.array_equal(e, e2) and np.array_equal(s, s2) and h == h2 and np.array_equal(n, n2)
    assert e[0] == 0.0 and np.isinf(e[-1]) and (np.diff(e[:-1]) >= 1.0).all() and 25 <= h <= 45
    b = synthetic_batch(5, 4)
    assert b["sigma"].shape == (4, 30) and b["nlayers"][2] == s.size
    assert np.array_equal(b["sigma"][2, : s.size], s)


def test_tdem_structs_and_tables(built_lib, oracle):
    """The time-domain window operator the product builds in C++ (gbp_tdem_tables.h) against the oracle's
    independent numpy/scipy construction (oracle_py.tdem_window_operator): host only, no GPU."""
    from geobipy_b200 import _lib, ops
    sv = ops.skytem_survey_struct()
    assert ops.n_channels(sv) == 45 and not ops.is_tdem(ops.resolve_system_struct())
    f, MR, MI, t = ops.tdem_window_operator(sv)
    S = oracle.make_tdem_system()
    assert np.allclose(f, np.array(S.freq[:32]), rtol=1e-14)
    MRo, MIo = np.array(S.MR[:45 * 32]).reshape(45, 32), np.array(S.MI[:45 * 32]).reshape(45, 32)
    assert np.all(np.abs(MR - MRo).max(axis=1) <= 1e-12 * np.abs(MRo).max(axis=1))
    assert np.all(np.abs(MI - MIo).max(axis=1) <= 1e-12 * np.abs(MIo).max(axis=1))
    assert np.allclose(t, np.array(S.t_centre[:45]))
    assert ops.flops_per_forward(sv, 10) == 32 * 22 * (75 * 10 + 39) + 2 * 45 * 64
    o = ops.make_options(initial_additive_error=[2e-14, 2e-13], minimum_depth=1.0, maximum_depth=550.0)
    assert o.n_systems == 2 and o.add_init == 2e-14 and o.add_init2 == 2e-13 and ops.n_depth(o) == 1209
    assert ops.chain_buffer_shapes(o, 3)["rel_hist"][0] == (3, 2, 99)
    oo = oracle.skytem_options()
    so = ops.make_options(**ops.SKYTEM_OPTIONS)
    for name, _ in _lib.OptionsC._fields_:
        assert getattr(oo, name) == getattr(so, name), name
    # invalid descriptions are rejected with a message, not a crash
    bad = ops.skytem_survey_struct()
    bad.sys[0].base_frequency = 25.0   # waveform no longer spans half a period
    with pytest.raises(_lib.GeobipyB200Error):
        ops.tdem_window_operator(bad)


def test_solve_z_option_keys_and_buffers(built_lib):
    """The options file's solve_z / maximum_z_change / z_proposal_variance (Point.set_priors pointcloud/Point.py:959-961,
    set_proposals :977-979) reach gbp_options; the height histogram is an output only when the height is sampled."""
    from geobipy_b200 import _lib, ops
    o = ops.make_options()
    assert o.solve_height == 0 and "height_hist" not in ops.DEFAULT_OUTPUTS and "height_hist" in _lib.BUFFER_FIELDS
    o = ops.make_options(solve_z=True, maximum_z_change=2.5, z_proposal_variance=0.04)
    assert o.solve_height == 1 and o.max_height_change == 2.5 and o.height_prop_var == 0.04
    shp = ops.chain_buffer_shapes(o, 5)
    assert shp["height_hist"] == ((5, 99), np.int32)
    assert _lib.S_CUR_HEIGHT == 29 and _lib.S_BEST_HEIGHT == 30 and _lib.S_HEIGHT_REF == 31 and _lib.NSCALARS == 32
    # the struct mirrors stay in step with the header (field order and size)
    import ctypes
    assert [n for n, _ in _lib.OptionsC._fields_][-4:] == ["solve_height", "pad_h_", "max_height_change", "height_prop_var"]
    assert ctypes.sizeof(_lib.OptionsC) == 288 and ctypes.sizeof(_lib.ChainBuffersC) == 13 * ctypes.sizeof(ctypes.c_void_p)


def test_options_from_a_reference_options_file(built_lib):
    """Keyword arguments that come from a reference options file: the height keys the shipped files carry are dead in
    the reference (Point.set_priors reads solve_z; checked on the live reference: solve_height=True leaves
    datapoint.z without a prior) and are ignored here too; unknowns that are not built raise."""
    import warnings
    from geobipy_b200 import api, ops
    shipped = dict(solve_height=False, maximum_height_change=1.0, height_proposal_variance=0.01, solve_calibration=False,
                   solve_transmitter_z=False, solve_receiver_pitch=False, n_markov_chains=1000)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        o = ops.options_from_reference(**shipped)
    assert o.solve_height == 0 and o.n_markov_chains == 1000
    with pytest.warns(UserWarning, match="solve_height"):
        o = ops.options_from_reference(**dict(shipped, solve_height=True))
    assert o.solve_height == 0
    o = ops.options_from_reference(**dict(shipped, solve_z=True, maximum_z_change=0.5, z_proposal_variance=0.02))
    assert o.solve_height == 1 and o.max_height_change == 0.5 and o.height_prop_var == 0.02
    for key in ("solve_transmitter_pitch", "solve_receiver_pitch", "solve_receiver_z", "solve_calibration", "solve_x"):
        with pytest.raises(NotImplementedError, match=key):
            ops.options_from_reference(**dict(shipped, **{key: True}))
    # the transmitter height of a time-domain datapoint IS built (KIND_TDEM_Z): same gbp_options fields as solve_z
    o = ops.options_from_reference(**dict(shipped, solve_transmitter_z=True, maximum_transmitter_z_change=2.0,
                                          transmitter_z_proposal_variance=0.04))
    assert o.solve_height == 1 and o.max_height_change == 2.0 and o.height_prop_var == 0.04 and o.height_key == "solve_transmitter_z"
    # the object-level mirror goes through the same filter
    with pytest.warns(UserWarning, match="solve_height"):
        inf = api.Inference1D(prng=np.random.default_rng(0), **dict(shipped, solve_height=True))
    assert inf.options.solve_height == 0
    with pytest.raises(NotImplementedError):
        api.Inference1D(prng=np.random.default_rng(0), **dict(shipped, solve_receiver_z=True))
    # ... and the transmitter height is only accepted for a time-domain datapoint
    inf = api.Inference1D(prng=np.random.default_rng(0), **dict(shipped, solve_transmitter_z=True))
    with pytest.raises(AssertionError):
        inf.initialize(api.FdemDataPoint(z=30.0, system=api.FdemSystem([380.0], api.CircularLoop(["z"], [1.0], [0.0], [0.0], [0.0]),
                                                                       api.CircularLoop(["z"], [1.0], [7.93], [0.0], [0.0]))))


def test_torch_ops_are_registered():
    """the three operators exist under torch.ops.geobipy_b200 with tensor-in / tensor-out schemas (the CUDA
    implementations call the C-ABI; without a device only the registration can be checked)."""
    import torch
    from geobipy_b200 import torch_ops  # noqa: F401
    for name in ("fdem_forward", "fdem_sensitivity", "rjmcmc_run"):
        op = getattr(torch.ops.geobipy_b200, name)
        assert "Tensor system" in str(op.default._schema)
    assert str(torch.ops.geobipy_b200.rjmcmc_run.default._schema).endswith("-> Tensor[]")
    with pytest.raises((NotImplementedError, RuntimeError)):   # no CPU implementation: the dispatcher refuses CPU tensors
        z = torch.zeros(1)
        torch.ops.geobipy_b200.fdem_forward(z.to(torch.uint8), z.to(torch.int32), z.double().reshape(1, 1), z.double().reshape(1, 1), z.double(), 32)
