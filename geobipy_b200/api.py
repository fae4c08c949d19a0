"""Host-side mirror of the reference's object interface for the per-sounding inference path.

Same names, argument meaning and error behaviour as the reference classes on this path, so that the
calls a GeoBIPy driver makes (`Inference3D.infer_serial`, geobipy/src/inversion/Inference3D.py:483-492)

    dp  = FdemDataPoint(x, y, z, elevation, data=..., system=FdemSystem.read("resolve.stm"))
    inf = Inference1D(**options, prng=...)        # options = keys of a reference options file
    inf.initialize(dp)
    failed = inf.infer(hdf_file_handle)

run on the GPU.  Everything numerical happens behind the C-ABI (ops.py); these classes only carry data.
Mirrored: FdemSystem (classes/system/FdemSystem.py), CircularLoop (classes/system/CircularLoop.py),
RectilinearMesh1D / Model (containers only), FdemDataPoint (classes/data/datapoint/FdemDataPoint.py,
DataPoint.py, EmDataPoint.py), Histogram (classes/statistics/Histogram.py: counts + mean/percentile),
Inference1D (inversion/Inference1D.py) and a batched Inference3D.infer replacement (`infer_batch`).
"""
import numpy as np

from . import _lib, ops

__all__ = ["CircularLoop", "FdemSystem", "RectilinearMesh1D", "Model", "FdemDataPoint", "Histogram", "StatArray",
           "Inference1D", "infer_batch"]

_ORI = {"x": 0, "y": 1, "z": 2}


class StatArray(np.ndarray):
    """An array (or 0-d scalar) that carries its name, units and - after an inference - its `posterior` Histogram, like
    the reference's StatArray (classes/core/StatArray.py): `model.values.posterior`, `model.mesh.nCells.posterior`,
    `datapoint.relative_error.posterior` ... are the attribute paths Inference1D.writeHdf serialises (:1050-1090)."""

    def __new__(cls, values, name=None, units=None, posterior=None):
        obj = np.asarray(values).view(cls)
        obj.name, obj.units, obj.posterior = name, units, posterior
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self.units = getattr(obj, "units", None)
        self.posterior = getattr(obj, "posterior", None)

    @property
    def hasPosterior(self):
        return self.posterior is not None


class CircularLoop:
    """Loop description: orientation ('x'|'y'|'z'), moment and offsets (classes/system/CircularLoop.py)."""

    def __init__(self, orientation=None, moment=None, x=None, y=None, z=None, **kwargs):
        n = 0 if orientation is None else len(orientation)
        self.orientation = [str(o).strip() for o in (orientation if orientation is not None else [])]
        self._orientation = np.asarray([_ORI[o] for o in self.orientation], dtype=np.int32)

        def arr(v):
            return np.zeros(n) if v is None else np.asarray(v, dtype=np.float64).reshape(-1)
        self.moment, self.x, self.y, self.z = arr(moment), arr(x), arr(y), arr(z)


class FdemSystem:
    """Frequency-domain acquisition system (classes/system/FdemSystem.py)."""

    def __init__(self, frequencies=None, transmitter=None, receiver=None):
        self.frequencies = np.asarray(frequencies if frequencies is not None else [], dtype=np.float64)
        self.transmitter = transmitter if transmitter is not None else CircularLoop()
        self.receiver = receiver if receiver is not None else CircularLoop()
        self._filename = None
        self._struct = None

    @classmethod
    def read(cls, filename):
        """Read an .stm file: header line, then `freq, tor, tmom, tx, ty, tz, ror, rmom, rx, ry, rz`
        per frequency (FdemSystem.read, FdemSystem.py:146-183)."""
        rows = []
        with open(filename) as f:
            next(f)
            for line in f:
                p = [t.strip() for t in line.strip().split(",")]
                if len(p) >= 11:
                    rows.append(p)
        col = lambda i: [r[i] for r in rows]  # noqa: E731
        num = lambda i: [float(r[i]) for r in rows]  # noqa: E731
        self = cls(num(0), CircularLoop(col(1), num(2), num(3), num(4), num(5)),
                   CircularLoop(col(6), num(7), num(8), num(9), num(10)))
        self._filename = filename
        return self

    @property
    def nFrequencies(self):
        return self.frequencies.size

    @property
    def tensor_id(self):
        return 1 + 3 * self.receiver._orientation + self.transmitter._orientation

    @property
    def loop_offsets(self):
        return np.vstack([self.receiver.x - self.transmitter.x, self.receiver.y - self.transmitter.y,
                          self.receiver.z - self.transmitter.z])

    @property
    def loop_separation(self):
        return np.linalg.norm(self.loop_offsets, axis=0)

    @property
    def lamda0(self):
        l0 = 10.0 ** (np.arange(120, dtype=np.float64) * 9.04226468670e-2 - 8.3885)
        return np.outer(1.0 / self.loop_separation, l0)

    @property
    def lamda1(self):
        l1 = 10.0 ** (np.arange(140, dtype=np.float64) * 8.7967143957e-2 - 7.91001919)
        return np.outer(1.0 / self.loop_separation, l1)

    @property
    def c_struct(self):
        if self._struct is None:
            t, r = self.transmitter, self.receiver
            self._struct = ops.make_system_struct(self.frequencies, t.orientation, t.moment, t.x, t.y, t.z,
                                                  r.orientation, r.moment, r.x, r.y, r.z)
        return self._struct


class RectilinearMesh1D:
    """Layer edges container (classes/mesh/RectilinearMesh1D.py): edges[0] = top, edges[-1] = inf."""

    def __init__(self, edges=None, centres=None, widths=None, **kwargs):
        if edges is None and widths is not None:
            edges = np.r_[0.0, np.cumsum(widths)]
        self.edges = StatArray(np.asarray(edges, dtype=np.float64), "Depth", "m")
        self._nCells = StatArray(np.int32(self.edges.size - 1), "Number of layers")

    @property
    def nCells(self):
        """Number of layers: an integer-like 0-d StatArray (its `posterior` is the layer-count histogram)."""
        return self._nCells

    @property
    def nEdges(self):
        return self.edges.size

    @property
    def widths(self):
        return np.diff(self.edges)

    @property
    def centres(self):
        return 0.5 * (self.edges[1:] + self.edges[:-1])


class Model:
    """Values on a mesh (classes/model/Model.py, container part)."""

    def __init__(self, mesh=None, values=None):
        self.mesh = mesh
        self.values = StatArray(np.asarray(values, dtype=np.float64) if values is not None else np.zeros(int(mesh.nCells)),
                                "Conductivity", "S/m")

    @property
    def posterior(self):
        """The hitmap (alias of `values.posterior`, the reference's path)."""
        return self.values.posterior

    @posterior.setter
    def posterior(self, h):
        self.values.posterior = h

    @property
    def nCells(self):
        return int(self.mesh.nCells)


class Histogram:
    """Posterior accumulator (classes/statistics/Histogram.py): `counts` on a 1-D or 2-D mesh.

    2-D hitmaps are [n_x (conductivity bins), n_y (depth cells)] like the reference's
    `model.values.posterior.counts`; x edges are in linear units, binned uniformly in ln."""

    def __init__(self, counts, x_edges=None, y_edges=None, log_x=False):
        self.counts = np.asarray(counts)
        self.x_edges, self.y_edges, self.log_x = x_edges, y_edges, log_x

    @property
    def values(self):
        return self.counts

    def _xc(self):
        e = np.log(self.x_edges) if self.log_x else np.asarray(self.x_edges)
        return 0.5 * (e[1:] + e[:-1])

    def mean(self):
        """Mean along the value axis for every depth cell (Mesh._mean, classes/mesh/Mesh.py:80)."""
        c = self._xc()
        h = self.counts.astype(np.float64)
        if h.ndim == 1:
            m = (h * c).sum() / max(h.sum(), 1.0)
        else:
            m = (h * c[:, None]).sum(axis=0) / np.maximum(h.sum(axis=0), 1.0)
        return np.exp(m) if self.log_x else m

    def percentile(self, percent):
        """Value-axis percentile per depth cell (Mesh._percentile, classes/mesh/Mesh.py:173-217)."""
        c = self._xc()
        h = np.atleast_2d(self.counts.T).T if self.counts.ndim == 1 else self.counts
        cs = np.cumsum(h, axis=0).astype(np.float64)
        tot = cs[-1]
        # first bin whose cumulative fraction reaches percent * 0.01 - both in fp64, as the reference computes them
        # (95.0 * 0.01 > 0.95: a fraction of exactly 0.95 goes to the next bin); an empty column -> the last bin
        frac = np.divide(cs, tot, out=np.zeros_like(cs), where=tot > 0.0)
        idx = np.minimum((frac < percent * 0.01).sum(axis=0), c.size - 1)
        out = c[idx]
        out = np.exp(out) if self.log_x else out
        return out if self.counts.ndim > 1 else out.item()

    def median(self):
        return self.percentile(50.0)

    def mode(self):
        """Centre of the fullest value bin per depth cell (Mesh._mode, classes/mesh/Mesh.py:138-165)."""
        c = self._xc()
        out = c[np.argmax(self.counts, axis=0)]
        out = np.exp(out) if self.log_x else out
        return out if self.counts.ndim > 1 else float(out)

    def credible_intervals(self, percent=90.0):
        """(median, low, high) per depth cell (Mesh._credible_intervals, Mesh.py:30-55): the 50 %, (100 - percent) / 2
        and 100 - (100 - percent) / 2 percentiles."""
        p = 0.5 * min(percent, 100.0 - percent)
        return self.percentile(50.0), self.percentile(p), self.percentile(100.0 - p)

    def credible_range(self, percent=90.0):
        """Width of the credible interval per depth cell (Mesh._credible_range, Mesh.py:58-78): in decades for the
        conductivity axis of a hitmap (its mesh has log = 10, Model.py:677-679), linear otherwise."""
        _, lo, hi = self.credible_intervals(percent)
        if self.log_x:
            return np.abs(np.log10(hi) - np.log10(lo))
        return np.abs(np.asarray(hi) - np.asarray(lo))

    def transparency(self, percent=95.0):
        """Credible range normalised to [0, 1] over this histogram, NaN -> 1 (Histogram.transparency,
        classes/statistics/Histogram.py:509-541)."""
        out = np.asarray(self.credible_range(percent), dtype=np.float64)
        mn, mx = np.nanmin(out), np.nanmax(out)
        out = (out - mn) / (mx - mn) if (mx - mn) > 0.0 else out - mn
        return np.where(np.isnan(out), 1.0, out)

    def opacity(self, percent=95.0):
        """1 - transparency (Histogram.opacity, Histogram.py:330-354)."""
        return 1.0 - self.transparency(percent)

    def opacity_level(self, percent=95.0):
        """Depth-cell centre found from the bottom up where the transparency drops to percent / 100
        (Histogram.opacity_level, Histogram.py:356-367; the reference's loop index may run to -1 = the last cell)."""
        t = self.transparency(percent)
        yc = 0.5 * (np.asarray(self.y_edges)[1:] + np.asarray(self.y_edges)[:-1])
        i = t.size - 1
        while t[i] > 0.01 * percent and i >= 0:
            i -= 1
        return yc[i]


class FdemDataPoint:
    """One frequency-domain sounding (FdemDataPoint.py / EmDataPoint.py / DataPoint.py).

    data / predictedData are [in-phase(F), quadrature(F)] in ppm."""

    def __init__(self, x=0.0, y=0.0, z=0.0, elevation=0.0, data=None, std=None, predictedData=None, system=None,
                 lineNumber=0.0, fiducial=0.0, precision=_lib.PRECISION_F64, **kwargs):
        if isinstance(system, str):
            system = FdemSystem.read(system)
        if isinstance(system, (list, tuple)):
            assert len(system) == 1, NotImplementedError("one FDEM system per datapoint")
            system = system[0]
        assert isinstance(system, FdemSystem), TypeError("system must have type FdemSystem")
        self.system = system
        self.x, self.y, self.z, self.elevation = float(x), float(y), float(z), float(elevation)
        self.lineNumber, self.fiducial = lineNumber, fiducial
        n = 2 * system.nFrequencies
        self.data = np.zeros(n) if data is None else np.asarray(data, dtype=np.float64).copy()
        self._std = np.full(n, 0.01) if std is None else np.asarray(std, dtype=np.float64).copy()
        self.predictedData = np.zeros(n) if predictedData is None else np.asarray(predictedData, dtype=np.float64).copy()
        assert self.data.size == n and self.predictedData.size == n, ValueError("data must have 2 x nFrequencies entries")
        self.relative_error = np.asarray([0.01])
        self.additive_error = np.asarray([0.0])
        self._use_errors = False
        self.sensitivity_matrix = None
        self.precision = precision
        self.posteriors = {}

    # -- EmDataPoint.active :44-56
    @property
    def active(self):
        d = self.data.copy()
        d[~(d > 0.0)] = np.nan
        return ~np.isnan(d)

    @property
    def n_active_channels(self):
        return int(self.active.sum())

    @property
    def nChannels(self):
        return self.data.size

    # -- DataPoint.std :268-282
    @property
    def std(self):
        if self._use_errors:
            assert np.all(self.relative_error > 0.0), ValueError("relative_error must be > 0.0")
            self._std = np.sqrt((self.relative_error[0] * self.data) ** 2 + self.additive_error[0] ** 2)
        return self._std

    def initialize(self, **kwargs):
        self.relative_error = np.atleast_1d(np.asarray(kwargs["initial_relative_error"], dtype=np.float64))
        self.additive_error = np.atleast_1d(np.asarray(kwargs["initial_additive_error"], dtype=np.float64))
        self._use_errors = True

    @property
    def deltaD(self):
        return self.predictedData - self.data

    def _model_arrays(self, mod):
        assert isinstance(mod, Model), TypeError("Invalid model class for forward modeling [1D]")
        assert np.isinf(mod.mesh.edges[-1]), ValueError("mod.edges must have last entry be infinity for forward modelling.")
        assert self.z >= mod.mesh.edges[0], "Sensor altitude must be above the top of the model"  # fdem1d.py:29
        L = mod.nCells
        assert 1 <= L <= _lib.MAXL, ValueError("1..%d layers supported" % _lib.MAXL)
        return (np.asarray([L], np.int32), mod.values.reshape(1, L).astype(np.float64),
                mod.mesh.widths.reshape(1, L).astype(np.float64), np.asarray([self.z - mod.mesh.edges[0]]))

    def forward(self, mod):
        """Fill predictedData from a 1-D layered model (FdemDataPoint.forward :524 -> fdem1dfwd)."""
        nl, s, t, a = self._model_arrays(mod)
        self.predictedData[:] = ops.fdem_forward(self.system.c_struct, nl, s, t, a, precision=self.precision)[0]

    def sensitivity(self, mod, **kwargs):
        """Jacobian d(predicted)/d ln(sigma), shape [nChannels, nCells] (FdemDataPoint.sensitivity :531)."""
        nl, s, t, a = self._model_arrays(mod)
        _, J = ops.fdem_forward(self.system.c_struct, nl, s, t, a, precision=self.precision, sensitivity=True)
        self.sensitivity_matrix = J[0]
        return self.sensitivity_matrix

    def fm_dlogc(self, mod):
        nl, s, t, a = self._model_arrays(mod)
        p, J = ops.fdem_forward(self.system.c_struct, nl, s, t, a, precision=self.precision, sensitivity=True)
        self.predictedData[:] = p[0]
        self.sensitivity_matrix = J[0]

    # -- DataPoint.data_misfit :502-525, likelihood :491-500 (MvNormal log-pdf :209-216)
    def data_misfit(self):
        a = self.active
        return float(np.sum((self.deltaD[a] / self.std[a]) ** 2))

    def likelihood(self, log=True):
        a = self.active
        var = self.std[a] ** 2
        ll = -0.5 * a.sum() * np.log(2.0 * np.pi) - 0.5 * np.sum(np.log(var)) - 0.5 * np.sum(self.deltaD[a] ** 2 / var)
        return float(ll) if log else float(np.exp(ll))


class Inference1D:
    """Per-sounding rjMCMC sampler with the reference's interface (inversion/Inference1D.py).

    `__init__` takes the keys of a reference options file (resolve_options) plus `prng`/`seed`;
    `initialize(datapoint)` then `infer(hdf_file_handle)` -> failed.  Afterwards the attributes the
    reference's `writeHdf` (:1050-1090) serialises are available: model (+ `.posterior` hitmap),
    n_cells_posterior, edges_posterior, relative_error_posterior, additive_error_posterior, best_model,
    data_misfit_v, acceptance_v, iteration, burned_in, burned_in_iteration, best_iteration, halfspace."""

    def __init__(self, covariance_scaling=0.75, ignore_likelihood=False, interactive_plot=False, multiplier=1.0,
                 n_markov_chains=100000, parameter_limits=None, prng=None, seed=None, save_hdf5=True, save_png=False,
                 solve_gradient=True, solve_parameter=False, update_plot_every=5000, precision=_lib.PRECISION_F32,
                 device=0, sounding_index=0, **kwargs):
        assert interactive_plot or save_hdf5, Exception("You have chosen to neither view or save the inversion results!")
        assert not ignore_likelihood, NotImplementedError("ignore_likelihood=True (prior sampling) is not on the GPU path")
        assert parameter_limits is None, NotImplementedError("parameter_limits")
        if seed is None:
            # any numpy Generator can seed the counter-based device stream
            seed = int(prng.integers(0, 2 ** 63 - 1)) if prng is not None else 0
        self.seed, self.sounding_index = int(seed) & (2 ** 64 - 1), int(sounding_index)
        self.n_markov_chains, self.update_plot_every = int(n_markov_chains), int(update_plot_every)
        self.multiplier = np.float64(multiplier)   # carried and serialised, as in the reference (Inference1D.py:88, :1067)
        # carried and serialised (Inference1D.py:88, :114, :1031): the options files of the reference set it
        self.interactive_plot, self.reciprocate_parameter, self.limits = bool(interactive_plot), bool(kwargs.get("reciprocate_parameters", False)), None
        self.precision, self.device = precision, device
        self._raw_options = dict(kwargs, covariance_scaling=covariance_scaling, n_markov_chains=n_markov_chains,
                                 update_plot_every=update_plot_every, solve_gradient=int(bool(solve_gradient)),
                                 solve_parameter=int(bool(solve_parameter)))
        self.options = ops.options_from_reference(**self._raw_options)
        self.user_options = kwargs
        self.datapoint = None

    def initialize(self, datapoint):
        from .tdem import TdemDataPoint, Tempest_datapoint
        assert isinstance(datapoint, (FdemDataPoint, TdemDataPoint)), TypeError("datapoint must be a FdemDataPoint or TdemDataPoint")
        self.datapoint = datapoint
        self._tdem = isinstance(datapoint, TdemDataPoint)
        self._tempest = isinstance(datapoint, Tempest_datapoint)
        if self._tempest:
            # tempest_options: initial_additive_error is the additive level of every CHANNEL; the sampled "additive error"
            # is one multiplier per component, starting at 1 (Tempest_datapoint.set_priors / set_proposals :478-510)
            kw = dict(self._raw_options)
            level = np.asarray(kw["initial_additive_error"], dtype=np.float64).reshape(-1)
            nc = datapoint.nSystems * datapoint.n_components
            kw["initial_additive_error"] = [1.0] * nc if nc > 1 else 1.0
            self.options = ops.options_from_reference(**kw)
            datapoint.initialize(initial_relative_error=self._raw_options["initial_relative_error"], initial_additive_error=level)
            assert (2 if self.options.n_systems > 1 else 1) == nc, ValueError("the error options must have one entry per component ({})".format(nc))
            self.iteration, self.burned_in, self.burned_in_iteration = 0, False, 0
            return
        o = self.options
        ns = datapoint.nSystems if self._tdem else 1
        assert (2 if o.n_systems > 1 else 1) == ns, ValueError(
            "the error options must be lists with one entry per system ({})".format(ns))
        if ns == 2:
            datapoint.initialize(initial_relative_error=[o.rel_init, o.rel_init2], initial_additive_error=[o.add_init, o.add_init2])
        else:
            datapoint.initialize(initial_relative_error=o.rel_init, initial_additive_error=o.add_init)
        hk = getattr(o, "height_key", None)
        if o.solve_height and hk is not None:
            # the reference reads solve_z for the datapoint's own height (Point.set_priors :959-961) and
            # solve_transmitter_z for the transmitter loop of a time-domain datapoint (Loop_pair / EmLoop)
            assert (hk == "solve_transmitter_z") == self._tdem, NotImplementedError(
                "%s is not built for a %s datapoint" % (hk, "time-domain" if self._tdem else "frequency-domain"))
        self.iteration, self.burned_in, self.burned_in_iteration = 0, False, 0

    def infer(self, hdf_file_handle=None, max_iterations=0, index=None):
        """Run the chain to the reference's termination rule.  With a file handle (any h5py-shaped group laid out by
        `createHdf` / `Inference2D.createHdf`) the result is written into it at the end, as the reference does
        (Inference1D.infer :679-688 -> writeHdf)."""
        dp = self.datapoint
        if dp.n_active_channels == 0:
            return True
        struct = dp.survey_struct() if self._tempest else (dp.c_struct if self._tdem else dp.system.c_struct)
        if not hasattr(dp, "z_input"):
            dp.z_input = float(dp.transmitter.z if self._tdem else dp.z)   # a second infer() starts from the input height again
        alt = dp.z_input
        r = ops.rjmcmc_run(struct, self.options, dp.data.reshape(1, -1), np.asarray([alt]), seed=self.seed,
                           first_index=self.sounding_index, max_iterations=max_iterations, precision=self.precision,
                           device=self.device)
        self._fill(r, 0)
        self._result = r
        if hdf_file_handle is not None:
            self.writeHdf(hdf_file_handle, index=index)
        return self.failed

    # -- HDF5 (Inference1D.createHdf :1002-1048, writeHdf :1050-1090) through geobipy_b200.hdf
    def _as_dataset(self, dp=None):
        """The datapoint as a one-sounding data set (what geobipy_b200.hdf lays files out from)."""
        dp = self.datapoint if dp is None else dp
        one = lambda v: np.asarray([float(v)])
        if self._tdem:
            from .tdem import TdemData, TempestData
            tx, rx = dp.transmitter, dp.receiver
            z_in = getattr(dp, "z_input", float(tx.z))
            geometry = np.asarray([[tx.pitch, tx.roll, tx.yaw, rx.x - tx.x, rx.y - tx.y, float(rx.z) - float(tx.z), rx.pitch, rx.roll, rx.yaw]], dtype=np.float64)
            if self._tempest:
                d = TempestData(dp.system, one(dp.lineNumber), one(dp.fiducial), one(dp.x), one(dp.y), one(z_in), one(dp.elevation), geometry,
                                np.asarray(dp.secondary_field, dtype=np.float64)[None], np.asarray(dp.primary_field, dtype=np.float64)[None])
                d.additive_error = np.asarray(dp.additive_error, dtype=np.float64)[None]
                return d
            return TdemData(dp.system, one(dp.lineNumber), one(dp.fiducial), one(dp.x), one(dp.y), one(z_in), one(dp.elevation), geometry,
                            np.asarray(dp.data, dtype=np.float64)[None])
        from .dataset import FdemData
        d = FdemData(dp.system)
        d.lineNumber, d.fiducial, d.x, d.y, d.elevation = one(dp.lineNumber), one(dp.fiducial), one(dp.x), one(dp.y), one(dp.elevation)
        d.z, d.data = one(getattr(dp, "z_input", dp.z)), np.asarray(dp.data, dtype=np.float64)[None]
        return d

    def createHdf(self, parent, add_axis=None):
        """Lay out `parent` for a line of soundings (`add_axis`: their number, or their fiducials as Inference2D.createHdf
        passes them): every group, dataset and attribute of the reference's file (geobipy_b200.hdf.create_line)."""
        from . import hdf
        assert self.datapoint is not None, ValueError("Inference needs a datapoint before creating HDF5 files.")
        if add_axis is None:
            raise NotImplementedError("createHdf without add_axis (a file for one sounding, no line axis) is not built: "
                                      "pass add_axis=1 for a one-sounding line")
        n = int(add_axis) if np.ndim(add_axis) == 0 else int(np.size(add_axis))
        hdf.create_line(parent, n, self.options, self._as_dataset(), n_markov_chains=self.n_markov_chains,
                        update_plot_every=self.update_plot_every, interactive_plot=self.interactive_plot,
                        reciprocate_parameter=self.reciprocate_parameter)
        return parent

    def writeHdf(self, parent, index=None):
        """Write this sounding's row (geobipy_b200.hdf.write_line): the chain's posteriors, then the best model / best
        datapoint over the same datasets, as the reference's writeHdf does.  `index` defaults to the position of the
        datapoint's fiducial among the file's sorted fiducials (:1054-1056)."""
        from . import hdf
        assert getattr(self, "_result", None) is not None, "run infer() first"
        if index is None:
            index = int(np.searchsorted(np.asarray(parent["data/fiducial/data"][()]), float(self.datapoint.fiducial)))
        bdp = self.best_datapoint   # (a Tempest line takes the predicted SECONDARY field and adds the primary field itself)
        predicted = bdp.predicted_secondary_field if self._tempest else bdp.predictedData
        hdf.write_line(parent, self._result, self.options, self._as_dataset(), np.asarray(predicted)[None],
                       rows=[int(index)], multiplier=float(self.multiplier))
        return parent

    def _fill(self, r, b):
        o, s = self.options, r["scalars"][b]
        self.halfspace = float(s[_lib.S_HALFSPACE])
        g = ops.posterior_grids(o, self.halfspace)
        self.iteration = int(s[_lib.S_ITER])
        self.burned_in = bool(s[_lib.S_BURNED_IN])
        self.burned_in_iteration = int(s[_lib.S_BURNED_IN_ITER])
        self.best_iteration = int(s[_lib.S_BEST_ITER])
        self.failed = bool(s[_lib.S_FAILED])
        self.data_misfit = float(s[_lib.S_CUR_MISFIT])
        self.prior, self.likelihood = float(s[_lib.S_CUR_PRIOR]), float(s[_lib.S_CUR_LIKELIHOOD])
        self.posterior = self.prior + self.likelihood
        self.best_posterior = float(s[_lib.S_BEST_POSTERIOR])
        self.acceptance_rate = 100.0 * s[_lib.S_N_ACCEPT] / max(self.iteration, 1)
        self.n_forward_evals = int(s[_lib.S_N_FORWARD])

        def model(k, sig, edges):
            k = int(k)
            return Model(RectilinearMesh1D(edges=edges[:k + 1]), sig[:k])
        self.model = model(s[_lib.S_CUR_K], r["cur_sigma"][b], r["cur_edges"][b])
        self.best_model = model(s[_lib.S_BEST_K], r["best_sigma"][b], r["best_edges"][b])
        # the attribute paths Inference1D.writeHdf serialises (:1050-1090): model.values.posterior,
        # model.mesh.nCells.posterior, model.mesh.edges.posterior, datapoint.relative_error / additive_error / z .posterior
        self.model.values.posterior = Histogram(r["hitmap"][b], g["sigma_edges"], g["depth_edges"], log_x=True)
        self.model.mesh.nCells.posterior = Histogram(r["ncells_hist"][b], np.arange(-0.5, o.max_layers + 1.0))
        self.model.mesh.edges.posterior = Histogram(r["edges_hist"][b], g["depth_edges"])
        self.hitmap = self.model.values.posterior
        self.n_cells_posterior, self.edges_posterior = self.model.mesh.nCells.posterior, self.model.mesh.edges.posterior
        self.data_misfit_v = r["misfit_trace"][b]
        self.acceptance_v = r["accept_trace"][b]
        dp = self.datapoint
        if o.n_systems > 1:  # one histogram per system, as DataPoint.set_relative_error_posterior builds them (:668-680)
            rel_post = [Histogram(r["rel_hist"][b][i], g["rel_edges"][i], log_x=True) for i in range(2)]
            add_post = [Histogram(r["add_hist"][b][i], g["add_edges"][i], log_x=True) for i in range(2)]
            dp.relative_error = StatArray([s[_lib.S_CUR_REL], s[_lib.S_CUR_REL2]], "Relative error", posterior=rel_post)
            if self._tempest:   # the sampled quantity is the multiplier of the fixed additive levels (one per component)
                dp.additive_error_multiplier = StatArray([s[_lib.S_CUR_ADD], s[_lib.S_CUR_ADD2]], "Multiplier", posterior=add_post)
            else:
                dp.additive_error = StatArray([s[_lib.S_CUR_ADD], s[_lib.S_CUR_ADD2]], "Additive error", posterior=add_post)
            self.best_relative_error = np.asarray([s[_lib.S_BEST_REL], s[_lib.S_BEST_REL2]])
            self.best_additive_error = np.asarray([s[_lib.S_BEST_ADD], s[_lib.S_BEST_ADD2]])
        else:
            rel_post = Histogram(r["rel_hist"][b], g["rel_edges"], log_x=True)
            add_post = Histogram(r["add_hist"][b], g["add_edges"], log_x=True)
            dp.relative_error = StatArray([s[_lib.S_CUR_REL]], "Relative error", posterior=rel_post)
            dp.additive_error = StatArray([s[_lib.S_CUR_ADD]], "Additive error", posterior=add_post)
            self.best_relative_error, self.best_additive_error = float(s[_lib.S_BEST_REL]), float(s[_lib.S_BEST_ADD])
        self.relative_error_posterior, self.additive_error_posterior = rel_post, add_post
        z_post = None
        if o.solve_height:  # solve_z: datapoint.z / best_datapoint.z and datapoint.z.posterior (Point.py:1013-1025)
            # the bins are relative to the centre of the height prior, which reset() re-centres on the sampled height
            # (Inference1D.py:984-994): the kernel reports it (S_HEIGHT_REF); dp.z_input keeps the height handed in
            z_ref = float(s[_lib.S_HEIGHT_REF])
            z_post = Histogram(r["height_hist"][b], z_ref + np.linspace(-o.max_height_change, o.max_height_change, o.n_err_bins + 1))
            self.height_posterior = z_post
            self.best_height = float(s[_lib.S_BEST_HEIGHT])
            if self._tdem:   # the transmitter moves, the receiver keeps its offset (Loop_pair.Geometry, Loop_pair.py:62-78)
                dz = float(s[_lib.S_CUR_HEIGHT]) - float(dp.transmitter.z)
                dp.transmitter.z = StatArray(float(dp.transmitter.z) + dz, "Height", "m", posterior=z_post)
                dp.receiver.z = float(dp.receiver.z) + dz
            else:
                dp.z = StatArray(float(s[_lib.S_CUR_HEIGHT]), "Height", "m", posterior=z_post)
        # best_datapoint: the datapoint at the highest-posterior state (errors, height, predicted data of best_model)
        import copy
        bdp = copy.copy(dp)
        if self._tdem:
            bdp.predicted_secondary_field = dp.predicted_secondary_field.copy()
        else:
            bdp.predictedData = dp.predictedData.copy()
        bdp.relative_error = StatArray(np.atleast_1d(self.best_relative_error), "Relative error")
        if self._tempest:
            bdp.additive_error_multiplier = StatArray(np.atleast_1d(self.best_additive_error), "Multiplier")
        else:
            bdp.additive_error = StatArray(np.atleast_1d(self.best_additive_error), "Additive error")
        if o.solve_height:
            if self._tdem:
                bdp.transmitter, bdp.receiver = copy.copy(dp.transmitter), copy.copy(dp.receiver)
                dzb = self.best_height - float(dp.transmitter.z)
                bdp.transmitter.z, bdp.receiver.z = self.best_height, float(dp.receiver.z) + dzb
            else:
                bdp.z = StatArray(self.best_height, "Height", "m")
        bdp.forward(self.best_model)
        self.best_datapoint = bdp
        dp.forward(self.model)

    def interface_probability(self):
        """edges histogram / sum (Inference2D.interface_probability, Inference2D.py:959-961)."""
        c = self.edges_posterior.counts.astype(np.float64)
        return c / max(c.sum(), 1.0)


def infer_batch(system, data, altitude, seed=0, precision=_lib.PRECISION_F32, device=0, first_index=0,
                outputs=ops.DEFAULT_OUTPUTS, max_iterations=0, **options):
    """Batched replacement of the `Inference3D.infer_serial` loop (Inference3D.py:458-492): all soundings of
    `data` [B, 2F] are inverted concurrently, one warp per sounding.  Returns the dict of posterior arrays
    (`include/geobipy_b200.h` gbp_chain_buffers) plus the options struct used."""
    s = system.c_struct if hasattr(system, "c_struct") else system
    opt = ops.options_from_reference(**options)
    r = ops.rjmcmc_run(s, opt, data, altitude, seed=seed, first_index=first_index, max_iterations=max_iterations,
                       precision=precision, device=device, outputs=outputs)
    r["options"] = opt
    return r
