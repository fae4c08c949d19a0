"""Host-side mirror of the reference's time-domain classes on the per-sounding inference path.

Mirrored (same names, argument meaning and error behaviour):
  TdemSystem    classes/system/TdemSystem_GAAEM.py (the wrapper of gatdaem1d.TDAEMSystem): `read(stm)`, `off_time`,
                `nTimes`, `components`
  TdemLoop      classes/system/CircularLoop.py as used by TdemDataPoint(transmitter_loop=, receiver_loop=)
  TdemDataPoint classes/data/datapoint/TdemDataPoint.py: data = secondary_field [windows of system 0, system 1, ...];
                `std` with per-system errors and the (t / 1 ms)^-1/2 additive scaling (:329-379); `forward`,
                `sensitivity`, `fm_dlogc` (:997-1055)
  TdemData      classes/data/dataset/TdemData.py read_csv: the column layout of the reference's skytem_*.csv
                (Line_number, Fiducial, Easting, Northing, Height, Elevation, tx_pitch .. rx_yaw, then the windows)
Everything numerical happens behind the C-ABI (ops.py / include/geobipy_b200.h gbp_tdem_*).
"""
import numpy as np

from . import _lib, ops
from .api import Model

__all__ = ["read_stm", "TdemSystem", "TdemLoop", "TdemDataPoint", "Tempest_datapoint", "TdemData", "TempestData"]


def read_stm(filename):
    """Parse a GA-AEM .stm file (block structured `Key = value` text; e.g. SkytemHM.stm) into the dictionary
    `ops.make_tdem_system_struct` consumes."""
    d, mode, wave, win = {}, None, [], []
    with open(filename) as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("//"):
                continue
            low = s.lower()
            if low.startswith("waveformcurrent begin"):
                mode = "wave"
            elif low.startswith("waveformcurrent end") or low.startswith("windowtimes end"):
                mode = None
            elif low.startswith("windowtimes begin"):
                mode = "win"
            elif mode in ("wave", "win"):
                a = s.split()
                (wave if mode == "wave" else win).append((float(a[0]), float(a[1])))
            elif "=" in s:
                k, v = (x.strip() for x in s.split("=", 1))
                d[k] = v
    assert str(d.get("Type", "Time Domain")).lower().startswith("time"), ValueError("not a time-domain system file")
    assert len(wave) >= 2 and len(win) >= 1, ValueError("system file has no waveform / windows: " + filename)
    cut = [float(x) for x in d.get("CutOffFrequency", "").split()]
    order = [int(float(x)) for x in d.get("Order", "").split()]
    return dict(
        name=d.get("Name", ""), base_frequency=float(d["BaseFrequency"]),
        digitising_frequency=float(d["WaveformDigitisingFrequency"]),
        n_turns=float(d.get("NumberOfTurns", 1)), peak_current=float(d.get("PeakCurrent", 1)),
        loop_area=float(d.get("LoopArea", 1)),
        waveform_time=[w[0] for w in wave], waveform_current=[w[1] for w in wave],
        window_start=[w[0] for w in win], window_end=[w[1] for w in win],
        filter_cutoff=cut, filter_order=order, loop_radius=float(d.get("ModellingLoopRadius", 0.0)),
        output_type=d.get("OutputType", "dB/dt"), x_scaling=float(d.get("XOutputScaling", 0)),
        y_scaling=float(d.get("YOutputScaling", 0)), z_scaling=float(d.get("ZOutputScaling", 1)),
        frequencies_per_decade=int(float(d.get("FrequenciesPerDecade", 5))),
        n_abscissae=int(float(d.get("NumberOfAbsiccaInHankelTransformEvaluation", 21))))


def format_stm(d):
    """A GA-AEM .stm text (list of lines) for a parsed description: `read_stm` of it gives the description back."""
    L = ["System Begin\n", "\tName = %s\n" % d.get("name", ""), "\tType = Time Domain\n", "\tTransmitter Begin\n",
         "\t\tNumberOfTurns = %g\n" % d.get("n_turns", 1), "\t\tPeakCurrent   = %g\n" % d.get("peak_current", 1),
         "\t\tLoopArea      = %g\n" % d.get("loop_area", 1), "\t\tBaseFrequency = %r\n" % float(d["base_frequency"]),
         "\t\tWaveformDigitisingFrequency = %r\n" % float(d["digitising_frequency"]), "\t\tWaveFormCurrent Begin\n"]
    L += ["%r\t%r\n" % (float(t), float(c)) for t, c in zip(d["waveform_time"], d["waveform_current"])]
    L += ["\t\tWaveFormCurrent End\n", "\tTransmitter End\n", "\tReceiver Begin\n",
          "\t\tNumberOfWindows = %d\n" % len(d["window_start"]), "\t\tWindowWeightingScheme = AreaUnderCurve\n",
          "\t\tWindowTimes Begin\n"]
    L += ["%r\t%r\n" % (float(a), float(b)) for a, b in zip(d["window_start"], d["window_end"])]
    L += ["\t\tWindowTimes End\n"]
    if len(d.get("filter_cutoff", [])):
        L += ["\t\tLowPassFilter Begin\n", "\t\t\tCutOffFrequency = %s\n" % " ".join("%r" % float(x) for x in d["filter_cutoff"]),
              "\t\t\tOrder           = %s\n" % " ".join("%d" % int(x) for x in d["filter_order"]), "\t\tLowPassFilter End\n"]
    L += ["\tReceiver End\n", "\tForwardModelling Begin\n", "\t\tModellingLoopRadius = %r\n" % float(d.get("loop_radius", 0.0)),
          "\t\tOutputType = %s\n" % d.get("output_type", "dB/dt"), "\t\tXOutputScaling = %g\n" % d.get("x_scaling", 0),
          "\t\tYOutputScaling = %g\n" % d.get("y_scaling", 0), "\t\tZOutputScaling = %g\n" % d.get("z_scaling", 1),
          "\t\tSecondaryFieldNormalisation  =  none\n", "\t\tFrequenciesPerDecade = %d\n" % d.get("frequencies_per_decade", 5),
          "\t\tNumberOfAbsiccaInHankelTransformEvaluation = %d\n" % d.get("n_abscissae", 21), "\tForwardModelling End\n",
          "System End\n"]
    return L


class TdemSystem:
    """One time-domain system (TdemSystem_GAAEM): built from an .stm file or a parsed description."""

    def __init__(self, system_filename=None, definition=None):
        if definition is None:
            import os
            assert system_filename is not None and os.path.exists(system_filename), \
                'Could not open file: ' + str(system_filename)   # TdemSystem_GAAEM.py:29
            definition = read_stm(system_filename)
        self.definition = definition
        self.filename = system_filename
        self._lines = None
        if system_filename is not None:
            with open(system_filename) as f:
                self._lines = f.readlines()
        self.off_time = 0.5 * (np.asarray(definition["window_start"]) + np.asarray(definition["window_end"]))
        assert np.min(np.diff(self.off_time)) > 0.0 if self.off_time.size > 1 else True, ValueError(
            "Receiver window times must monotonically increase for system " + str(system_filename))
        # the components with a non-zero Output Scaling, x before z (TdemSystem_GAAEM / TdemDataPoint.forward :1008-1016)
        self.components = [c for c, k, dflt in (('x', "x_scaling", 0.0), ('z', "z_scaling", 1.0)) if float(definition.get(k, dflt)) != 0.0]

    @classmethod
    def read(cls, system_filename):
        return cls(system_filename)

    @property
    def stm_lines(self):
        """The .stm file line by line - what the reference keeps as `TdemSystem_GAAEM.string` and stores in its result
        files (`toHdf`, TdemSystem_GAAEM.py:114-118).  A system built from a parsed description is written out again."""
        return self._lines if self._lines is not None else format_stm(self.definition)

    @property
    def nTimes(self):
        return self.off_time.size

    @property
    def n_components(self):
        return len(self.components)

    @property
    def isGA(self):
        return True


class TdemLoop:
    """Transmitter / receiver loop of a time-domain datapoint (CircularLoop): position offsets and attitude.
    Only zero pitch / roll / yaw are supported on the GPU path."""

    def __init__(self, x=0.0, y=0.0, z=0.0, elevation=0.0, orientation='z', moment=1.0, pitch=0.0, roll=0.0, yaw=0.0,
                 radius=None, **kwargs):
        self.x, self.y, self.z = float(x), float(y), float(z)
        self.orientation, self.moment, self.radius = orientation, float(moment), radius
        self.pitch, self.roll, self.yaw = float(pitch), float(roll), float(yaw)


class TdemDataPoint:
    """One time-domain sounding with 1-2 systems (TdemDataPoint.py).  Data are dBz/dt window averages in
    V/(A m^4), system 0's windows first."""

    def __init__(self, x=0.0, y=0.0, z=0.0, elevation=0.0, primary_field=None, secondary_field=None, relative_error=None,
                 additive_error=None, std=None, predicted_primary_field=None, predicted_secondary_field=None,
                 system=None, transmitter_loop=None, receiver_loop=None, lineNumber=0.0, fiducial=0.0,
                 precision=_lib.PRECISION_F64, **kwargs):
        if isinstance(system, (str, TdemSystem)):
            system = [system]
        assert system is not None and 1 <= len(system) <= _lib.TD_MAXSYS, ValueError("1 or 2 systems per datapoint")
        self.system = [s if isinstance(s, TdemSystem) else TdemSystem(s) for s in system]
        self.x, self.y, self.z, self.elevation = float(x), float(y), float(z), float(elevation)
        self.lineNumber, self.fiducial = lineNumber, fiducial
        self.transmitter = transmitter_loop if transmitter_loop is not None else TdemLoop(z=self.z)
        self.receiver = receiver_loop if receiver_loop is not None else TdemLoop(x=-13.0, z=self.z + 2.0)
        for lp in (self.transmitter, self.receiver):
            assert lp.pitch == 0.0 and lp.roll == 0.0 and lp.yaw == 0.0, NotImplementedError(
                "loop pitch / roll / yaw are not supported on the GPU path")
        n = self.nChannels
        self.secondary_field = np.zeros(n) if secondary_field is None else np.asarray(secondary_field, np.float64).copy()
        self.predicted_secondary_field = (np.zeros(n) if predicted_secondary_field is None
                                          else np.asarray(predicted_secondary_field, np.float64).copy())
        assert self.secondary_field.size == n, ValueError("secondary_field must have %d entries" % n)
        self._std = np.full(n, 0.01) if std is None else np.asarray(std, np.float64).copy()
        ns = self.nSystems
        self.relative_error = np.full(ns, 0.01) if relative_error is None else np.asarray(relative_error, np.float64).reshape(-1)
        self.additive_error = np.zeros(ns) if additive_error is None else np.asarray(additive_error, np.float64).reshape(-1)
        assert self.relative_error.size == ns, ValueError("relative_error must be a list of size equal to the number of systems {}".format(ns))
        assert self.additive_error.size == ns, ValueError("additive_error must be a list of size equal to the number of systems {}".format(ns))
        self._use_errors = relative_error is not None
        self.sensitivity_matrix = None
        self.precision = precision
        self._struct = None

    # -- layout
    @property
    def nSystems(self):
        return len(self.system)

    @property
    def nTimes(self):
        return np.asarray([s.nTimes for s in self.system])

    @property
    def nChannels(self):
        return int((self.nTimes * self.n_components).sum())

    @property
    def components(self):
        return list(self.system[0].components)

    @property
    def n_components(self):
        return len(self.components)

    def off_time(self, system=0):
        return self.system[system].off_time

    def _systemIndices(self, system=0):
        o = np.r_[0, np.cumsum(self.nTimes * self.n_components)]
        return np.s_[o[system]:o[system + 1]]

    def _component_indices(self, component=0, system=0):
        """Channels of component `component` (position in `components`) of a system (TdemDataPoint._component_indices)."""
        o = np.r_[0, np.cumsum(self.nTimes * self.n_components)][system] + component * self.nTimes[system]
        return np.s_[o:o + self.nTimes[system]]

    @property
    def data(self):
        return self.secondary_field

    @property
    def predictedData(self):
        return self.predicted_secondary_field

    @property
    def c_struct(self):
        if self._struct is None:
            off = (self.receiver.x - self.transmitter.x, self.receiver.y - self.transmitter.y, self.receiver.z - self.transmitter.z)
            self._struct = ops.make_tdem_survey_struct([s.definition for s in self.system], off)
        return self._struct

    # -- EmDataPoint.active :44-56
    @property
    def active(self):
        return self.secondary_field > 0.0

    @property
    def n_active_channels(self):
        return int(self.active.sum())

    # -- TdemDataPoint.std :329-379
    @property
    def std(self):
        if self._use_errors:
            assert np.all(self.relative_error > 0.0), ValueError('relative_error must be > 0.0')
            for i in range(self.nSystems):
                ic = self._systemIndices(i)
                rel = self.relative_error[i] * self.secondary_field[ic]
                add = np.exp(np.log(self.additive_error[i]) - 0.5 * (np.log(self.off_time(i)) - np.log(1e-3)))
                self._std[ic] = np.sqrt(rel ** 2 + add ** 2)
        return self._std

    def initialize(self, **kwargs):
        self.relative_error = np.asarray(kwargs['initial_relative_error'], np.float64).reshape(-1)
        self.additive_error = np.asarray(kwargs['initial_additive_error'], np.float64).reshape(-1)
        self._use_errors = True

    @property
    def deltaD(self):
        return self.predicted_secondary_field - self.secondary_field

    def _model_arrays(self, mod):
        assert isinstance(mod, Model), TypeError("Invalid model class {} for forward modeling [1D]".format(type(mod)))
        assert np.isinf(mod.mesh.edges[-1]), ValueError("mod.edges must have last entry be infinity for forward modelling.")
        alt = self.transmitter.z - mod.mesh.edges[0]
        assert self.z >= mod.mesh.edges[0] and alt > 0.0, "Sensor altitude must be above the top of the model"  # tdem1d.py:32
        L = mod.nCells
        assert 1 <= L <= _lib.MAXL, ValueError("1..%d layers supported" % _lib.MAXL)
        return (np.asarray([L], np.int32), mod.values.reshape(1, L).astype(np.float64),
                mod.mesh.widths.reshape(1, L).astype(np.float64), np.asarray([alt]))

    def forward(self, mod):
        """Fill predicted_secondary_field (and predicted_primary_field: one value per system and component) from a 1-D
        layered model (TdemDataPoint.forward :997-1022)."""
        nl, s, t, a = self._model_arrays(mod)
        self.predicted_secondary_field[:] = ops.forward(self.c_struct, nl, s, t, a, precision=self.precision)[0]
        self.predicted_primary_field = ops.tdem_primary_field(self.c_struct)

    def sensitivity(self, mod, ix=None, model_changed=False):
        """d(predicted)/d ln(sigma) [nChannels, nCells] (TdemDataPoint.sensitivity :1024, gaTdem1dsen's sigma scaling)."""
        nl, s, t, a = self._model_arrays(mod)
        _, J = ops.forward(self.c_struct, nl, s, t, a, precision=self.precision, sensitivity=True)
        self.sensitivity_matrix = J[0] if ix is None else J[0][:, ix]
        return self.sensitivity_matrix

    def fm_dlogc(self, mod):
        nl, s, t, a = self._model_arrays(mod)
        p, J = ops.forward(self.c_struct, nl, s, t, a, precision=self.precision, sensitivity=True)
        self.predicted_secondary_field[:] = p[0]
        self.sensitivity_matrix = J[0]

    # -- DataPoint.data_misfit :502-525, likelihood :491-500
    def data_misfit(self):
        a = self.active
        return float(np.sum((self.deltaD[a] / self.std[a]) ** 2))

    def likelihood(self, log=True):
        a = self.active
        var = self.std[a] ** 2
        ll = -0.5 * a.sum() * np.log(2.0 * np.pi) - 0.5 * np.sum(np.log(var)) - 0.5 * np.sum(self.deltaD[a] ** 2 / var)
        return float(ll) if log else float(np.exp(ll))


class Tempest_datapoint(TdemDataPoint):
    """A fixed-wing Tempest sounding (classes/data/datapoint/Tempest_datapoint.py): X and Z components of the B field in fT;
    `data` / `predictedData` are secondary + primary field per component (:107-127), the error model has one relative
    error per component and an additive error per channel (:141-176).  Forward and Jacobian run on the GPU; the sampler
    for this datapoint type (its error model, the pitch / offset unknowns of tempest_options) is not built."""

    def __init__(self, *args, primary_field=None, additive_error_multiplier=None, **kwargs):
        super().__init__(*args, **kwargs)
        n = self.nSystems * self.n_components
        self.primary_field = np.zeros(n) if primary_field is None else np.asarray(primary_field, np.float64).reshape(n)
        self.predicted_primary_field = np.zeros(n)
        self.relative_error = np.full(n, 0.01)
        self.additive_error = np.zeros(self.nChannels)
        self.additive_error_multiplier = np.ones(n) if additive_error_multiplier is None else np.asarray(additive_error_multiplier, np.float64).reshape(n)

    @property
    def data(self):
        out = self.secondary_field.copy()
        for i in range(self.n_components):
            ic = self._component_indices(i, 0)
            out[ic] = self.secondary_field[ic] + self.primary_field[i]
        return out

    @property
    def predictedData(self):
        out = self.predicted_secondary_field.copy()
        for i in range(self.n_components):
            ic = self._component_indices(i, 0)
            out[ic] = self.predicted_secondary_field[ic] + self.predicted_primary_field[i]
        return out

    @property
    def std(self):
        assert np.all(self.relative_error > 0.0), ValueError('relative_error must be > 0.0')
        d, out = self.data, np.empty(self.nChannels)
        for j in range(self.n_components):
            ic = self._component_indices(j, 0)
            out[ic] = np.sqrt((self.relative_error[j] * d[ic]) ** 2 + (self.additive_error_multiplier[j] * self.additive_error[ic]) ** 2)
        return out

    def initialize(self, **kwargs):
        """Tempest_datapoint.initialize :274-278 -> DataPoint.initialize :530-532: one relative error per component, the
        additive error of every channel (the options file's initial_additive_error vector), multipliers of 1."""
        n = self.nSystems * self.n_components
        self.relative_error = np.asarray(kwargs['initial_relative_error'], np.float64).reshape(n)
        self.additive_error = np.asarray(kwargs['initial_additive_error'], np.float64).reshape(self.nChannels)
        self.additive_error_multiplier = np.ones(n)

    def survey_struct(self, with_errors=True):
        """gbp_tdem_survey of this datapoint type; with the additive level of every channel it selects the Tempest error
        model of the sampler (include/geobipy_b200.h gbp_tdem_survey.error_model)."""
        off = (self.receiver.x - self.transmitter.x, self.receiver.y - self.transmitter.y, float(self.receiver.z) - float(self.transmitter.z))
        return ops.make_tdem_survey_struct([s.definition for s in self.system], off,
                                           additive_level=self.additive_error if with_errors else None)


class TdemData:
    """A time-domain survey file in the reference's CSV layout (TdemData.read_csv): one row per sounding."""

    GEOMETRY = ("tx_pitch", "tx_roll", "tx_yaw", "txrx_dx", "txrx_dy", "txrx_dz", "rx_pitch", "rx_roll", "rx_yaw")

    def __init__(self, system, line_number, fiducial, x, y, height, elevation, geometry, data):
        self.system = system
        self.line_number, self.fiducial, self.x, self.y = line_number, fiducial, x, y
        self.height, self.elevation, self.geometry, self.data = height, elevation, geometry, data
        n, ns = np.shape(data)[0], len(system)
        self.relative_error, self.additive_error = np.full((n, ns), 0.01), np.zeros((n, ns))     # Data.__init__ :94-95
        self._std = None

    @classmethod
    def read_csv(cls, data_filename, system):
        """`system`: list of .stm file names or TdemSystem objects (one per moment).  Columns: Line_number (or Line),
        Fiducial, Easting (X), Northing (Y), Height, Elevation, the nine geometry columns, then
        sum(nTimes) window columns, system 0 first; non-positive or NaN values are inactive channels."""
        import pandas as pd
        if isinstance(system, (str, TdemSystem)):
            system = [system]
        system = [s if isinstance(s, TdemSystem) else TdemSystem(s) for s in system]
        df = pd.read_csv(data_filename, index_col=False, skipinitialspace=True)   # the reference's parser settings (TdemData.read_csv)
        low = {c.strip().lower(): c for c in df.columns}

        def col(*names, default=None):
            for n in names:
                if n in low:
                    return df[low[n]].to_numpy(dtype=np.float64)
            assert default is not None, ValueError("column %s missing from %s" % (names[0], data_filename))
            return np.full(len(df), default)
        n = sum(s.nTimes * s.n_components for s in system)
        # TdemData._csv_channels (classes/data/dataset/TdemData.py:611-633): the data are the columns whose name contains
        # off_time / x_time / y_time / z_time, those that also contain "err" are their standard deviations; on_time
        # columns and the primary field (px, py, pz) are not data
        dcols, ecols = [], []
        for c in df.columns:
            t = c.strip().lower()
            if t in cls.GEOMETRY or "on_time" in t:
                continue
            if any(x in t for x in ("off_time", "x_time", "y_time", "z_time")):
                (ecols if "err" in t else dcols).append(c)
        assert len(dcols) == n, Exception("Number of off time columns {} in {} does not match total number of times {} in system files".format(
            len(dcols), data_filename, n))
        if ecols:
            assert len(ecols) == len(dcols), Exception(
                "Number of Off time standard deviation estimates does not match number of Off time data columns in file {}".format(data_filename))
        geometry = np.stack([col(g, default=(-13.0 if g == "txrx_dx" else 2.0 if g == "txrx_dz" else 0.0)) for g in cls.GEOMETRY], axis=1)
        self = cls(system, col("line_number", "line"), col("fiducial", "fid"), col("easting", "x"), col("northing", "y"),
                   col("height", "z"), col("elevation", "dtm", default=0.0), geometry,
                   df[dcols].to_numpy(dtype=np.float64))
        # the primary field of every component (columns px / py / pz, sorted by name: TdemData._csv_channels :631-633)
        pcols = sorted(c for c in df.columns if c.strip().lower() in ("px", "py", "pz"))
        self._read_primary(df[pcols].to_numpy(dtype=np.float64) if pcols else None)
        # TdemData.read_csv :520-525: without error columns std = 0.1 * data (kept as std_from_file: the `std` getter
        # recomputes from the per-system errors, as the reference's does)
        self.std = df[ecols].to_numpy(dtype=np.float64) if ecols else self._std_without_columns()
        return self

    def _read_primary(self, primary):
        pass   # a TdemData keeps the secondary field only (TdemData.read_csv :520-525); TempestData overrides

    def _std_without_columns(self):
        return 0.1 * self.data

    @property
    def std(self):
        """TdemData.std (classes/data/dataset/TdemData.py:269-278): recomputed on every access from the per-system
        relative errors (1 % unless set) and additive errors (0); what the file gave stays in `std_from_file`."""
        out = np.empty_like(self.data)
        o = 0
        for i, s_ in enumerate(self.system):
            j = slice(o, o + s_.nTimes)
            out[:, j] = np.sqrt((self.relative_error[:, i][:, None] * self.data[:, j]) ** 2 + (self.additive_error[:, i] ** 2)[:, None])
            o += s_.nTimes
        return out

    @std.setter
    def std(self, values):
        self._std = values

    @property
    def std_from_file(self):
        return getattr(self, "_std", None)

    @property
    def nPoints(self):
        return self.data.shape[0]

    @property
    def nChannels(self):
        return self.data.shape[1]

    def subset(self, idx):
        """The soundings `idx` as a data set of their own (Data.__getitem__ of the reference)."""
        out = TdemData(self.system, self.line_number[idx], self.fiducial[idx], self.x[idx], self.y[idx], self.height[idx],
                       np.asarray(self.elevation)[idx], self.geometry[idx], self.data[idx])
        if self._std is not None:
            out.std = np.asarray(self._std)[idx]
        out.relative_error, out.additive_error = self.relative_error[idx], self.additive_error[idx]
        return out

    # names the survey driver (dataset.Inference3D) shares with FdemData
    @property
    def z(self):
        return self.height

    def _loop(self, dx, dy, dz, att):
        n = self.nPoints
        r = float(self.system[0].definition.get("loop_radius", 0.0))
        return dict(x=self.x + dx, y=self.y + dy, z=self.height + dz, elevation=np.asarray(self.elevation, dtype=np.float64),
                    pitch=self.geometry[:, att], roll=self.geometry[:, att + 1], yaw=self.geometry[:, att + 2],
                    moment=np.ones(n), orientation=np.full(n, 2, dtype=np.int32), radius=np.full(n, r))

    @property
    def transmitter(self):
        """Per-sounding transmitter loops as arrays (TdemData.loop_pair.transmitter, a CircularLoops): position = the
        sounding's, attitude from the tx_* columns, radius = the systems' ModellingLoopRadius, z-directed unit moment."""
        return self._loop(0.0, 0.0, 0.0, 0)

    @property
    def receiver(self):
        g = self.geometry
        return self._loop(g[:, 3], g[:, 4], g[:, 5], 6)

    @property
    def lineNumber(self):
        return self.line_number

    @property
    def c_struct(self):
        return self.survey_struct()

    def survey_struct(self):
        """One gbp_tdem_survey for the whole file: the GPU path needs a constant tx->rx offset and zero attitudes."""
        g = self.geometry
        assert np.all(g[:, [0, 1, 2, 6, 7, 8]] == 0.0), NotImplementedError("loop pitch / roll / yaw are not supported on the GPU path")
        assert np.all(g[:, 3:6] == g[0, 3:6]), NotImplementedError("the transmitter-receiver offset must be constant over the file")
        return ops.make_tdem_survey_struct([s.definition for s in self.system], tuple(g[0, 3:6]))

    def datapoint(self, i):
        """TdemDataPoint of row i (TdemData.datapoint)."""
        g = self.geometry[i]
        tx = TdemLoop(x=self.x[i], y=self.y[i], z=self.height[i])
        rx = TdemLoop(x=self.x[i] + g[3], y=self.y[i] + g[4], z=self.height[i] + g[5])
        return TdemDataPoint(self.x[i], self.y[i], self.height[i], self.elevation[i], secondary_field=self.data[i],
                             system=self.system, transmitter_loop=tx, receiver_loop=rx, lineNumber=self.line_number[i],
                             fiducial=self.fiducial[i])


class TempestData(TdemData):
    """A fixed-wing Tempest survey (classes/data/dataset/TempestData.py): one system, X and Z components of the B field in
    fT.  `secondary_field` [n, C] is what the file's S0X_time* / S0Z_time* columns hold, `primary_field` [n, components]
    its PX / PZ columns (:251-252); `data` = secondary + primary per component (what a Tempest_datapoint inverts,
    Tempest_datapoint.py:107-127).  Errors per COMPONENT: `relative_error` [n, components], `additive_error` [n, C] (the
    additive level of every channel) times `additive_error_multiplier` [n, components] (TempestData.py:68-123)."""

    def __init__(self, system, line_number, fiducial, x, y, height, elevation, geometry, secondary_field, primary_field=None):
        super().__init__(system, line_number, fiducial, x, y, height, elevation, geometry, secondary_field)
        assert len(self.system) == 1, NotImplementedError("a Tempest data set has one system")
        n, nc = self._secondary.shape[0], self.n_components
        self.primary_field = np.zeros((n, nc)) if primary_field is None else np.asarray(primary_field, np.float64).reshape(n, nc)
        # what the reference's reader leaves behind (TempestData.__init__ :59-65 as read_csv returns it): all zero until the
        # inversion's options set them
        self.relative_error = np.zeros((n, nc))
        self.additive_error = np.zeros((n, self.secondary_field.shape[1]))
        self.additive_error_multiplier = np.zeros((n, nc))
        self._std = np.zeros((n, self.secondary_field.shape[1]))

    def _std_without_columns(self):
        return np.zeros_like(self._secondary)   # TempestData.read_csv :255-258 assigns std only when the file has error columns

    def _read_primary(self, primary):
        if primary is not None:
            assert primary.shape[1] == self.n_components, Exception("one primary-field column (PX / PY / PZ) per component")
            self.primary_field = primary

    @property
    def n_components(self):
        return self.system[0].n_components

    @property
    def secondary_field(self):
        return self._secondary

    @property
    def nPoints(self):
        return self._secondary.shape[0]

    # TdemData keeps its channels in `data`; here that name is secondary + primary (TempestData / Tempest_datapoint.data)
    @property
    def data(self):
        nt = self.system[0].nTimes
        return self._secondary + np.repeat(self.primary_field, nt, axis=1)

    @data.setter
    def data(self, values):
        self._secondary = np.asarray(values, dtype=np.float64)

    @property
    def std(self):
        """The data set's std is the getter TempestData inherits from TdemData (classes/data/dataset/TdemData.py:269-278):
        once a relative error is set, every channel of the one system takes the FIRST component's relative error and the
        FIRST channel's additive level; before that, what the file gave (zeros without error columns).  The error model
        the inversion uses is the datapoint's (`datapoint(i).std`, Tempest_datapoint.py:141-176)."""
        if self.relative_error.max() > 0.0:
            self._std = np.sqrt((self.relative_error[:, :1] * self.data) ** 2 + self.additive_error[:, :1] ** 2)
        return self._std

    @std.setter
    def std(self, values):
        self._std = values

    def subset(self, idx):
        out = TempestData(self.system, self.line_number[idx], self.fiducial[idx], self.x[idx], self.y[idx], self.height[idx],
                          np.asarray(self.elevation)[idx], self.geometry[idx], self._secondary[idx], self.primary_field[idx])
        out.std = np.asarray(self._std)[idx]
        out.relative_error, out.additive_error = self.relative_error[idx], self.additive_error[idx]
        out.additive_error_multiplier = self.additive_error_multiplier[idx]
        return out

    def survey_struct(self, additive_level=None):
        """gbp_tdem_survey of the file; with the additive level of every channel (tempest_options' initial_additive_error)
        it selects the Tempest error model of the sampler."""
        g = self.geometry
        assert np.all(g[:, [0, 1, 2, 6, 7, 8]] == 0.0), NotImplementedError("loop pitch / roll / yaw are not supported on the GPU path")
        assert np.all(g[:, 3:6] == g[0, 3:6]), NotImplementedError("the transmitter-receiver offset must be constant over the file")
        return ops.make_tdem_survey_struct([s.definition for s in self.system], tuple(g[0, 3:6]), additive_level=additive_level)

    def datapoint(self, i):
        g = self.geometry[i]
        tx = TdemLoop(x=self.x[i], y=self.y[i], z=self.height[i])
        rx = TdemLoop(x=self.x[i] + g[3], y=self.y[i] + g[4], z=self.height[i] + g[5])
        dp = Tempest_datapoint(self.x[i], self.y[i], self.height[i], self.elevation[i], secondary_field=self._secondary[i],
                               primary_field=self.primary_field[i], system=self.system, transmitter_loop=tx, receiver_loop=rx,
                               lineNumber=self.line_number[i], fiducial=self.fiducial[i])
        if self.relative_error[i].max() > 0.0:    # errors the data set carries go with the datapoint (TdemData.datapoint)
            dp.relative_error, dp.additive_error = self.relative_error[i].copy(), self.additive_error[i].copy()
        return dp
