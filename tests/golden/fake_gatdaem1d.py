"""BUILD CONTAINER ONLY: a stand-in for the absent third-party module `gatdaem1d` (GeoscienceAustralia/ga-aem)
so that the UNMODIFIED reference's time-domain classes (TdemDataPoint, TdemSystem_GAAEM, Loop_pair, Model.Earth)
and its Inference1D can be driven here.  The forward arithmetic behind it is the oracle's restatement
(oracle/tdem1d_oracle.c); what the recordings made through it pin is everything AROUND the forward: the
time-domain error model (TdemDataPoint.std :329-379), the per-system error priors / joint proposals, the
Hessian / gradient assembly with a 45-channel dual-moment Jacobian, priors, likelihood and proposal densities
as the reference's own Python computes them.

Interface reproduced (call sites: classes/system/TdemSystem_GAAEM.py:8-40, .../TD/tdem1d.py:89-154,
classes/model/Model.py:152-159, classes/system/Loop_pair.py:62-78)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle_py as O  # noqa: E402

_DUAL = None


def _dual():
    global _DUAL
    if _DUAL is None:
        _DUAL = (O.make_tdem_system(), O.skytem_definitions())
    return _DUAL


class Earth:
    def __init__(self, conductivity, thickness):
        self.conductivity = np.asarray(conductivity, dtype=np.float64).copy()
        self.thickness = np.asarray(thickness, dtype=np.float64).copy()


class Geometry:
    def __init__(self, tx_height, tx_roll, tx_pitch, tx_yaw, txrx_dx, txrx_dy, txrx_dz, rx_roll, rx_pitch, rx_yaw):
        assert tx_roll == tx_pitch == tx_yaw == rx_roll == rx_pitch == rx_yaw == 0.0
        self.tx_height = float(tx_height)
        self.offset = (float(txrx_dx), float(txrx_dy), float(txrx_dz))


class _Response:
    def __init__(self, sz):
        z = np.zeros_like(sz)
        self.PX = self.PY = self.PZ = np.zeros(1)
        self.SX, self.SY, self.SZ = z, z, sz


class _Components:
    pass


class _GenericSystem:
    """Any single system that is not one of the two SkyTEM moments (the fixed-wing Tempest system: X and Z components of
    the B field, primary field): the oracle's generalised restatement, one oracle system per receiver offset."""
    CONDUCTIVITYDERIVATIVE = 1

    def __init__(self, stmfile):
        sys.path.insert(0, os.path.join(HERE, "..", ".."))
        from geobipy_b200 import tdem
        self.d = tdem.read_stm(stmfile)
        self.comps = O.tdem_components(self.d)
        self.nw = len(self.d["window_start"])
        self.windows = types.SimpleNamespace(centre=0.5 * (np.asarray(self.d["window_start"]) + np.asarray(self.d["window_end"])))
        self.waveform = types.SimpleNamespace()
        self._cache, self._last = {}, None

    def loopRadius(self):
        return float(self.d.get("loop_radius", 0.0))

    def _sys(self, G):
        if G.offset not in self._cache:
            self._cache[G.offset] = O.make_tdem_system([self.d], rx_offset=G.offset)
        return self._cache[G.offset]

    def _split(self, v, G=None):
        """our channel order (x windows, z windows; z up) -> gatdaem1d's SX / SY / SZ (z down: the reference negates it)"""
        r = _Components()
        zero = np.zeros((self.nw,) + np.shape(v)[1:])
        r.SX, r.SY, r.SZ = zero, zero, zero
        o = 0
        for c in self.comps:
            if c == "x":
                r.SX = v[o:o + self.nw]
            else:
                r.SZ = -v[o:o + self.nw]
            o += self.nw
        if G is not None:
            p = O.tdem_primary_field(self.d, G.offset)
            r.PX = r.PY = r.PZ = 0.0
            for c, val in zip(self.comps, p):
                if c == "x":
                    r.PX = val
                else:
                    r.PZ = -val
        return r

    def _thk(self, E):
        return np.r_[E.thickness, 1.0]

    def forwardmodel(self, G, E):
        self._last = (G, E)
        return self._split(O.tdem_forward(self._sys(G), G.tx_height, E.conductivity, self._thk(E)), G)

    def fm_dlogc(self, G, E):
        self._last = (G, E)
        out = O.tdem_forward(self._sys(G), G.tx_height, E.conductivity, self._thk(E))
        J = self._split(O.tdem_sensitivity(self._sys(G), G.tx_height, E.conductivity, self._thk(E)))
        return self._split(out, G), J.SX.T, np.zeros_like(J.SX.T), J.SZ.T

    def derivative(self, dtype, layer):
        G, E = self._last
        J = O.tdem_sensitivity(self._sys(G), G.tx_height, E.conductivity, self._thk(E))
        return self._split(J[:, layer - 1] / E.conductivity[layer - 1])   # d/d sigma: the reference re-multiplies by sigma


class TDAEMSystem:
    CONDUCTIVITYDERIVATIVE = 1

    def __init__(self, stmfile):
        text = open(stmfile).read()
        first = [ln.split("=")[1].strip().lower() for ln in text.splitlines() if ln.strip().startswith("OutputType")]
        self._generic = None
        if first and first[0] == "b":     # not a SkyTEM moment: the generalised restatement (Tempest)
            g = self._generic = _GenericSystem(stmfile)
            self.windows, self.waveform = g.windows, g.waveform
            self.loopRadius, self.forwardmodel, self.fm_dlogc, self.derivative = g.loopRadius, g.forwardmodel, g.fm_dlogc, g.derivative
            return
        base, self._radius = None, 0.0
        for line in open(stmfile):
            if "BaseFrequency" in line:
                base = float(line.split("=")[1])
            if "ModellingLoopRadius" in line and not line.strip().startswith("//"):
                self._radius = float(line.split("=")[1])
        tsys, defs = _dual()
        self._index = [i for i, d in enumerate(defs) if d["base_frequency"] == base][0]
        d = defs[self._index]
        o = int(sum(tsys.n_win[:self._index]))
        self._slice = slice(o, o + tsys.n_win[self._index])
        self.windows = types.SimpleNamespace(centre=0.5 * (np.asarray(d["window_start"]) + np.asarray(d["window_end"])))
        self.waveform = types.SimpleNamespace()
        self._last = None

    def loopRadius(self):
        return self._radius

    def _thk(self, E):
        return np.r_[E.thickness, 1.0]

    def forwardmodel(self, G, E):
        tsys, _ = _dual()
        assert G.offset == (-13.0, 0.0, 2.0)
        self._last = (G, E)
        out = O.tdem_forward(tsys, G.tx_height, E.conductivity, self._thk(E))
        return _Response(-out[self._slice])          # the reference negates the z component

    def fm_dlogc(self, G, E):
        tsys, _ = _dual()
        self._last = (G, E)
        out = O.tdem_forward(tsys, G.tx_height, E.conductivity, self._thk(E))
        J = O.tdem_sensitivity(tsys, G.tx_height, E.conductivity, self._thk(E))[self._slice]
        z = np.zeros_like(J.T)
        return _Response(-out[self._slice]), z, z, -J.T

    def derivative(self, dtype, layer):
        tsys, _ = _dual()
        G, E = self._last
        J = O.tdem_sensitivity(tsys, G.tx_height, E.conductivity, self._thk(E))[self._slice]
        return _Response(-J[:, layer - 1] / E.conductivity[layer - 1])   # d/d sigma: the reference re-multiplies by sigma


def install():
    m = types.ModuleType("gatdaem1d")
    m.Earth, m.Geometry, m.TDAEMSystem = Earth, Geometry, TDAEMSystem
    sys.modules["gatdaem1d"] = m
