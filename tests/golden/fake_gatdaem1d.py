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
        assert (txrx_dx, txrx_dy, txrx_dz) == (-13.0, 0.0, 2.0)
        self.tx_height = float(tx_height)


class _Response:
    def __init__(self, sz):
        z = np.zeros_like(sz)
        self.PX = self.PY = self.PZ = np.zeros(1)
        self.SX, self.SY, self.SZ = z, z, sz


class TDAEMSystem:
    CONDUCTIVITYDERIVATIVE = 1

    def __init__(self, stmfile):
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
