"""Prototype (build container only): accurate TDEM forward vs the reference's SkyTEM known-answer CSVs."""
import sys, os
import numpy as np
from scipy.special import j0, j1
from scipy.interpolate import CubicSpline

MU0 = 4e-7 * np.pi
DATA = "/root/reference/tests/data_checks"
MODELS = {'glacial': [1e-2, 1e-1, 0.03333333], 'saline_clay': [1e-2, 1e-1, 1.], 'resistive_dolomites': [2e-2, 2e-3, 2e-2],
          'resistive_basement': [1e-2, 1e-1, 1e-4], 'coastal_salt_water': [1., 1e-2, 5e-2], 'ice_over_salt_water': [1e-4, 1e-2, 1]}

HM = dict(base=30.0, dig=491520.0,
          wt=np.array([-4.00E-03, -3.91E-03, -3.81E-03, -3.72E-03, -3.68E-03, -2.30E-03, -1.01E-03, 0.0, 3.25E-06, 1.00E-04, 2.02E-04, 2.82E-04, 3.08E-04, 3.13E-04, 3.15E-04, 3.17E-04, 3.19E-04, 0.012666667]),
          wa=np.array([0, 3.17E-01, 6.30E-01, 8.79E-01, 9.61E-01, 9.74E-01, 9.88E-01, 1.0, 9.91E-01, 7.02E-01, 3.78E-01, 1.16E-01, 2.79E-02, 1.21E-02, 6.61E-03, 3.03E-03, 0, 0]),
          win=np.array([[3.796E-04, 3.872E-04], [3.876E-04, 3.972E-04], [3.976E-04, 4.102E-04], [4.106E-04, 4.262E-04], [4.266E-04, 4.462E-04], [4.466E-04, 4.712E-04], [4.716E-04, 5.022E-04], [5.026E-04, 5.422E-04], [5.426E-04, 5.932E-04], [5.936E-04, 6.562E-04], [6.566E-04, 7.372E-04], [7.376E-04, 8.382E-04], [8.386E-04, 9.652E-04], [9.656E-04, 1.126E-03], [1.127E-03, 1.328E-03], [1.329E-03, 1.583E-03], [1.584E-03, 1.905E-03], [1.906E-03, 2.311E-03], [2.312E-03, 2.822E-03], [2.823E-03, 3.468E-03], [3.469E-03, 4.260E-03], [4.261E-03, 5.228E-03], [5.229E-03, 6.413E-03], [6.414E-03, 7.865E-03], [7.866E-03, 9.641E-03], [9.642E-03, 1.182E-02]]))
LM = dict(base=210.0, dig=3440640.0,
          wt=np.array([-8.00E-04, -7.65E-04, -6.28E-04, -4.35E-04, -9.22E-05, 0.0, 2.20E-07, 4.90E-07, 1.09E-06, 1.69E-06, 3.31E-06, 3.90E-06, 4.47E-06, 5.50E-06, 6.58E-06, 7.27E-06, 8.01E-06, 9.68E-06, 1.17E-05, 1.46E-05, 1.581E-03]),
          wa=np.array([0, 6.34E-02, 2.50E-01, 4.75E-01, 8.90E-01, 1.0, 9.97E-01, 9.80E-01, 9.10E-01, 8.16E-01, 5.37E-01, 4.47E-01, 3.70E-01, 2.56E-01, 1.68E-01, 1.26E-01, 9.08E-02, 4.08E-02, 1.30E-02, 0, 0]),
          win=np.array([[1.828E-05, 2.285E-05], [2.328E-05, 2.885E-05], [2.928E-05, 3.685E-05], [3.728E-05, 4.685E-05], [4.728E-05, 5.985E-05], [6.027E-05, 7.587E-05], [7.627E-05, 9.587E-05], [9.627E-05, 1.209E-04], [1.213E-04, 1.519E-04], [1.523E-04, 1.919E-04], [1.923E-04, 2.429E-04], [2.433E-04, 3.059E-04], [3.063E-04, 3.869E-04], [3.873E-04, 4.879E-04], [4.883E-04, 6.149E-04], [6.153E-04, 7.759E-04], [7.763E-04, 9.779E-04], [9.783E-04, 1.233E-03], [1.233E-03, 1.555E-03]]))
FILT = [(300000.0, 1), (210000.0, 2)]


def rte(lam, omega, sigma, thick):
    """TE reflection coefficient of a layered half-space seen from the air; lam [n], scalar omega."""
    L = len(sigma)
    u = [np.sqrt(lam ** 2 + 1j * omega * MU0 * s) for s in sigma]
    Y = u[L - 1]                       # admittance ~ u (common factor 1/(i w mu) dropped)
    for k in range(L - 2, -1, -1):
        t = np.tanh(u[k] * thick[k])
        Y = u[k] * (Y + u[k] * t) / (u[k] + Y * t)
    return (lam - Y) / (lam + Y)


def Bz_secondary(freqs, sigma, thick, h_tx, h_rx, r, nlam=400, loop_radius=0.0, lam_nodes=None):
    """mu0 * Hz secondary of a unit vertical magnetic dipole, e^{+iwt}, z up positive."""
    zh = h_tx + h_rx
    out = np.zeros(len(freqs), complex)
    if lam_nodes is None:
        loglam = np.linspace(np.log(1e-7), np.log(60.0 / zh), nlam)
    else:
        loglam = lam_nodes
    lam = np.exp(loglam)
    dl = loglam[1] - loglam[0]
    w = np.full(len(lam), dl); w[0] *= 0.5; w[-1] *= 0.5
    geom = lam ** 3 * np.exp(-lam * zh) * j0(lam * r)        # lam^2 dlam = lam^3 dloglam
    if loop_radius > 0:
        x = lam * loop_radius
        geom = geom * 2 * j1(x) / x
    for i, f in enumerate(freqs):
        R = rte(lam, 2 * np.pi * f, sigma, thick)
        out[i] = MU0 / (4 * np.pi) * np.sum(w * geom * R)
    return out


def butter(f, fc, order):
    s = 1j * f / fc
    if order == 1:
        return 1.0 / (1 + s)
    if order == 2:
        return 1.0 / (s * s + np.sqrt(2) * s + 1)
    raise ValueError


def window_operator(sysd, fnodes, nyq=None, filters=FILT, spline_bc='not-a-knot'):
    """A[w, n] complex such that window_w = sum_n 2 Re(A[w,n] S(f_n)), odd harmonics f_n up to nyq."""
    T = 1.0 / sysd['base']
    if nyq is None:
        nyq = sysd['dig'] / 2
    n = np.arange(1, int(nyq / sysd['base']) + 1, 2)
    f = n * sysd['base']
    w = 2 * np.pi * f
    wt, wa = sysd['wt'], sysd['wa']
    slopes = np.diff(wa) / np.diff(wt)
    # Fourier coefficient of dI/dt over the full period; second half period is minus the first => x2 for odd n
    d = np.zeros(len(f), complex)
    for j, s in enumerate(slopes):
        if s == 0:
            continue
        d += s * (np.exp(-1j * w * wt[j]) - np.exp(-1j * w * wt[j + 1])) / (1j * w)
    d *= 2.0 / T
    F = np.ones(len(f), complex)
    for fc, o in filters:
        F *= butter(f, fc, o)
    ta, tb = sysd['win'][:, 0:1], sysd['win'][:, 1:2]
    A = d * F * (np.exp(1j * w * tb) - np.exp(1j * w * ta)) / (1j * w * (tb - ta))
    return f, A


def forward(sysd, sigma, thick, h_tx=30.0, dz=2.0, r=13.0, fpd=5, loop_radius=0.0, exact=False, nlam=400, lam_nodes=None, nyq=None, filters=FILT):
    f, A = window_operator(sysd, None, nyq=nyq, filters=filters)
    if exact:
        fn = np.exp(np.linspace(np.log(f[0] * 0.999), np.log(f[-1] * 1.001), 40 * 6))
    else:
        nd = np.log10(f[-1] / f[0])
        nn = int(np.ceil(nd * fpd)) + 1
        fn = f[0] * 10 ** (np.arange(nn) / fpd)
    S = Bz_secondary(fn, sigma, thick, h_tx, h_tx + dz, r, loop_radius=loop_radius, nlam=nlam, lam_nodes=lam_nodes)
    lf = np.log10(fn)
    Sf = CubicSpline(lf, S.real)(np.log10(f)) + 1j * CubicSpline(lf, S.imag)(np.log10(f))
    return 2 * np.real(A @ Sf)


def load(name):
    a = np.loadtxt(os.path.join(DATA, "skytem_%s_clean.csv" % name), delimiter=",", skiprows=1)
    return a


if __name__ == "__main__":
    zwedge = np.linspace(50.0, 1.0, 79)
    zdeep = np.linspace(75.0, 500.0, 79)
    for name in MODELS:
        a = load(name)
        for i in (0, 40, 78):
            sig = np.array(MODELS[name]); thick = np.array([zwedge[i], zdeep[i] - zwedge[i]])
            ref = a[i, 15:]
            for kw in (dict(exact=True), dict(exact=True, loop_radius=10.416), dict(fpd=5)):
                p = -np.r_[forward(HM, sig, thick, **kw), forward(LM, sig, thick, **kw)]
                e = p / ref - 1
                print(name, i, kw, "HM max %.2e LM max %.2e" % (np.abs(e[:26]).max(), np.abs(e[26:]).max()), np.round(e[[0, 10, 25, 26, 35, 44]], 4))
