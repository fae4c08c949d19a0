"""Synthetic RESOLVE soundings (BASELINE.json configs[1]).

Pure NumPy description of the *inputs*; the responses themselves come from the CUDA forward
operator (or, in tests, from the oracle).  The per-sounding generator is
``numpy.random.default_rng(1234 + i)`` so any sounding can be regenerated independently on
any rank.
"""
import numpy as np

__all__ = ["synthetic_sounding", "synthetic_batch", "skytem_noise_std"]


def synthetic_sounding(i, max_depth=150.0, n_channels=12):
    """True model, sensor height and unit-normal noise vector of synthetic sounding ``i``.

    L ~ U{1..10}; interfaces log-uniform in [1, max_depth] m with >= 1 m spacing;
    log10(sigma) ~ U[-3, 0]; height ~ U[25, 45] m.  RESOLVE (BASELINE configs[1]): max_depth 150, 12 channels;
    SkyTEM (configs[3]): max_depth 400, 45 channels.
    Returns (edges[L+1] with edges[0]=0, edges[L]=inf; sigma[L]; height; noise[n_channels]).
    """
    rng = np.random.default_rng(1234 + int(i))
    L = int(rng.integers(1, 11))
    while True:
        e = np.sort(np.exp(rng.uniform(np.log(1.0), np.log(max_depth), L - 1)))
        z = np.r_[0.0, e]
        if L == 1 or np.min(np.diff(z)) >= 1.0:
            break
    sigma = 10.0 ** rng.uniform(-3.0, 0.0, L)
    height = rng.uniform(25.0, 45.0)
    noise = rng.standard_normal(n_channels)
    return np.r_[0.0, e, np.inf], sigma, height, noise


def skytem_noise_std(clean, t_centre, n_win, rel=0.05, add=(2e-14, 2e-13)):
    """Noise model of the synthetic SkyTEM data = the reference's own (TdemDataPoint.std :329-379 with the
    initial errors of skytem_options): sqrt((rel d)^2 + (add_s sqrt(1 ms / t))^2)."""
    a = np.concatenate([np.full(n, v) for n, v in zip(n_win, add)]) * np.sqrt(1e-3 / np.asarray(t_centre))
    return np.sqrt((rel * np.asarray(clean)) ** 2 + a ** 2)


def synthetic_batch(first, count, max_layers=30, max_depth=150.0, n_channels=12):
    """Padded arrays for soundings first .. first+count-1.

    Returns dict(nlayers[int32 B], sigma[B, max_layers], thickness[B, max_layers], height[B], noise[B, C]).
    Unused layer slots hold sigma = 1, thickness = inf.
    """
    nl = np.zeros(count, np.int32)
    sig = np.ones((count, max_layers))
    thk = np.full((count, max_layers), np.inf)
    h = np.zeros(count)
    noise = np.zeros((count, n_channels))
    for j in range(count):
        e, s, hh, n = synthetic_sounding(first + j, max_depth, n_channels)
        L = s.size
        nl[j] = L
        sig[j, :L] = s
        thk[j, :L] = np.diff(e)
        h[j] = hh
        noise[j] = n
    return dict(nlayers=nl, sigma=sig, thickness=thk, height=h, noise=noise)
