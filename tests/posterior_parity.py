"""Shared checker of posterior-model statistics against ensembles of LIVE-REFERENCE chains (tests/golden/ref_chain_*,
ref_tdem_chain_*; tests/golden/make_golden.py chain / tdem_chain).  Test infrastructure, used by the CPU tests of the
oracle and by the -m gpu tests of the fp32 production kernels.

SURVEY.md 8(d) parity list, as it can be stated for this sampler.  10k-iteration chains of the reference mix slowly - its
own chains on the same data differ from each other by tens of conductivity bins, 0.15-0.47 in acceptance rate and 1.75-2.95
in mean layer count (sounding 3) - so every statistic is compared with the ENSEMBLE of reference chains:

  profiles      5 % / 50 % / 95 % conductivity bin per depth cell (Mesh._percentile, classes/mesh/Mesh.py:173-217) of the
                pooled chains under test inside the envelope of the reference chains widened by 2 bins, for a stated
                fraction of the depth cells above the depth of investigation (Inference2D.compute_doi :493-532 applied
                to the pooled reference chains; at most the top 60 m);
  interfaces    most frequent interface cell within 2 depth cells of one of the reference ensemble's three most frequent
                cells or of the most frequent cell of one of its chains (the interface histograms of 10k-iteration chains
                are multi-modal: the 15 reference chains of sounding 1 peak at cells 4 ... 47);
  layers        mean layer count within +-0.5 of the reference ensemble's mean (or twice its standard error, if larger);
  acceptance    acceptance rate within +-5 points of the reference ensemble's mean (or twice its standard error);
  misfit        the reference's post-burn-in misfit does NOT sit on n_active (its chains burn in the first time the misfit
                dips below n_active, Inference1D.py:726, and then hover at 2-12 x n_active with resolve_options' error
                priors), so "centred on n_active" is checked as what it is: the median misfit after iteration 5000 of
                the chains under test lies within the reference chains' range (x / 1.5), the minimum misfit reached dips
                as low as the reference's, and the burned-in fraction agrees within twice its binomial standard error.
"""
import numpy as np


def percentile_bins(hitmap, p):
    """first conductivity bin whose cumulative count reaches p (0..1) of the column total, per depth cell"""
    c = np.cumsum(np.asarray(hitmap, dtype=np.int64), axis=0)
    tot = c[-1]
    return np.array([np.searchsorted(c[:, j], p * tot[j]) for j in range(c.shape[1])])


def doi_cell(pooled_ref, percent=67.0, top=120):
    """Depth cell of the depth of investigation of the pooled reference chains: opacity = 1 - (credible range normalised
    to [0, 1]) (Histogram.opacity, Histogram.py:330-354), searched from the bottom up (compute_doi :514-517)."""
    lo, hi = percentile_bins(pooled_ref, 0.05), percentile_bins(pooled_ref, 0.95)
    rng = (hi - lo).astype(np.float64)
    t = (rng - rng.min()) / max(rng.max() - rng.min(), 1e-300)
    op = 1.0 - t
    j = op.size - 1
    while op[j] < 0.01 * percent and j >= 1:
        j -= 1
    return int(min(max(j, 20), top))


def chain_stats(r, n_burn_min=5000):
    """Per-chain statistics from a dict with hitmap / edges_hist / ncells_hist / misfit_trace and either the
    reference fixture's fields (iterations, accept_trace, burned_in) or `scalars`-derived ones (iterations, n_accept)."""
    it = int(r["iterations"])
    nc = np.asarray(r["ncells_hist"], dtype=np.float64)
    mt = np.asarray(r["misfit_trace"], dtype=np.float64)[:it]
    acc = float(r["acceptance"]) if "acceptance" in r else float(np.asarray(r["accept_trace"])[:it + 1].mean())
    return dict(acceptance=acc, kbar=float((nc * np.arange(nc.size)).sum() / max(nc.sum(), 1.0)),
                misfit_median=float(np.median(mt[min(n_burn_min, it - 1):])), misfit_min=float(mt.min()),
                burned_in=bool(r["burned_in"]))


def compare(refs, runs, top=120):
    """refs: live-reference chains; runs: chains under test (same keys).  Returns the metrics; `check` asserts."""
    pooled_ref = sum(np.asarray(r["hitmap"], dtype=np.int64) for r in refs)
    pooled = sum(np.asarray(r["hitmap"], dtype=np.int64) for r in runs)
    nz = doi_cell(pooled_ref, top=top)
    out = {"cells": nz}
    for p in (0.05, 0.5, 0.95):
        rp = np.array([percentile_bins(np.asarray(r["hitmap"])[:, :nz], p) for r in refs])
        op = percentile_bins(pooled[:, :nz], p)
        out["inside_p%g" % (100 * p)] = float(((op >= rp.min(axis=0) - 2) & (op <= rp.max(axis=0) + 2)).mean())
        out["pooled_rms_p%g" % (100 * p)] = float(np.sqrt(np.mean((op - percentile_bins(pooled_ref[:, :nz], p)) ** 2.0)))
    re = sum(np.asarray(r["edges_hist"], dtype=np.int64) for r in refs)
    oe = sum(np.asarray(r["edges_hist"], dtype=np.int64) for r in runs)
    out["interface_peak"] = int(oe.argmax())
    out["interface_ref_top3"] = [int(i) for i in np.argsort(re)[-3:]]
    peaks = np.r_[np.argsort(re)[-3:], [int(np.asarray(r["edges_hist"]).argmax()) for r in refs]]
    out["interface_ref_chain_peaks"] = sorted(set(int(i) for i in peaks[3:]))
    out["interface_offset"] = int(np.min(np.abs(peaks - oe.argmax())))
    rs, os_ = [chain_stats(r) for r in refs], [chain_stats(r) for r in runs]
    for k in ("acceptance", "kbar"):
        a = np.array([s[k] for s in rs])
        out[k + "_ref"], out[k + "_ref_se"] = float(a.mean()), float(a.std(ddof=1) / np.sqrt(a.size))
        out[k] = float(np.mean([s[k] for s in os_]))
    rm = np.array([s["misfit_median"] for s in rs])
    out["misfit_median"] = float(np.median([s["misfit_median"] for s in os_]))
    out["misfit_ref_range"] = [float(rm.min()), float(rm.max())]
    out["misfit_min"] = float(np.min([s["misfit_min"] for s in os_]))
    out["misfit_min_ref"] = float(np.min([s["misfit_min"] for s in rs]))
    out["burned_in"] = float(np.mean([s["burned_in"] for s in os_]))
    out["burned_in_ref"] = float(np.mean([s["burned_in"] for s in rs]))
    out["n_ref"], out["n"] = len(rs), len(os_)
    return out


def check(m, inside=(0.85, 0.9, 0.85), peak_cells=2, layers=0.5, acceptance=0.05):
    """The tolerances of SURVEY.md 8(d) on the metrics of `compare`."""
    for p, frac in zip((5, 50, 95), inside):
        assert m["inside_p%g" % p] >= frac, ("p%g profile" % p, m)
    assert m["interface_offset"] <= peak_cells, ("interface peak", m)
    assert abs(m["kbar"] - m["kbar_ref"]) <= max(layers, 2.0 * m["kbar_ref_se"]), ("mean layer count", m)
    assert abs(m["acceptance"] - m["acceptance_ref"]) <= max(acceptance, 2.0 * m["acceptance_ref_se"]), ("acceptance", m)
    lo, hi = m["misfit_ref_range"]
    assert lo / 1.5 <= m["misfit_median"] <= hi * 1.5, ("misfit after burn-in", m)
    assert m["misfit_min"] <= 1.25 * m["misfit_min_ref"], ("lowest misfit", m)
    se = np.sqrt(max(m["burned_in_ref"] * (1.0 - m["burned_in_ref"]), 0.25 / m["n_ref"]) / m["n_ref"])
    assert abs(m["burned_in"] - m["burned_in_ref"]) <= max(0.25, 2.0 * se), ("burned-in fraction", m)
