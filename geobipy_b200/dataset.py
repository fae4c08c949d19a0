"""Survey-level input and output either side of the hot path.

Input  : `FdemData.read_csv` reads a reference-format frequency-domain CSV (same header conventions as
         geobipy/src/classes/data/dataset/FdemData.py:520-683, Data.py:488-528, pointcloud/Point.py:336-383)
         into the batched device layout the sampler takes.
Driver : `Inference3D` replaces the per-record loop of geobipy/src/inversion/Inference3D.py:458-492 - one
         call inverts every sounding of the survey concurrently (optionally sharded over the GPUs of the
         node, parallel.run_sharded) - and collects, per flight line, the arrays Inference1D.writeHdf
         (:1050-1090) / Inference2D.createHdf (:2001-2016) store.
Output : one `<line>.npz` per flight line with those arrays (h5py is not available in this image, so the
         HDF5 container itself is not written; the array names follow the reference's dataset names), plus
         the posterior summaries the reference derives afterwards: mean / median / 5-95 % credible interval
         of conductivity with depth (Mesh._mean, Mesh._percentile, classes/mesh/Mesh.py:80, :173-217) and the
         interface probability (Inference2D.interface_probability, Inference2D.py:959-961).
"""
import csv
import os

import numpy as np

from . import _lib, api, ops

__all__ = ["FdemData", "Inference3D", "summarise", "opacity_and_doi"]

_LINE = ("line", "linenumber", "line_number")
_FID = ("fid", "fiducial", "id")
_X = ("e", "x", "easting")
_Y = ("n", "y", "northing")
_Z = ("alt", "altitude", "laser", "bheight", "height")
_ELEV = ("dtm", "dem_elev", "dem_np", "topo", "elev", "elevation")


def _csv_channels(header):
    """Column roles from the header, as the reference assigns them."""
    roles = dict(line=None, fid=None, x=None, y=None, z=None, elev=None, inphase=[], quad=[], inerr=[], quaderr=[])
    for ch in header:
        c = ch.strip().lower()
        if c in _LINE:
            roles["line"] = ch
        elif c in _FID:
            roles["fid"] = ch
        elif c in _X:
            roles["x"] = ch
        elif c in _Y:
            roles["y"] = ch
        elif c in _Z:
            roles["z"] = ch
        elif c in _ELEV:
            roles["elev"] = ch
        elif any(lbl in c for lbl in ("cpi", "i_", "in_phase")):
            (roles["inerr"] if "err" in c else roles["inphase"]).append(ch)
        elif any(lbl in c for lbl in ("cpq", "q_", "quad")):
            (roles["quaderr"] if "err" in c else roles["quad"]).append(ch)
    assert roles["line"] is not None and roles["fid"] is not None, Exception("File must contain columns for line and fiducial.")
    assert None not in (roles["x"], roles["y"], roles["z"]), Exception(
        "File must contain columns for easting, northing, height. May also have an elevation column")
    return roles


class FdemData:
    """Whole-survey frequency-domain table (classes/data/dataset/FdemData.py, reader part)."""

    def __init__(self, system=None):
        if isinstance(system, str):
            system = api.FdemSystem.read(system)
        self.system = system
        self.lineNumber = self.fiducial = self.x = self.y = self.z = self.elevation = None
        self.data = self._std = None
        self.relative_error, self.additive_error = 0.01, 0.0     # Data.__init__ :94-95

    @property
    def std(self):
        """sqrt((relative_error data)^2 + additive_error^2), recomputed on every access with the data set's relative error
        (1 % unless set) and additive error (0) - the reference's getter (Data.std, classes/data/dataset/Data.py:376-384),
        which also overrides standard deviations read from `*err*` columns; those stay available as `std_from_file`."""
        if self.data is None:
            return None
        return np.sqrt((self.relative_error * self.data) ** 2 + self.additive_error ** 2)

    @std.setter
    def std(self, values):
        self._std = values

    @property
    def std_from_file(self):
        return self._std

    @property
    def nPoints(self):
        return 0 if self.data is None else self.data.shape[0]

    @property
    def nChannels(self):
        return 2 * self.system.nFrequencies

    @classmethod
    def read_csv(cls, dataFilename, system):
        """Read a data file whose header names Line, Fiducial, Easting, Northing, Height[, Elevation] and one
        in-phase + one quadrature column per frequency (any order; `*err*` columns are uncertainties)."""
        self = cls(system=system)
        # numbers are parsed as the reference parses them (pandas.read_csv with its default float parser, comma or
        # whitespace separated: FdemData.read_csv :555-559), so that a file gives bit-identical arrays in both
        import pandas as pd
        with open(dataFilename) as f:
            sample = f.readline()
        if "," in sample:
            df = pd.read_csv(dataFilename, index_col=False, skipinitialspace=True)
        else:
            df = pd.read_csv(dataFilename, index_col=False, sep=r"\s+", skipinitialspace=True)
        df = df.replace("NaN", np.nan)
        header = [str(h).strip() for h in df.columns]
        df.columns = header
        roles = _csv_channels(header)

        def num(name):
            return df[name].to_numpy(dtype=np.float64)
        body = df
        self.lineNumber, self.fiducial = num(roles["line"]), num(roles["fid"])
        self.x, self.y, self.z = num(roles["x"]), num(roles["y"]), num(roles["z"])
        self.elevation = num(roles["elev"]) if roles["elev"] else np.zeros(len(body))
        chans = roles["inphase"] + roles["quad"]
        assert len(chans) == self.nChannels, ValueError(
            "file has %d data channels, the system needs %d" % (len(chans), self.nChannels))
        self.data = np.stack([num(c) for c in chans], axis=1)
        errs = roles["inerr"] + roles["quaderr"]
        self.std = np.stack([num(c) for c in errs], axis=1) if errs else 0.1 * self.data
        return self

    def datapoint(self, i):
        return api.FdemDataPoint(x=self.x[i], y=self.y[i], z=self.z[i], elevation=self.elevation[i], data=self.data[i],
                                 std=self.std[i], system=self.system, lineNumber=self.lineNumber[i], fiducial=self.fiducial[i])

    def line(self, line_number):
        return np.flatnonzero(self.lineNumber == line_number)

    def subset(self, idx):
        """The soundings `idx` as a data set of their own (Data.__getitem__ of the reference)."""
        out = FdemData(self.system)
        for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data"):
            setattr(out, k, np.asarray(getattr(self, k))[idx])
        out.std = None if self._std is None else np.asarray(self._std)[idx]
        out.relative_error, out.additive_error = self.relative_error, self.additive_error
        return out

    @property
    def lines(self):
        return np.unique(self.lineNumber)


def _percentile_bins(hitmap, percent):
    """First bin whose cumulative count reaches percent of the column total (Mesh._percentile)."""
    cs = np.cumsum(hitmap, axis=-2, dtype=np.float64)
    tot = cs[..., -1:, :]
    frac = np.divide(cs, tot, out=np.zeros_like(cs), where=tot > 0.0)   # the reference's rule, ties included (Mesh.py:196-208)
    return np.minimum((frac < percent * 0.01).sum(axis=-2), hitmap.shape[-2] - 1)


def summarise(result, opt, line_id=None, doi_percent=67.0, credible_percent=90.0):
    """Host (numpy) statement of `summarise_device` - the checker of the device kernels in the tests, and what turns
    host-resident result arrays into summaries: conductivity mean / p5 / p50 / p95 / mode per depth cell [B, n_depth],
    credible range (decades), opacity and depth of investigation per flight line, interface probability."""
    hm = np.asarray(result["hitmap"])
    hs = np.asarray(result["scalars"])[:, _lib.S_HALFSPACE]
    s = np.log(1.0 + opt.factor) * opt.sigma_bins_nstd
    edges = np.linspace(-s, s, opt.n_sigma_bins + 1)[None, :] + np.log(hs)[:, None]   # ln sigma bin edges
    centres = 0.5 * (edges[:, 1:] + edges[:, :-1])
    h = hm.astype(np.float64)
    tot = np.maximum(h.sum(axis=1), 1.0)
    out = {"mean": np.exp((h * centres[:, :, None]).sum(axis=1) / tot)}
    for p in (5.0, 50.0, 95.0):
        idx = _percentile_bins(hm, p)
        out["p%g" % p] = np.exp(np.take_along_axis(centres, idx, axis=1))
    out["mode"] = np.exp(np.take_along_axis(centres, np.argmax(hm, axis=1), axis=1))
    hp = 0.5 * min(credible_percent, 100.0 - credible_percent)
    dxl = (edges[:, 1] - edges[:, 0])[:, None]
    out["credible_range"] = (_percentile_bins(hm, 100.0 - hp) - _percentile_bins(hm, hp)) * dxl / np.log(10.0)
    depth_edges = np.arange(0.0, 1.1 * opt.max_edge, 0.5 * opt.min_width)
    out["opacity"], out["doi"] = np.zeros_like(out["credible_range"]), np.zeros(hm.shape[0])
    lines = np.zeros(hm.shape[0]) if line_id is None else np.asarray(line_id)
    for ln in np.unique(lines):
        m = lines == ln
        out["opacity"][m], out["doi"][m] = opacity_and_doi_from_range(out["credible_range"][m], depth_edges, doi_percent)
    eh = np.asarray(result["edges_hist"]).astype(np.float64)
    out["interface_probability"] = eh / np.maximum(eh.sum(axis=1, keepdims=True), 1.0)
    out["depth_edges"] = np.arange(0.0, 1.1 * opt.max_edge, 0.5 * opt.min_width)
    if "height_hist" in result:
        # sampled sensor height (solve_z).  The histogram bins z - z_ref, z_ref = the centre of the height prior: the
        # input height, re-centred on the sampled height by every reset() (scalar slot S_HEIGHT_REF).  height_mean [B]
        # is the posterior mean height itself; Inference3D.infer adds height_change_mean = height_mean - input height.
        hh = np.asarray(result["height_hist"]).astype(np.float64)
        c = -opt.max_height_change + (np.arange(hh.shape[1]) + 0.5) * (2.0 * opt.max_height_change / hh.shape[1])
        out["height_mean"] = (np.asarray(result["scalars"])[:, _lib.S_HEIGHT_REF]
                              + (hh * c[None, :]).sum(axis=1) / np.maximum(hh.sum(axis=1), 1.0))
    return out


def summarise_device(result, opt, line_id=None, doi_percent=67.0, credible_percent=90.0):
    """The posterior summaries of `summarise` plus mode, credible range, opacity and depth of investigation, computed ON
    THE DEVICE from the torch CUDA result arrays of `ops.rjmcmc_run` (hand-written kernels gbp_summarise_posterior /
    gbp_opacity_doi).  line_id [B] (integers): the flight line of every sounding - opacity is the
    credible range normalised over the line (Inference2D.compute_opacity, Inference2D.py:1011-1023), the depth of
    investigation its compute_doi (:493-532); None = one line.  Returns a dict of torch tensors (conductivities in S/m)."""
    import torch
    hm = result["hitmap"]
    dev = hm.device
    B, ns, nd = hm.shape
    s = np.log(1.0 + opt.factor) * opt.sigma_bins_nstd
    dx = 2.0 * s / opt.n_sigma_bins
    lo = torch.log(result["scalars"][:, _lib.S_HALFSPACE].to(torch.float64)) - s
    r = ops.summarise_posterior(hm.contiguous(), lo, dx, percentiles=(5.0, 50.0, 95.0), credible_percent=credible_percent)
    if line_id is None:
        grp, ng = torch.zeros(B, dtype=torch.int32, device=dev), 1
    else:
        u, inv = np.unique(np.asarray(line_id), return_inverse=True)
        grp, ng = torch.as_tensor(inv.astype(np.int32), device=dev), int(u.size)
    opacity, doi_cell, _ = ops.opacity_doi(r["range_bins"], group=grp, n_groups=ng, doi_percent=doi_percent)
    depth_edges = np.arange(0.0, 1.1 * opt.max_edge, 0.5 * opt.min_width)
    centres = torch.as_tensor(0.5 * (depth_edges[1:] + depth_edges[:-1]), device=dev)
    eh = result["edges_hist"].to(torch.float64)
    out = {"mean": torch.exp(r["mean"]), "p5": torch.exp(r["pct"][0]), "p50": torch.exp(r["pct"][1]), "p95": torch.exp(r["pct"][2]),
           "mode": torch.exp(r["mode"]), "credible_range": r["credible_range"], "opacity": opacity,
           "doi": centres[doi_cell.long()], "interface_probability": eh / eh.sum(dim=1, keepdim=True).clamp_min(1.0),
           "depth_edges": torch.as_tensor(depth_edges, device=dev)}
    if "height_hist" in result:
        hh = result["height_hist"].to(torch.float64)
        c = torch.as_tensor(-opt.max_height_change + (np.arange(hh.shape[1]) + 0.5) * (2.0 * opt.max_height_change / hh.shape[1]), device=dev)
        out["height_mean"] = result["scalars"][:, _lib.S_HEIGHT_REF] + (hh * c[None, :]).sum(dim=1) / hh.sum(dim=1).clamp_min(1.0)
    return out


def opacity_and_doi(p_low, p_high, depth_edges, doi_percent=67.0):
    """Opacity and depth of investigation of one flight line from its credible intervals (restated from
    Inference2D.compute_opacity inversion/Inference2D.py:1011-1023 = Histogram.opacity / transparency
    classes/statistics/Histogram.py:330-354, 509-541 over Mesh._credible_range classes/mesh/Mesh.py:58-78, and
    Inference2D.compute_doi :493-532; no HDF5 reader in this image, so not run against the reference's own).

    p_low / p_high [n_soundings, n_depth]: the 5 % / 95 % conductivities (compute_opacity's percent = 90).  The credible
    range is their ratio in decades (the hitmap's value axis has log = 10, Model.py:677-679); transparency is that
    range normalised by its minimum and maximum over the WHOLE line, NaN -> 1; opacity = 1 - transparency.  The depth
    of investigation is the centre of the deepest cell whose opacity reaches doi_percent / 100, searched from the
    bottom up and never above the second cell from the top (the loop of compute_doi stops at j = 1... 0).
    Returns (opacity [n, n_depth], doi [n])."""
    lo, hi = np.asarray(p_low, dtype=np.float64), np.asarray(p_high, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rng = np.abs(np.log10(hi) - np.log10(lo))
    return opacity_and_doi_from_range(rng, depth_edges, doi_percent)


def opacity_and_doi_from_range(credible_range, depth_edges, doi_percent=67.0):
    """The same from the credible range itself [n_soundings, n_depth] (Mesh._credible_range of the counts)."""
    rng = np.asarray(credible_range, dtype=np.float64)
    mn, mx = np.nanmin(rng), np.nanmax(rng)
    t = (rng - mn) / (mx - mn) if (mx - mn) > 0.0 else rng - mn
    t = np.where(np.isnan(t), 1.0, t)
    opacity = 1.0 - t
    e = np.asarray(depth_edges, dtype=np.float64)
    centres = 0.5 * (e[1:] + e[:-1])
    p = 0.01 * doi_percent
    n, nz = opacity.shape
    doi = np.empty(n)
    for i in range(n):
        j = nz - 1
        while opacity[i, j] < p and j >= 1:   # compute_doi :514-517
            j -= 1
        doi[i] = centres[j]
    return opacity, doi


class Inference3D:
    """Survey driver: `Inference3D(data).infer(**options)` (inversion/Inference3D.py:451-492, :503-635)."""

    def __init__(self, data, prng=None, seed=None, world=None):
        from .tdem import TdemData
        assert isinstance(data, (FdemData, TdemData)), TypeError("data must be a FdemData or a TdemData")
        self.data = data
        if seed is None:
            seed = int(prng.integers(0, 2 ** 63 - 1)) if prng is not None else 0
        self.seed = int(seed) & (2 ** 64 - 1)
        self.results = None
        self.options = None

    def infer(self, index=None, fiducial=None, line_number=None, precision=_lib.PRECISION_F32, device=0,
              max_iterations=0, sharded=False, **options):
        """Invert every sounding (or the one selected by index / fiducial+line_number, as the reference's
        `infer`).  Sounding i always uses random stream (seed, i)."""
        d = self.data
        if index is None and fiducial is not None:
            index = int(np.flatnonzero((d.fiducial == fiducial) & ((d.lineNumber == line_number) if line_number is not None else True))[0])
        sel = np.arange(d.nPoints) if index is None else np.atleast_1d(index)
        from .tdem import TempestData
        if isinstance(d, TempestData):
            # tempest_options: initial_additive_error is the additive level of every CHANNEL; what is sampled is one
            # multiplier per component, starting at 1 (Tempest_datapoint.set_priors / set_proposals :478-510)
            level = np.asarray(options["initial_additive_error"], dtype=np.float64).reshape(-1)
            assert level.size == d.nChannels, ValueError("initial_additive_error must hold the additive level of every channel ({})".format(d.nChannels))
            nc = d.n_components
            options = dict(options, initial_additive_error=[1.0] * nc if nc > 1 else 1.0)
            d.additive_error = np.tile(level, (d.nPoints, 1))
            sysc = d.survey_struct(additive_level=level)
        else:
            sysc = d.c_struct if hasattr(d, "c_struct") else d.system.c_struct   # time-domain surveys: systems + tx-rx offset
        opt = ops.options_from_reference(**options)
        self.options = opt
        self._struct = sysc
        import torch
        dev = torch.device("cuda", device)

        def up(a):
            return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        lines = np.asarray(d.lineNumber)[sel]
        if sharded:
            from . import parallel
            res = parallel.run_sharded(sysc, opt, d.data[sel], d.z[sel], seed=self.seed, precision=precision,
                                       outputs=ops.DEFAULT_OUTPUTS, max_iterations=max_iterations)
            if res is None:
                return None
        elif sel.size == d.nPoints:
            res = ops.rjmcmc_run(sysc, opt, up(d.data), up(d.z), seed=self.seed, first_index=0, max_iterations=max_iterations,
                                 precision=precision)
        else:
            parts = [ops.rjmcmc_run(sysc, opt, up(d.data[i:i + 1]), up(d.z[i:i + 1]), seed=self.seed, first_index=int(i),
                                    max_iterations=max_iterations, precision=precision) for i in sel]
            res = {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0]}
        # posterior summaries on the device, then everything comes to the host once
        summ = summarise_device(res, opt, line_id=lines)
        res = {k: v.cpu().numpy() for k, v in res.items()}
        res["index"] = sel
        res.update({"summary_" + k: v.cpu().numpy() for k, v in summ.items()})
        if "summary_height_mean" in res:
            res["summary_height_change_mean"] = res["summary_height_mean"] - np.asarray(d.z, dtype=np.float64)[sel]
        self.results = res
        return res

    def save(self, directory, format="h5", precision=_lib.PRECISION_F64, device=0):
        """One result file per flight line, `<line>.h5` in the reference's HDF5 layout (what `Inference3D.create_hdf5`
        + every sounding's `Inference1D.writeHdf` leave behind, inversion/Inference3D.py:276-340; `geobipy_b200.hdf`): the
        file `Inference2D` / `Inference3D` of the reference open for their sections and maps.  The predicted data of every
        sounding's best model come from ONE call of the forward operator.  `format="npz"` keeps round 1's flat archive
        (plus the per-line summaries the device computed)."""
        assert self.results is not None, "run infer() first"
        os.makedirs(directory, exist_ok=True)
        r, d = self.results, self.data
        files = []
        if format in ("h5", "hdf5"):
            from . import hdf
            from .tdem import TdemData
            opt = self.options
            sysc = self._struct
            idx_all = r["index"]
            k = r["scalars"][:, _lib.S_BEST_K].astype(np.int32)
            sig = np.where(np.isnan(r["best_sigma"]), 1.0, r["best_sigma"])
            thk = np.diff(np.where(np.isfinite(r["best_edges"]), r["best_edges"], 0.0), axis=1)
            thk[thk <= 0.0] = 1.0   # the half-space and the padding: ignored by the operator
            alt = r["scalars"][:, _lib.S_BEST_HEIGHT] if opt.solve_height else np.asarray(d.z, dtype=np.float64)[idx_all]
            predicted = ops.forward(sysc, k, sig, thk, alt, precision=precision, device=device)
            for ln in np.unique(np.asarray(d.lineNumber)[idx_all]):
                m = np.asarray(d.lineNumber)[idx_all] == ln
                sub = d.subset(idx_all[m])
                res = {key: v[m] for key, v in r.items() if isinstance(v, np.ndarray) and v.shape[:1] == m.shape and not key.startswith("summary_")}
                files.append(hdf.save_line(os.path.join(directory, "%g.h5" % ln), res, opt, sub, predicted[m]))
            return files
        assert format == "npz", ValueError("format must be 'h5' or 'npz'")
        for ln in np.unique(d.lineNumber[r["index"]]):
            m = d.lineNumber[r["index"]] == ln
            idx = r["index"][m]
            path = os.path.join(directory, "%s.npz" % (("%g" % ln)))
            line = {}
            if "summary_opacity" in r:   # per-line products of Inference2D (opacity, doi)
                line["opacity"], line["doi"] = r["summary_opacity"][m], r["summary_doi"][m]
            np.savez_compressed(
                path, line_number=ln, fiducial=d.fiducial[idx], x=d.x[idx], y=d.y[idx], z=d.z[idx],
                elevation=d.elevation[idx], data=d.data[idx], **line,
                **{k: (v[m] if isinstance(v, np.ndarray) and v.shape[:1] == m.shape else v) for k, v in r.items()})
            files.append(path)
        return files
