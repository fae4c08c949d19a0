"""Result files in the reference's HDF5 layout, written for a whole flight line at once.

The reference creates one `<line>.h5` per flight line (`Inference3D._create_HDF5_dataset` inversion/Inference3D.py:312-340
-> `Inference2D.createHdf` inversion/Inference2D.py:2001-2016 -> `Inference1D.createHdf` inversion/Inference1D.py:1002-1048)
and every sounding then writes its row (`Inference1D.writeHdf` :1050-1090: the current datapoint / model with their
posteriors, then the best datapoint / model over the same datasets).  Each class of the reference contributes a nested
group (`DataArray.createHdf` classes/core/DataArray.py:1011-1097, `StatArray.createHdf` classes/statistics/StatArray.py:738,
`RectilinearMesh1D._create_hdf_2d` classes/mesh/RectilinearMesh1D.py:1657-1684, `Model.createHdf` classes/model/Model.py:853,
`DataPoint.createHdf` classes/data/datapoint/DataPoint.py:746-773, `FdemDataPoint.createHdf` FdemDataPoint.py:282-294,
`TdemDataPoint.createHdf` TdemDataPoint.py:603-626, `Point.createHdf` classes/pointcloud/Point.py:1403-1427).

Here the tree of a line is built from the arrays the sampler returns for ALL its soundings (`ops.rjmcmc_run`), one
vectorised write per dataset instead of one h5py call per sounding and dataset.  Group / dataset names, shapes, dtypes,
fill values and attributes are those of the reference: `tests/golden/hdf_layout_*.npz` holds the trees the reference's own
createHdf / writeHdf code produced (through `h5lite` standing in for h5py) and `tests/test_hdf.py` compares entry by entry.

`parent` is any h5py-like group: a real `h5py.File` where h5py is installed, else `geobipy_b200.h5lite.File` (this image),
which serialises to a real HDF5 file on close.  `h5()` returns whichever module is available.
"""
import numpy as np

from . import _lib, ops

__all__ = ["h5", "create_line", "write_line", "save_line", "read_line"]

_SIEMENS = "$\\frac{S}{m}$"
_VM2 = "$\\frac{V}{m^{2}}$"
_DEG = "$^{o}$"


def h5():
    """The HDF5 module to write with: h5py when installed, else the bundled h5lite."""
    try:
        import h5py
        return h5py
    except ImportError:
        from . import h5lite
        return h5lite


# ------------------------------------------------------------------------------------------ the reference's building blocks
def _group(parent, name, repr_=None, label=None, units=None):
    g = parent.create_group(name)
    if repr_ is not None:
        g.attrs["repr"] = repr_
    if label is not None:
        g.attrs["name"] = label
    if units is not None:
        g.attrs["units"] = units
    return g


def _data_array(parent, name, shape, dtype=np.float64, fill=np.nan, label=None, units=None, repr_="DataArray", values=None):
    """DataArray.createHdf (+ writeHdf when `values` is given): group {repr, name, units} with one dataset 'data'."""
    g = _group(parent, name, repr_, label, units)
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    if values is not None:
        g.create_dataset("data", data=np.ascontiguousarray(values, dtype=dtype).reshape(shape))
    else:
        g.create_dataset("data", shape, dtype=dtype, fillvalue=fill)
    return g


def _x_axis(parent, n):
    """The axis a line adds: RectilinearMesh1D(centres = 0 .. n-1) (Inference1D.createHdf turns the fiducials into their
    count, :1013-1015; RectilinearMesh1D._create_hdf_2d :1662-1667)."""
    g = _group(parent, "x", "RectilinearMesh1D")
    _data_array(g, "edges", n + 1, label="", units="", values=np.arange(n + 1, dtype=np.float64) - 0.5)
    g.create_dataset("dimension", data=np.asarray([0], dtype=np.int32))
    return g


def _histogram(parent, name, n, edges, label, units, log=None, relative=False, y_dimension=1):
    """Histogram over a 1-D mesh, one row per sounding: Histogram(mesh = RectilinearMesh2D(x = soundings, y = bins))."""
    h = _group(parent, name, "Histogram")
    m = _group(h, "mesh", "RectilinearMesh2D")
    _x_axis(m, n)
    y = _group(m, "y", "RectilinearMesh1D")
    if log is not None:
        y.create_dataset("log", data=np.int64(log))
    if relative:
        _data_array(y, "relative_to", n, label=label, units=units)
    _data_array(y, "edges", edges.size, label=label, units=units, values=edges)
    y.create_dataset("dimension", data=np.asarray([y_dimension], dtype=np.int32))
    _data_array(h, "values", (n, edges.size - 1), dtype=np.int32, fill=0, label="Frequency")
    return h


def _loop(parent, name, n, repr_="CircularLoop", values=None):
    """CircularLoop(s).createHdf (classes/system/CircularLoop.py:95-104, EmLoop.py:418-431)."""
    g = _group(parent, name, repr_)
    spec = (("x", "Easting", "m"), ("y", "Northing", "m"), ("z", "Height", "m"), ("elevation", "Elevation", "m"),
            ("pitch", "Pitch", _DEG), ("roll", "Roll", _DEG), ("yaw", "Yaw", _DEG), ("moment", "Moment", ""),
            ("orientation", "Orientation", ""), ("radius", "Radius", "m"))
    for key, label, units in spec:
        dt = np.int32 if key == "orientation" else np.float64
        v = None if values is None else values[key]
        _data_array(g, key, n, dtype=dt, fill=(0 if dt is np.int32 else np.nan), label=label, units=units, values=v)
    return g


_ORI = {"x": 0, "y": 1, "z": 2}


def _grids(opt):
    """Bin edges of the posteriors as the reference stores them (relative, log10 where its meshes are logarithmic)."""
    g = ops.posterior_grids(opt, 1.0)
    ln10 = np.log(10.0)
    out = {"sigma_rel": np.log(g["sigma_edges"]) / ln10, "depth": np.asarray(g["depth_edges"], dtype=np.float64),
           "ncells": np.arange(opt.max_layers + 1.0) + 0.5}
    out["ncells"] = np.r_[-0.5, out["ncells"]]
    for key in ("rel_edges", "add_edges"):
        e = np.atleast_2d(np.asarray(g[key], dtype=np.float64))
        # relative_to = 0.5 (max - min) of the bins (DataPoint.set_relative_error_posterior :673), a log10 mesh
        out[key] = [(np.log10(b) - np.log10(0.5 * (b.max() - b.min())), np.log10(0.5 * (b.max() - b.min()))) for b in e]
        # the same bins on a mesh without log (Tempest_datapoint.set_additive_error_posterior :535-547)
        out[key + "_linear"] = [(b - 0.5 * (b.max() - b.min()), 0.5 * (b.max() - b.min())) for b in e]
    return out


# ------------------------------------------------------------------------------------------ create
def create_line(parent, n, opt, data, n_markov_chains=None, update_plot_every=None, interactive_plot=False,
                reciprocate_parameter=True):
    """`Inference2D.createHdf(parent, inference1d)` for a line of `n` soundings: every group, dataset, attribute and fill
    value of the reference's file.  `data` is the line's FdemData / TdemData (systems, channel count)."""
    from .tdem import TdemData, TempestData
    tdem = isinstance(data, TdemData)
    tempest = isinstance(data, TempestData)   # errors per component, additive levels per channel, data in fT
    nsys = 2 if opt.n_systems > 1 else 1      # entries of the error arrays (systems; components of a Tempest datapoint)
    G = _grids(opt)
    C = int(data.nChannels)
    d_units = "fT" if tempest else (_VM2 if tdem else "ppm")
    tz = tdem and opt.solve_height            # solve_transmitter_z: the transmitter loop's z carries the posterior

    # ---- data (DataPoint.createHdf + Point.createHdf)
    d = _group(parent, "data", "TempestData" if tempest else ("TdemData" if tdem else "FdemData"))
    for key, label in (("x", "Easting"), ("y", "Northing"), ("z", "Height"), ("elevation", "Elevation")):
        if key == "z" and opt.solve_height and not tdem:    # a sampled height is a StatArray with its posterior (Point.py:1013-1020)
            z = _data_array(d, "z", n, label=label, units="m", repr_="StatArray")
            z.create_dataset("n_posteriors", data=np.int64(1))
            dz = opt.max_height_change
            _histogram(z, "posterior", n, np.linspace(-dz, dz, opt.n_err_bins + 1), "Height", "m", relative=True)
        else:
            _data_array(d, key, n, label=label, units="m")
    _data_array(d, "fiducial", n, label="fiducial")
    _data_array(d, "line_number", n, label="Line number")
    _data_array(d, "data", (n, C), label="Data" if tempest else ("Secondary field" if tdem else "Frequency domain data"), units=d_units)
    _data_array(d, "std", (n, C), label="Standard deviation", units=d_units)
    _data_array(d, "predicted_data", (n, C), label="Predicted Data" if (tempest or not tdem) else "Predicted secondary field",
                units=d_units)
    errs = [("relative_error", "$\\epsilon_{Relative}$" if tempest else "$\\epsilon_{Relative}x10^{2}$", "%", "rel_edges", 10)]
    if tempest:
        # Tempest_datapoint.createHdf :566-575: the additive level of every channel is a plain array, the sampled
        # multiplier (one per component) carries the posteriors - log-spaced bins kept on a LINEAR mesh (:535-547)
        _data_array(d, "additive_error", (n, C), label="$\\epsilon_{additive}$", units=d_units)
        errs.append(("additive_error_multiplier", "Multiplier", None, "add_edges_linear", None))
    else:
        errs.append(("additive_error", "$\\epsilon_{Additive}$", d_units, "add_edges", 10))
    for key, label, units, gkey, log in errs:
        e = _data_array(d, key, (n, nsys) if nsys > 1 else n, label=label, units=units, repr_="StatArray")
        e.create_dataset("n_posteriors", data=np.int64(nsys))
        for i in range(nsys):
            edges, _ = G[gkey][i]
            _histogram(e, "posterior%d" % i if nsys > 1 else "posterior", n, edges, label, units, log=log, relative=True)
    if tdem:
        d.create_dataset("nSystems", data=np.int64(len(data.system)))
        for i, s in enumerate(data.system):
            g = _group(d, "System%d" % i, "TdemSystem")
            g.attrs["data"] = list(s.stm_lines)          # the .stm file, line by line (TdemSystem_GAAEM.toHdf)
        d.create_dataset("components", data=np.asarray([_ORI[c] for c in data.system[0].components], dtype=np.int32))
        lp = _group(d, "loop_pair", "Loop_pair")
        for key, label in (("x", "Easting"), ("y", "Northing"), ("z", "Height"), ("elevation", "Elevation")):
            _data_array(lp, key, n, label=label, units="m")
        t = _loop(lp, "transmitter", n, "CircularLoops")
        if tz:   # a sampled transmitter height: StatArray with its posterior (EmLoop.createHdf -> Point.createHdf :1403-1427)
            z = t["z"]
            z.attrs["repr"] = "StatArray"
            z.create_dataset("n_posteriors", data=np.int64(1))
            dz = opt.max_height_change
            _histogram(z, "posterior", n, np.linspace(-dz, dz, opt.n_err_bins + 1), "Height", "m", relative=True)
        _loop(lp, "receiver", n, "CircularLoops")
        pshape = (n, data.system[0].n_components) if tempest else n
        _data_array(d, "primary_field", pshape, label="Primary field", units=d_units)
        _data_array(d, "secondary_field", (n, C), label="Secondary field", units=d_units)
        _data_array(d, "predicted_primary_field", pshape, label="Predicted primary field", units=d_units)
        _data_array(d, "predicted_secondary_field", (n, C), label="Predicted secondary field", units=d_units)
    else:
        s = data.system
        g = _group(d, "sys", "FdemSystem")
        F = s.nFrequencies
        _data_array(g, "freq", F, label="Frequencies", units="Hz", values=s.frequencies)
        for name, lp in (("T", s.transmitter), ("R", s.receiver)):
            z0 = np.zeros(F)
            _loop(g, name, F, values=dict(x=lp.x, y=lp.y, z=lp.z, elevation=z0, pitch=z0, roll=z0, yaw=z0, moment=lp.moment,
                                          orientation=[_ORI[o] for o in lp.orientation], radius=z0))

    # ---- scalars of the inversion (Inference1D.createHdf :1020-1040)
    N = int(opt.n_markov_chains if n_markov_chains is None else n_markov_chains)
    parent.create_dataset("update_plot_every", data=np.int32(opt.update_plot_every if update_plot_every is None else update_plot_every))
    parent.create_dataset("interactive_plot", data=np.bool_(interactive_plot))
    parent.create_dataset("reciprocate_parameter", data=np.bool_(reciprocate_parameter))
    parent.create_dataset("n_markov_chains", data=np.int64(N))
    parent.create_dataset("nsystems", data=np.int64(len(data.system) if tdem else 1))
    for key in ("iteration", "burned_in_iteration", "best_iteration"):
        parent.create_dataset(key, (n,), dtype=np.int64, fillvalue=0)
    parent.create_dataset("burned_in", (n,), dtype=np.bool_, fillvalue=0)
    for key in ("multiplier", "invtime", "savetime"):
        parent.create_dataset(key, (n,), dtype=np.float64, fillvalue=np.nan)
    _data_array(parent, "acceptance_rate", (n, 2 * N), dtype=np.uint8, fill=0, label="% Acceptance")
    _data_array(parent, "phids", (n, 2 * N), label="Data Misfit")
    _data_array(parent, "halfspace", n, label="halfspace", units=_SIEMENS)

    # ---- model (Model.createHdf of the model padded to max_cells, RectilinearMesh2D_stitched)
    ml, nz = opt.max_layers, G["depth"].size - 1
    m = _group(parent, "model", "Model")
    mesh = _group(m, "mesh", "RectilinearMesh2D_stitched")
    _x_axis(mesh, n)
    y = mesh.create_group("y")
    e = _data_array(y, "edges", (n, ml + 1), repr_="StatArray")
    e.create_dataset("n_posteriors", data=np.int64(1))
    _histogram(e, "posterior", n, G["depth"], "Depth", "m", y_dimension=0)
    _data_array(y, "relative_to", n)
    nc = _data_array(mesh, "nCells", n, dtype=np.int32, fill=0, label="Number of cells", repr_="StatArray")
    nc.create_dataset("n_posteriors", data=np.int64(1))
    _histogram(nc, "posterior", n, G["ncells"], "# of Layers", "", y_dimension=0)
    v = _data_array(m, "values", (n, ml), label="Conductivity", units=_SIEMENS, repr_="StatArray")
    v.create_dataset("n_posteriors", data=np.int64(1))
    h = _group(v, "posterior", "Histogram")
    hm = _group(h, "mesh", "RectilinearMesh3D")
    _x_axis(hm, n)
    hy = _group(hm, "y", "RectilinearMesh1D")
    hy.create_dataset("log", data=np.int64(10))
    _data_array(hy, "relative_to", n)
    _data_array(hy, "edges", G["sigma_rel"].size, label="Conductivity", units=_SIEMENS, values=G["sigma_rel"])
    hy.create_dataset("dimension", data=np.asarray([1], dtype=np.int32))
    hz = _group(hm, "z", "RectilinearMesh1D")
    _data_array(hz, "edges", nz + 1, label="Depth", units="m", values=G["depth"])
    hz.create_dataset("dimension", data=np.asarray([2], dtype=np.int32))
    _data_array(h, "values", (n, opt.n_sigma_bins, nz), dtype=np.int32, fill=0, label="Frequency")
    return parent


# ------------------------------------------------------------------------------------------ write
def _std(tdem, data, rel, add, off_times=None):
    """DataPoint.std :268-282 / TdemDataPoint.std :329-379 for every sounding of the block (rel, add: [n, systems])."""
    data = np.asarray(data, dtype=np.float64)
    if not tdem:
        return np.sqrt((rel[:, :1] * data) ** 2 + add[:, :1] ** 2)
    out = np.empty_like(data)
    o = 0
    for i, t in enumerate(off_times):
        c = slice(o, o + t.size)
        a = np.exp(np.log(add[:, i:i + 1]) - 0.5 * (np.log(t)[None, :] - np.log(1e-3)))
        out[:, c] = np.sqrt((rel[:, i:i + 1] * data[:, c]) ** 2 + a ** 2)
        o += t.size
    return out


def write_line(parent, res, opt, data, predicted_best, rows=None, multiplier=1.0):
    """`Inference1D.writeHdf` for every sounding of a block at once.

    res: the arrays `ops.rjmcmc_run` returned for the block (numpy, leading dimension m); data: the block's FdemData /
    TdemData / TempestData (m soundings); predicted_best [m, C]: forward response of each sounding's best model (the caller
    runs the forward operator once for the block; the secondary field - a Tempest line adds the primary field here); rows: where the block's soundings sit in the file (default 0 .. m-1 - the
    reference sorts a line by fiducial, Inference2D.createHdf :2011-2012)."""
    from .tdem import TdemData, TempestData
    tdem = isinstance(data, TdemData)
    tempest = isinstance(data, TempestData)
    s = np.asarray(res["scalars"], dtype=np.float64)
    m = s.shape[0]
    rows = np.arange(m) if rows is None else np.asarray(rows)
    nsys = 2 if opt.n_systems > 1 else 1
    G = _grids(opt)
    L = _lib

    def put(path, values):
        ds = parent[path]
        ds[rows] = np.asarray(values).astype(ds.dtype, copy=False)

    d = "data/"
    put(d + "x/data", data.x)
    put(d + "y/data", data.y)
    put(d + "elevation/data", data.elevation)
    put(d + "fiducial/data", data.fiducial)
    put(d + "line_number/data", data.lineNumber)
    obs = np.asarray(data.data, dtype=np.float64)
    put(d + "data/data", obs)
    # the best datapoint is written last over the same datasets (Inference1D.writeHdf :1079-1081): errors, height and
    # predicted data of the highest-posterior state; the posteriors stay those of the chain
    rel = s[:, [L.S_BEST_REL, L.S_BEST_REL2][:nsys]]
    add = s[:, [L.S_BEST_ADD, L.S_BEST_ADD2][:nsys]]
    put(d + "relative_error/data", rel if nsys > 1 else rel[:, 0])
    off = [sy.off_time for sy in data.system] if tdem else None
    if tempest:   # the sampled "additive error" is the multiplier of the channels' additive levels, one per component
        nt = data.system[0].nTimes
        level = np.asarray(data.additive_error, dtype=np.float64)
        put(d + "additive_error/data", level)
        put(d + "additive_error_multiplier/data", add)
        put(d + "std/data", np.sqrt((np.repeat(rel, nt, axis=1) * obs) ** 2 + (np.repeat(add, nt, axis=1) * level) ** 2))
    else:
        put(d + "additive_error/data", add if nsys > 1 else add[:, 0])
        put(d + "std/data", _std(tdem, obs, rel, add, off))
    if tempest:   # predicted data = predicted secondary + predicted primary field, which follows from the geometry alone
        pprim = np.tile(ops.tdem_primary_field(data.survey_struct()), (m, 1))
        put(d + "predicted_data/data", np.asarray(predicted_best) + np.repeat(pprim, nt, axis=1))
    else:
        put(d + "predicted_data/data", predicted_best)
    z_in = np.asarray(data.z, dtype=np.float64)
    z_best = s[:, L.S_BEST_HEIGHT] if opt.solve_height else z_in
    for key, hkey, gkey in (("relative_error", "rel_hist", "rel_edges"),
                            ("additive_error_multiplier" if tempest else "additive_error", "add_hist",
                             "add_edges_linear" if tempest else "add_edges")):
        hist = np.asarray(res[hkey]).reshape(m, nsys, -1)
        for i in range(nsys):
            p = d + key + ("/posterior%d" % i if nsys > 1 else "/posterior")
            put(p + "/values/data", hist[:, i])
            # (the reference hands every system's histogram the relative_to of the first one, StatArray.writeHdf)
            put(p + "/mesh/y/relative_to/data", np.full(m, G[gkey][0][1]))
    if tdem:
        put(d + "z/data", z_in)
        if tempest:   # data = secondary + primary per component
            put(d + "secondary_field/data", data.secondary_field)
            put(d + "predicted_secondary_field/data", predicted_best)
            put(d + "primary_field/data", data.primary_field)
            put(d + "predicted_primary_field/data", pprim)
        else:
            put(d + "secondary_field/data", obs)
            put(d + "predicted_secondary_field/data", predicted_best)
            put(d + "primary_field/data", np.zeros(m))
            put(d + "predicted_primary_field/data", np.zeros(m))
        tx, rx = data.transmitter, data.receiver
        lp = d + "loop_pair/"
        # a sampled transmitter height: the forward sees both loops move (Loop_pair.Geometry hands gatdaem1d the height
        # and the OFFSET, Loop_pair.py:62-78), the stored receiver loop keeps the z it was given
        dz = z_best - z_in
        put(lp + "x/data", rx["x"] - tx["x"])
        put(lp + "y/data", rx["y"] - tx["y"])
        put(lp + "z/data", rx["z"] - tx["z"])
        put(lp + "elevation/data", np.zeros(m))
        for name, q in (("transmitter", tx), ("receiver", rx)):
            for key in ("x", "y", "elevation", "pitch", "roll", "yaw", "moment", "radius"):
                put(lp + name + "/" + key + "/data", q[key])
            put(lp + name + "/z/data", q["z"] + (dz if name == "transmitter" else 0.0))
            put(lp + name + "/orientation/data", q["orientation"])
        if opt.solve_height:
            put(lp + "transmitter/z/posterior/values/data", res["height_hist"])
            put(lp + "transmitter/z/posterior/mesh/y/relative_to/data", s[:, L.S_HEIGHT_REF])
    else:
        put(d + "z/data", z_best)
        if opt.solve_height:
            put(d + "z/posterior/values/data", res["height_hist"])
            put(d + "z/posterior/mesh/y/relative_to/data", s[:, L.S_HEIGHT_REF])

    put("iteration", s[:, L.S_ITER])
    put("burned_in_iteration", s[:, L.S_BURNED_IN_ITER])
    put("best_iteration", s[:, L.S_BEST_ITER])
    put("burned_in", s[:, L.S_BURNED_IN] != 0)
    put("multiplier", np.full(m, multiplier))
    put("acceptance_rate/data", res["accept_trace"])
    put("phids/data", res["misfit_trace"])
    put("halfspace/data", s[:, L.S_HALFSPACE])

    put("model/mesh/nCells/data", s[:, L.S_BEST_K])
    put("model/mesh/nCells/posterior/values/data", res["ncells_hist"])
    put("model/mesh/y/edges/data", res["best_edges"])
    put("model/mesh/y/edges/posterior/values/data", res["edges_hist"])
    put("model/values/data", res["best_sigma"])
    put("model/values/posterior/values/data", res["hitmap"])
    put("model/values/posterior/mesh/y/relative_to/data", np.log10(s[:, L.S_HALFSPACE]))
    return parent


def save_line(filename, res, opt, data, predicted_best, **kwargs):
    """Create `<filename>` for one flight line (soundings sorted by fiducial, as the reference's files) and write it."""
    order = np.argsort(np.asarray(data.fiducial), kind="stable")
    rows = np.empty_like(order)
    rows[order] = np.arange(order.size)
    with h5().File(filename, "w") as f:
        create_line(f, order.size, opt, data, **{k: v for k, v in kwargs.items() if k != "multiplier"})
        write_line(f, res, opt, data, predicted_best, rows=rows, multiplier=kwargs.get("multiplier", 1.0))
    return filename


def read_line(filename_or_group):
    """The arrays of a line file back in the form `ops.rjmcmc_run` returns them (what `write_line` took): posterior histograms,
    traces, best models and the scalar columns that the file holds (best errors, heights, iterations, burn-in, half-space),
    plus `fiducial`, `line_number`, `x`, `y`, `z`, `elevation`, `data`, `predicted_data`, `std` and the bin `edges` of every
    posterior - enough for `dataset.summarise_device` and the per-line products without the reference installed.  Works on the
    files of the reference's own writer as well (same layout)."""
    f = h5().File(filename_or_group, "r") if isinstance(filename_or_group, str) else filename_or_group
    L = _lib

    def get(path):
        return np.asarray(f[path][()])
    n = int(get("data/fiducial/data").size)
    out = dict(hitmap=get("model/values/posterior/values/data"), edges_hist=get("model/mesh/y/edges/posterior/values/data"),
               ncells_hist=get("model/mesh/nCells/posterior/values/data"), misfit_trace=get("phids/data"),
               accept_trace=get("acceptance_rate/data"), best_sigma=get("model/values/data"), best_edges=get("model/mesh/y/edges/data"))
    for key in ("fiducial", "line_number", "x", "y", "z", "elevation", "data", "predicted_data", "std"):
        out[key] = get("data/%s/data" % key)
    s = np.zeros((n, L.NSCALARS))
    s[:, L.S_ITER], s[:, L.S_BURNED_IN_ITER], s[:, L.S_BEST_ITER] = get("iteration"), get("burned_in_iteration"), get("best_iteration")
    s[:, L.S_BURNED_IN], s[:, L.S_HALFSPACE], s[:, L.S_BEST_K] = get("burned_in"), get("halfspace/data"), get("model/mesh/nCells/data")
    s[:, L.S_BEST_HEIGHT] = s[:, L.S_HEIGHT_REF] = out["z"]
    tempest = "data/additive_error_multiplier" in f
    rel = get("data/relative_error/data").reshape(n, -1)
    add = get("data/additive_error_multiplier/data" if tempest else "data/additive_error/data").reshape(n, -1)
    s[:, L.S_BEST_REL], s[:, L.S_BEST_ADD] = rel[:, 0], add[:, 0]
    nsys = rel.shape[1]
    if nsys > 1:
        s[:, L.S_BEST_REL2], s[:, L.S_BEST_ADD2] = rel[:, 1], add[:, 1]
    hists = {}
    for key, name in (("relative_error", "rel_hist"), ("additive_error_multiplier" if tempest else "additive_error", "add_hist")):
        posts = ["data/%s/posterior%s" % (key, ("%d" % i) if nsys > 1 else "") for i in range(nsys)]
        hists[name] = np.stack([get(p_ + "/values/data") for p_ in posts], axis=1)
        out[name + "_edges"] = np.stack([get(p_ + "/mesh/y/edges/data") for p_ in posts])
    out["rel_hist"], out["add_hist"] = (hists["rel_hist"], hists["add_hist"]) if nsys > 1 else (hists["rel_hist"][:, 0], hists["add_hist"][:, 0])
    for zpath in ("data/z", "data/loop_pair/transmitter/z"):     # a sampled (transmitter) height and its posterior
        if zpath + "/posterior" in f:
            out["height_hist"] = get(zpath + "/posterior/values/data")
            s[:, L.S_BEST_HEIGHT] = get(zpath + "/data")
            s[:, L.S_HEIGHT_REF] = get(zpath + "/posterior/mesh/y/relative_to/data")
            out["height_edges"] = get(zpath + "/posterior/mesh/y/edges/data")
    out["scalars"] = s
    # the posteriors' bins as stored: conductivity bins are log10 and relative to each sounding's half-space value
    out["sigma_edges_log10_relative"] = get("model/values/posterior/mesh/y/edges/data")
    out["sigma_relative_to_log10"] = get("model/values/posterior/mesh/y/relative_to/data")
    out["depth_edges"] = get("model/values/posterior/mesh/z/edges/data")
    if tempest:
        out["additive_level"] = get("data/additive_error/data")
        for key in ("primary_field", "secondary_field", "predicted_primary_field", "predicted_secondary_field"):
            out[key] = get("data/%s/data" % key)
    return out
