"""Survey-level reader / driver (geobipy_b200/dataset.py).  The CSV layout is the reference's
(tests/data_checks/resolve_*_clean.csv, documentation_source/source/supplementary/data/resolve_*.csv)."""
import os

import numpy as np
import pytest

HEADER = ("Line_number,Fiducial,Easting,Northing,Height,Elevation,In_Phase_380.0,In_Phase_1776.0,In_Phase_3345.0,"
          "In_Phase_8171.0,In_Phase_41020.0,In_Phase_129550.0,Quadrature_380.0,Quadrature_1776.0,Quadrature_3345.0,"
          "Quadrature_8171.0,Quadrature_41020.0,Quadrature_129550.0")
STM = ("freq, tor, tmom, tx, ty, tzoff, ror, rmom, rx, ry, rzoff\n380, z, 1, 0, 0, 0, z, 1, 7.93, 0, 0\n"
       "1776, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0\n3345, x, -1, 0, 0, 0, x, 1, 9.03, 0, 0\n8171, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0\n"
       "41020, z, 1, 0, 0, 0, z, 1, 7.91, 0, 0\n129550, z, 1, 0, 0, 0, z, 1, 7.89, 0, 0\n")


def _write_survey(tmp_path, golden_dir, n=6):
    g = np.load(os.path.join(golden_dir, "resolve_clean.npz"))
    stm = tmp_path / "resolve.stm"
    stm.write_text(STM)
    rows = [HEADER]
    for i in range(n):
        line = 100.0 if i < n // 2 else 200.0
        vals = [line, i, float(i), 0.0, 30.0, 0.0] + list(g["data"][0, i * 10])
        rows.append(",".join(repr(float(v)) if j != 1 else str(int(v)) for j, v in enumerate(vals)))
    rows[2] = rows[2].replace(rows[2].split(",")[8], "NaN")   # one missing channel
    csvf = tmp_path / "resolve_glacial_clean.csv"
    csvf.write_text("\n".join(rows) + "\n")
    return str(csvf), str(stm), g


def test_read_csv_reference_layout(tmp_path, golden_dir, built_lib):
    from geobipy_b200.dataset import FdemData
    csvf, stm, g = _write_survey(tmp_path, golden_dir)
    d = FdemData.read_csv(csvf, stm)
    assert d.nPoints == 6 and d.nChannels == 12 and d.data.shape == (6, 12)
    assert list(d.lines) == [100.0, 200.0] and list(d.line(200.0)) == [3, 4, 5]
    assert np.array_equal(d.fiducial, np.arange(6.0)) and np.all(d.z == 30.0) and np.all(d.elevation == 0.0)
    assert np.isnan(d.data[1, 2]) and np.allclose(d.data[0], g["data"][0, 0])
    # FdemData.read_csv stores 0.1 x data when the file has no error columns (:580-583), but the `std` GETTER of the
    # reference recomputes from the data set's relative error of 1 % (Data.std :376-384; pinned on the reference's own
    # reader in tests/test_readers.py)
    assert np.allclose(d.std_from_file[0], 0.1 * d.data[0]) and np.allclose(d.std[0], 0.01 * d.data[0])
    dp = d.datapoint(3)
    assert dp.fiducial == 3.0 and dp.lineNumber == 200.0 and dp.n_active_channels == 12
    assert d.datapoint(1).n_active_channels == 11             # NaN channel is inactive (EmDataPoint.active)
    with pytest.raises(AssertionError):                        # header without a fiducial column
        bad = tmp_path / "bad.csv"
        bad.write_text(HEADER.replace("Fiducial", "Foo") + "\n" + "0,0,0,0,30,0," + ",".join(["1"] * 12) + "\n")
        FdemData.read_csv(str(bad), stm)


def test_summaries_match_histogram_class(built_lib):
    """dataset.summarise (batched) against api.Histogram (per sounding; mirrors Mesh._mean/_percentile)."""
    from geobipy_b200 import api, dataset, ops, _lib
    opt = ops.make_options()
    rng = np.random.default_rng(1)
    nd = ops.n_depth(opt)
    hm = rng.integers(0, 30, (2, opt.n_sigma_bins, nd)).astype(np.int32)
    sc = np.zeros((2, _lib.NSCALARS))
    sc[:, _lib.S_HALFSPACE] = [0.01, 0.2]
    res = dict(hitmap=hm, scalars=sc, edges_hist=rng.integers(0, 5, (2, nd)).astype(np.int32))
    s = dataset.summarise(res, opt)
    for b in range(2):
        g = ops.posterior_grids(opt, sc[b, _lib.S_HALFSPACE])
        h = api.Histogram(hm[b], g["sigma_edges"], g["depth_edges"], log_x=True)
        assert np.allclose(s["mean"][b], h.mean()) and np.allclose(s["p50"][b], h.median())
        assert np.allclose(s["p5"][b], h.percentile(5.0)) and np.allclose(s["p95"][b], h.percentile(95.0))
        assert abs(s["interface_probability"][b].sum() - 1.0) < 1e-12
    assert "height_mean" not in s
    # sampled sensor height (solve_z): posterior mean height on the prior's bins (Point.set_z_posterior :1013-1020),
    # which are centred on S_HEIGHT_REF (the input height, re-centred by every reset())
    optz = ops.make_options(solve_z=True, maximum_z_change=2.0, z_proposal_variance=0.01)
    hh = np.zeros((2, 99), np.int32)
    hh[0, 49] = 7                     # the centre bin: z = z0
    hh[1, 98], hh[1, 97] = 3, 1       # the two top bins
    scz = sc.copy()
    scz[:, _lib.S_HEIGHT_REF] = [30.0, 31.25]   # sounding 1 went through a reset(): prior re-centred 1.25 m higher
    sz = dataset.summarise(dict(res, scalars=scz, height_hist=hh), optz)
    e = np.linspace(-2.0, 2.0, 100)
    assert abs(sz["height_mean"][0] - 30.0) < 1e-12
    assert np.isclose(sz["height_mean"][1], 31.25 + (3 * 0.5 * (e[98] + e[99]) + 0.5 * (e[97] + e[98])) / 4)
    hz = api.Histogram(hh[1], 31.25 + e)
    assert np.isclose(hz.mean(), sz["height_mean"][1])


def test_histogram_mirror_matches_reference_histogram(golden_dir):
    """api.Histogram against what the reference's own Histogram / Mesh methods return on recorded hitmaps
    (tests/golden/posterior_summaries.npz, make_golden.py summaries): mean, median, mode, percentiles, credible range,
    transparency, opacity, opacity level - including the reference's round-off at exact percentile ties (sounding 2 of the
    file has 10 000 counts per column and cumulative fractions of exactly 0.95)."""
    from geobipy_b200 import api
    g = np.load(os.path.join(golden_dir, "posterior_summaries.npz"))
    ties = 0
    for n in range(3):
        P = lambda k: g["s%d_%s" % (n, k)]   # noqa: E731
        h = api.Histogram(P("hitmap"), P("x_edges"), P("y_edges"), log_x=True)
        cs = np.cumsum(P("hitmap"), axis=0).astype(np.float64)
        for name, mine in (("mean", h.mean()), ("median", h.median()), ("mode", h.mode())):
            assert np.allclose(mine, P(name), rtol=1e-12, atol=0.0), (n, name)
        # Histogram.percentile works on the pmf (counts / grand total, Histogram.py:43-49, :387): where a cumulative count
        # hits percent x total EXACTLY its answer depends on the round-off of that normalisation (bin i or i + 1); median,
        # credible range, transparency and opacity work on the counts and are reproduced exactly, ties included
        for name, pc, mine in (("p5", 5.0, h.percentile(5.0)), ("p95", 95.0, h.percentile(95.0))):
            tie = (np.abs(cs - pc * 0.01 * cs[-1]) < 1e-9 * np.maximum(cs[-1], 1.0)).any(axis=0)
            assert np.allclose(mine[~tie], P(name)[~tie], rtol=1e-12, atol=0.0), (n, name)
            bw = np.log(P("x_edges")[1] / P("x_edges")[0])
            assert np.all(np.abs(np.log(mine[tie] / P(name)[tie])) <= bw * (1 + 1e-9)), (n, name)
        assert np.allclose(h.credible_range(90.0), P("credible_range90"), rtol=0.0, atol=1e-12), n
        assert np.allclose(h.transparency(90.0), P("transparency90"), rtol=0.0, atol=1e-12), n
        assert np.allclose(h.opacity(90.0), P("opacity90"), rtol=0.0, atol=1e-12), n
        assert np.allclose(h.transparency(95.0), P("transparency95"), rtol=0.0, atol=1e-12), n
        assert h.opacity_level(95.0) == float(P("opacity_level95")), n
        cs = np.cumsum(P("hitmap"), axis=0)
        ties += int(((cs == 0.95 * cs[-1]) & (cs[-1] > 0)).any())
    assert ties >= 1   # the tie rule is exercised


@pytest.mark.gpu
def test_inference3d_end_to_end(tmp_path, golden_dir, built_lib):
    from geobipy_b200 import _lib
    from geobipy_b200.dataset import FdemData, Inference3D
    _lib.require_cuda()
    csvf, stm, g = _write_survey(tmp_path, golden_dir)
    d = FdemData.read_csv(csvf, stm)
    inv = Inference3D(d, seed=5)
    r = inv.infer(n_markov_chains=400, max_iterations=300)
    assert r["hitmap"].shape == (6, 250, 440) and (r["scalars"][:, _lib.S_ITER] == 300).all()
    assert r["summary_p50"].shape == (6, 440) and np.all(r["summary_p5"] <= r["summary_p95"])
    # the summaries were computed on the device (gbp_summarise_posterior / gbp_opacity_doi): the host statement agrees
    from geobipy_b200 import dataset
    host = dataset.summarise(r, inv.options, line_id=d.lineNumber)
    for k in ("mean", "p5", "p50", "p95", "mode", "interface_probability"):
        assert np.allclose(r["summary_" + k], host[k], rtol=1e-12, atol=0.0), k
    assert np.allclose(r["summary_credible_range"], host["credible_range"], rtol=0.0, atol=1e-12)
    assert np.allclose(r["summary_opacity"], host["opacity"], rtol=0.0, atol=1e-12) and np.array_equal(r["summary_doi"], host["doi"])
    assert r["summary_opacity"][:3].max() == 1.0 and r["summary_opacity"][3:].max() == 1.0   # normalised per flight line
    one = Inference3D(d, seed=5).infer(index=4, n_markov_chains=400, max_iterations=300)
    assert np.array_equal(one["hitmap"][0], r["hitmap"][4])       # (seed, sounding index) fixes the stream
    byfid = Inference3D(d, seed=5).infer(fiducial=4.0, line_number=200.0, n_markov_chains=400, max_iterations=300)
    assert np.array_equal(byfid["scalars"], one["scalars"])
    files = inv.save(str(tmp_path / "out"), format="npz")
    assert [os.path.basename(f) for f in files] == ["100.npz", "200.npz"]
    z = np.load(files[1])
    assert z["hitmap"].shape == (3, 250, 440) and list(z["fiducial"]) == [3.0, 4.0, 5.0]
    # the reference's own output: one HDF5 file per line in the layout of Inference2D.createHdf / Inference1D.writeHdf
    # (tests/test_hdf.py pins the layout on the reference's own writer and readers)
    from geobipy_b200 import api, h5lite
    files = inv.save(str(tmp_path / "out"))
    assert [os.path.basename(f) for f in files] == ["100.h5", "200.h5"]
    f = h5lite.File(files[1], "r")
    assert f["model/values/posterior/values/data"].shape == (3, 250, 440) and list(f["data/fiducial/data"][()]) == [3.0, 4.0, 5.0]
    assert np.array_equal(f["model/values/posterior/values/data"][()], r["hitmap"][3:]) and f.attrs == {} and f["data"].attrs["repr"] == "FdemData"
    assert np.array_equal(f["iteration"][()], r["scalars"][3:, _lib.S_ITER]) and f["model/mesh/nCells/data"][1] == r["scalars"][4, _lib.S_BEST_K]
    assert np.array_equal(f["phids/data"][()], r["misfit_trace"][3:]) and f["acceptance_rate/data"].dtype == np.uint8
    kb = int(r["scalars"][4, _lib.S_BEST_K])
    best = api.Model(api.RectilinearMesh1D(edges=r["best_edges"][4, :kb + 1]), r["best_sigma"][4, :kb])
    dp = d.datapoint(4)
    dp.forward(best)     # predicted data of the best model, as best_datapoint.writeHdf leaves them
    assert np.allclose(f["data/predicted_data/data"][1], dp.predictedData, rtol=1e-10)
    assert np.allclose(f["data/std/data"][1], np.sqrt((r["scalars"][4, _lib.S_BEST_REL] * d.data[4]) ** 2 + r["scalars"][4, _lib.S_BEST_ADD] ** 2))


def test_opacity_and_doi_of_a_line():
    """Inference2D.compute_opacity (:1011-1023) / compute_doi (:493-532) restated on the p5 / p95 summaries: a hand
    example with 2 soundings x 5 depth cells."""
    from geobipy_b200.dataset import opacity_and_doi
    p5 = np.array([[1e-2, 1e-2, 1e-2, 1e-3, 1e-4], [1e-2, 1e-2, 1e-2, 1e-2, 1e-2]])
    p95 = np.array([[2e-2, 2e-2, 1e-1, 1e-1, 1e0], [2e-2, 2e-2, 2e-2, 2e-2, np.nan]])
    edges = np.arange(6) * 10.0
    op, doi = opacity_and_doi(p5, p95, edges)
    rng = np.abs(np.log10(p95) - np.log10(p5))              # decades: [0.301, 0.301, 1, 2, 4], [0.301 x4, nan]
    t = (rng - np.log10(2.0)) / (4.0 - np.log10(2.0))       # normalised over the whole line
    t[1, 4] = 1.0                                           # NaN -> fully transparent
    assert np.allclose(op, 1.0 - t)
    # sounding 0: opacity [1, 1, 0.811, 0.541, 0] -> deepest cell with opacity >= 0.67 is cell 2 (centre 25 m);
    # sounding 1: cell 4 is transparent, cell 3 opaque -> 35 m
    assert np.allclose(doi, [25.0, 35.0])
    # nothing reaches the cut-off below the top: the search stops at the top cell (compute_doi's j >= 1 guard)
    op2, doi2 = opacity_and_doi(np.array([[1e-3, 1e-4, 1e-4]]), np.array([[1e-3, 1e0, 1e0]]), np.arange(4) * 1.0, doi_percent=100.0)
    assert np.allclose(op2, [[1.0, 0.0, 0.0]]) and doi2[0] == 0.5
    # a line with one and the same credible range everywhere is fully opaque
    op3, doi3 = opacity_and_doi(np.full((2, 3), 1e-2), np.full((2, 3), 1e-1), np.arange(4) * 2.0)
    assert np.allclose(op3, 1.0) and np.allclose(doi3, 5.0)


def test_inference3d_save_writes_line_products(tmp_path, golden_dir, built_lib):
    """Inference3D.save on fabricated results (no GPU): one file per line with the per-line opacity / doi."""
    from geobipy_b200 import _lib, dataset, ops
    csvf, stm, g = _write_survey(tmp_path, golden_dir)
    d = dataset.FdemData.read_csv(csvf, stm)
    opt = ops.make_options()
    nd = ops.n_depth(opt)
    rng = np.random.default_rng(2)
    sc = np.zeros((6, _lib.NSCALARS))
    sc[:, _lib.S_HALFSPACE] = 0.02
    res = dict(hitmap=rng.integers(0, 20, (6, opt.n_sigma_bins, nd)).astype(np.int32), scalars=sc,
               edges_hist=rng.integers(0, 5, (6, nd)).astype(np.int32), index=np.arange(6))
    res.update({"summary_" + k: v for k, v in dataset.summarise(res, opt, line_id=d.lineNumber).items()})
    inv = dataset.Inference3D(d, seed=1)
    inv.results, inv.options = res, opt
    files = inv.save(str(tmp_path / "out"), format="npz")
    assert [os.path.basename(f) for f in files] == ["100.npz", "200.npz"]
    z = np.load(files[0])
    assert z["hitmap"].shape == (3, opt.n_sigma_bins, nd) and z["opacity"].shape == (3, nd) and z["doi"].shape == (3,)
    assert z["opacity"].min() >= 0.0 and z["opacity"].max() <= 1.0 and z["opacity"].max() == 1.0
    op, doi = dataset.opacity_and_doi(res["summary_p5"][:3], res["summary_p95"][:3], res["summary_depth_edges"])
    assert np.allclose(z["opacity"], op, rtol=0.0, atol=1e-12) and np.array_equal(z["doi"], doi)


def test_histogram_credible_range_opacity_mode(built_lib):
    """The per-sounding summaries of the reference's Histogram class (mode, credible intervals / range, transparency,
    opacity: Histogram.py:113-127, 308-367, 509-541; Mesh.py:30-78, 138-165) on the mirror, against direct numpy and
    against the per-line opacity_and_doi when the line is that one sounding."""
    from geobipy_b200 import api, dataset, ops
    opt = ops.make_options()
    nd = ops.n_depth(opt)
    rng = np.random.default_rng(7)
    # a posterior that widens with depth
    hm = np.zeros((opt.n_sigma_bins, nd), np.int32)
    for j in range(nd):
        w = 2 + j // 8
        hm[:, j] = np.bincount(np.clip(rng.normal(120, w, 400).astype(int), 0, opt.n_sigma_bins - 1), minlength=opt.n_sigma_bins)
    g = ops.posterior_grids(opt, 0.03)
    h = api.Histogram(hm, g["sigma_edges"], g["depth_edges"], log_x=True)
    c = 0.5 * (np.log(g["sigma_edges"])[1:] + np.log(g["sigma_edges"])[:-1])
    assert np.allclose(np.log(h.mode()), c[np.argmax(hm, axis=0)])
    med, lo, hi = h.credible_intervals(90.0)
    assert np.array_equal(med, h.median()) and np.array_equal(lo, h.percentile(5.0)) and np.array_equal(hi, h.percentile(95.0))
    assert np.all(lo <= med) and np.all(med <= hi)
    r = h.credible_range(90.0)
    assert np.allclose(r, np.log10(hi / lo)) and r[-1] > r[0]
    t, op = h.transparency(90.0), h.opacity(90.0)
    assert t.min() == 0.0 and t.max() == 1.0 and np.allclose(op, 1.0 - t)
    op_line, doi = dataset.opacity_and_doi(lo[None, :], hi[None, :], g["depth_edges"])
    assert np.allclose(op_line[0], op)
    yc = 0.5 * (g["depth_edges"][1:] + g["depth_edges"][:-1])
    assert doi[0] == yc[np.flatnonzero(op >= 0.67).max()]
    lvl = h.opacity_level(50.0)   # the reference uses `percent` for the interval AND for the threshold
    assert lvl == yc[np.flatnonzero(h.transparency(50.0) <= 0.5).max()]
    # 1-D histograms (error / height posteriors)
    h1 = api.Histogram(np.array([0, 1, 5, 2, 0]), np.linspace(0.0, 5.0, 6))
    assert h1.mode() == 2.5 and h1.credible_range(50.0) == 0.0 and h1.credible_range(90.0) == 2.0
