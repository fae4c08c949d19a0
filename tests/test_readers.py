"""The data-file readers on the files the REFERENCE ships, against the reference's own readers.

tests/golden/reader_files/ holds the heads of four data files of documentation_source/source/supplementary/data (a
comma-separated synthetic RESOLVE file, a whitespace-separated field file whose columns are named differently, the
one-sounding file, a SkyTEM dual-moment file) with their system files; tests/golden/readers.npz is what
FdemData.read_csv (classes/data/dataset/FdemData.py:520-610) and TdemData.read_csv (classes/data/dataset/TdemData.py:
401-560) of the unmodified reference returned for exactly those files (tests/golden/make_golden.py readers)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = os.path.join(HERE, "golden", "reader_files")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "golden", "readers.npz"))


@pytest.mark.parametrize("name,stm", [("resolve_glacial.csv", "resolve.stm"), ("Resolve_small.txt", "FdemSystem1.stm"),
                                      ("Resolve_single.txt", "FdemSystem1.stm")])
def test_fdem_reader_on_the_reference_files(ref, name, stm):
    from geobipy_b200.dataset import FdemData
    d = FdemData.read_csv(os.path.join(FILES, name), os.path.join(FILES, stm))
    key = name.split(".")[0]
    assert d.nPoints == ref[key + "/data"].shape[0] and d.nChannels == ref[key + "/data"].shape[1]
    for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data", "std"):
        got = np.asarray(getattr(d, k), dtype=np.float64)
        assert got.shape == ref[key + "/" + k].shape and np.array_equal(got, ref[key + "/" + k], equal_nan=True), (name, k)
    assert np.array_equal(d.system.frequencies, ref[key + "/frequencies"])
    assert np.allclose(d.system.loop_separation, ref[key + "/loop_separation"], rtol=1e-15)


def test_tdem_reader_on_the_reference_file(ref):
    from geobipy_b200.tdem import TdemData
    d = TdemData.read_csv(os.path.join(FILES, "skytem_glacial.csv"), [os.path.join(FILES, "SkytemHM.stm"), os.path.join(FILES, "SkytemLM.stm")])
    key = "skytem_glacial/"
    assert d.nPoints == 6 and d.nChannels == 45
    for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data", "std"):
        got = np.asarray(getattr(d, k), dtype=np.float64)
        assert got.shape == ref[key + k].shape and np.array_equal(got, ref[key + k], equal_nan=True), k
    for who in ("transmitter", "receiver"):
        lp = getattr(d, who)
        for k in ("x", "y", "z", "pitch", "roll", "yaw", "radius"):
            assert np.array_equal(np.asarray(lp[k], dtype=np.float64), ref[key + who + "_" + k]), (who, k)
    assert np.allclose(d.system[0].off_time, ref[key + "off_time0"], rtol=1e-15) and np.allclose(d.system[1].off_time, ref[key + "off_time1"], rtol=1e-15)


def test_tempest_reader_on_the_reference_file():
    """TempestData.read_csv on the head of the Tempest file the reference ships, against the reference's own TempestData.read_csv
    (classes/data/dataset/TempestData.py:140-273; tests/golden/make_golden.py readers_tempest): secondary field per component,
    primary field PX / PZ, data = secondary + primary, the loops, and the reader's all-zero error arrays."""
    from geobipy_b200.tdem import TempestData, Tempest_datapoint
    ref = np.load(os.path.join(HERE, "golden", "readers_tempest.npz"))
    d = TempestData.read_csv(os.path.join(FILES, "tempest_glacial.csv"), os.path.join(FILES, "tempest.stm"))
    assert d.nPoints == 6 and d.nChannels == 30 and list(d.system[0].components) == list(ref["components"])
    for k in ("lineNumber", "fiducial", "x", "y", "z", "elevation", "data", "std", "primary_field", "secondary_field", "relative_error",
              "additive_error", "additive_error_multiplier"):
        got = np.asarray(getattr(d, k), dtype=np.float64)
        assert got.shape == ref[k].shape and np.array_equal(got, ref[k], equal_nan=True), k
    for who in ("transmitter", "receiver"):
        lp = getattr(d, who)
        for k in ("x", "y", "z", "pitch", "roll", "yaw", "radius"):
            assert np.array_equal(np.asarray(lp[k], dtype=np.float64), ref[who + "_" + k]), (who, k)
    assert np.allclose(d.system[0].off_time, ref["off_time0"], rtol=1e-15)
    # errors set on the data set: its own std (the getter inherited from TdemData) and the datapoint's (Tempest_datapoint.std)
    d.relative_error = np.tile([0.001, 0.002], (6, 1))
    d.additive_error = np.tile(np.linspace(0.01, 0.02, 30), (6, 1))
    assert np.allclose(d.std, ref["std_with_errors"], rtol=1e-14)
    dp = d.datapoint(2)
    assert isinstance(dp, Tempest_datapoint) and np.array_equal(dp.data, ref["dp2_data"]) and np.allclose(dp.std, ref["dp2_std"], rtol=1e-14)
