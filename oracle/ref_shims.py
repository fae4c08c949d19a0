"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Import shims that let the *unmodified* reference package (DOI-USGS/geobipy, mounted
read-only at /root/reference in the build container) be imported and driven without its
plotting / HDF5 / MPI dependencies, none of which exist in this image.

Used only by
  * tests/golden/make_golden.py   (generates the committed golden vectors),
  * oracle validation scripts run in the build container.
The GPU box has no /root/reference; nothing under tests -m gpu, smoke() or bench.py
imports this file.

The stubs only have to satisfy import-time attribute look-ups of
geobipy/src/base/plotting.py:19-74 and friends; no plotting call is ever made.
"""
import functools
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GEOBIPY_REFERENCE", "/root/reference")


class _Anything:
    """Object that swallows any attribute access / call (for plotting stubs)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __contains__(self, item):
        return False

    def __getitem__(self, item):
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        # class-like names must be real classes (used in isinstance / subclassing)
        if name[:1].isupper():
            cls = type(name, (object,), {"__init__": lambda self, *a, **k: None})
            setattr(self, name, cls)
            return cls
        val = _Anything()
        setattr(self, name, val)
        return val


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = _StubModule(name)
    m.__path__ = []  # behave like a package so sub-imports work
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serve a stub for any sub-module of an already stubbed top-level package."""

    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if isinstance(sys.modules.get(top), _StubModule):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_STUBBED = [
    "h5py",
    "matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.gridspec",
    "matplotlib.axes", "matplotlib.cm", "matplotlib.colors", "matplotlib.collections",
    "matplotlib.patches", "matplotlib.colorbar", "matplotlib.ticker", "matplotlib.lines",
    "matplotlib.animation", "matplotlib.pylab", "matplotlib.path", "matplotlib.markers",
    "matplotlib.backends", "matplotlib.backends.backend_pdf", "matplotlib.dates",
    "matplotlib.image", "matplotlib.transforms", "matplotlib.widgets",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.mplot3d",
    "progressbar", "pyvista", "pygmt", "numba_kdtree", "lmfit", "netCDF4", "empymod",
    "mpi4py", "sklearn.mixture",
]


def install():
    """Install the stubs (idempotent) and put the reference on sys.path."""
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    for name in _STUBBED:
        try:
            if name.split(".")[0] in ("sklearn",):
                importlib.import_module(name)
                continue
        except Exception:
            pass
        try:
            if name not in sys.modules:
                importlib.import_module(name)
        except Exception:
            _stub(name)

    # import-time touches of geobipy/src/base/plotting.py:19-74
    mpl = sys.modules.get("matplotlib")
    if isinstance(mpl, _StubModule):
        class _Colormaps(dict):
            def __contains__(self, item):
                return True  # make_colourmap() returns early

            def register(self, *a, **k):
                pass

        mpl.colormaps = _Colormaps()
        mc = sys.modules["matplotlib.colors"]
        mc.hex2color = lambda h: (0.0, 0.0, 0.0)
        mc.to_rgba = lambda c, *a, **k: (0.0, 0.0, 0.0, 1.0)
        mc.ListedColormap = type("ListedColormap", (object,), {"__init__": lambda self, *a, **k: None})

    # cached_property -> functools
    if "cached_property" not in sys.modules:
        cp = types.ModuleType("cached_property")
        cp.cached_property = functools.cached_property
        sys.modules["cached_property"] = cp

    # SciPy >= 1.18 moved a private helper the reference imports by name
    # (geobipy/src/base/interpolation.py:11).
    try:
        import scipy.interpolate as si
        try:
            from scipy.interpolate.interpnd import _ndim_coords_from_arrays  # noqa: F401
        except Exception:
            from scipy.interpolate import _interpnd
            mod = types.ModuleType("scipy.interpolate.interpnd")
            mod._ndim_coords_from_arrays = _interpnd._ndim_coords_from_arrays
            for k in dir(_interpnd):
                if not hasattr(mod, k):
                    try:
                        setattr(mod, k, getattr(_interpnd, k))
                    except Exception:
                        pass
            sys.modules["scipy.interpolate.interpnd"] = mod
            si.interpnd = mod
    except Exception:
        pass

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Return the imported reference package (container only)."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree %s is not present on this machine" % REFERENCE_ROOT)
    install()
    return importlib.import_module("geobipy")


def load_numba_kernels():
    """Load fdem1d_numba.py standalone by path (it imports only numpy + numba)."""
    path = os.path.join(REFERENCE_ROOT, "geobipy/src/classes/forwardmodelling/Electromagnetic/FD/fdem1d_numba.py")
    spec = importlib.util.spec_from_file_location("ref_fdem1d_numba", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
