"""A small HDF5 container with the h5py calls the reference's result files use - for images without h5py / libhdf5.

The reference writes its results through h5py (`Inference1D.createHdf / writeHdf` inversion/Inference1D.py:1002-1090,
`DataArray.createHdf` classes/core/DataArray.py:1011-1097, `base/HDF/hdfWrite.py`).  This module provides the subset of
that API those code paths touch - `File`, `Group.create_group / create_dataset / get / __getitem__ / keys / attrs`,
`Dataset.shape / dtype / ndim / __getitem__ / __setitem__ / attrs` - over an in-memory tree, and serialises the tree to a
real HDF5 file on `close()` (`h5lite_format.py`: superblock v2, version-2 object headers with compact link storage,
contiguous datasets; the reader in the same module parses it back).  `geobipy_b200/hdf.py` writes the reference's result
layout against EITHER this module or the real h5py when it is installed (`hdf.h5()` picks).

Not a general HDF5 library: no chunking, compression, references, variable-length types other than UTF-8 string
attributes, or partial I/O - the tree lives in memory until the file is closed.
"""
import os

import numpy as np

__all__ = ["File", "Group", "Dataset", "AttributeManager"]


class AttributeManager(dict):
    """`obj.attrs`: a dict of strings / numbers / small arrays."""

    def create(self, name, data, **kwargs):
        self[name] = data

    def modify(self, name, value):
        self[name] = value


def _split(path):
    return [p for p in str(path).split("/") if p]


class _Node:
    def __init__(self, parent, key):
        self._parent, self._key = parent, key
        self.attrs = AttributeManager()

    @property
    def name(self):
        if self._parent is None:
            return "/"
        base = self._parent.name
        return base + self._key if base.endswith("/") else base + "/" + self._key

    @property
    def parent(self):
        return self._parent if self._parent is not None else self

    @property
    def file(self):
        n = self
        while n._parent is not None:
            n = n._parent
        return n


class Dataset(_Node):
    """A dataset held as one numpy array.  h5py semantics kept: shape () or (n, ...), `fillvalue`, `ds[...]` reads a
    copy-free view, `ds[...] = x` writes in place, `ds[()]` the whole array (a numpy scalar for shape ())."""

    def __init__(self, parent, key, array, fillvalue=None):
        super().__init__(parent, key)
        self._a = array
        self.fillvalue = fillvalue

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    ndim = property(lambda self: self._a.ndim)
    size = property(lambda self: self._a.size)
    nbytes = property(lambda self: self._a.nbytes)

    def __len__(self):
        if self._a.ndim == 0:
            raise TypeError("Attempt to take len() of scalar dataset")
        return self._a.shape[0]

    def __getitem__(self, idx):
        r = self._a[idx]
        return r[()] if isinstance(r, np.ndarray) and r.ndim == 0 else r

    def __setitem__(self, idx, value):
        if self._a.ndim == 0:
            self._a[()] = value
        else:
            self._a[idx] = value

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._a, dtype=dtype)

    def __iter__(self):
        return iter(self._a)

    def __repr__(self):
        return '<h5lite dataset "%s": shape %s, type "%s">' % (self._key, self.shape, self.dtype.str)

    def astype(self, dtype):
        return np.asarray(self._a).astype(dtype)

    def read_direct(self, dest, source_sel=None, dest_sel=None):
        dest[dest_sel if dest_sel is not None else Ellipsis] = self._a[source_sel if source_sel is not None else Ellipsis]


class Group(_Node):
    def __init__(self, parent=None, key=""):
        super().__init__(parent, key)
        self._children = {}

    # -- navigation
    def _walk(self, path, create=False):
        node = self.file if str(path).startswith("/") else self
        for p in _split(path):
            if not isinstance(node, Group):
                raise KeyError("%s is not a group" % node.name)
            if p not in node._children:
                if not create:
                    raise KeyError("Unable to open object (object '%s' doesn't exist)" % p)
                node._children[p] = Group(node, p)
            node = node._children[p]
        return node

    def __getitem__(self, path):
        return self._walk(path)

    def get(self, path, default=None):
        try:
            return self._walk(path)
        except KeyError:
            return default

    def __contains__(self, path):
        return self.get(path) is not None

    def __iter__(self):
        return iter(self._children)

    def __len__(self):
        return len(self._children)

    def keys(self):
        return self._children.keys()

    def values(self):
        return self._children.values()

    def items(self):
        return self._children.items()

    def __delitem__(self, path):
        parts = _split(path)
        del self._walk("/".join(parts[:-1]))._children[parts[-1]]

    def visititems(self, fn):
        def rec(g, prefix):
            for k, v in g._children.items():
                r = fn(prefix + k, v)
                if r is not None:
                    return r
                if isinstance(v, Group):
                    r = rec(v, prefix + k + "/")
                    if r is not None:
                        return r
        return rec(self, "")

    # -- creation
    def _place(self, path):
        parts = _split(path)
        if not parts:
            raise ValueError("empty name")
        g = self._walk("/".join(parts[:-1]), create=True) if len(parts) > 1 else (self.file if str(path).startswith("/") else self)
        if parts[-1] in g._children:
            raise ValueError("Unable to create link (name already exists): %s" % path)
        return g, parts[-1]

    def create_group(self, path, **kwargs):
        g, key = self._place(path)
        g._children[key] = Group(g, key)
        return g._children[key]

    def require_group(self, path):
        n = self.get(path)
        return n if n is not None else self.create_group(path)

    def create_dataset(self, path, shape=None, dtype=None, data=None, fillvalue=None, **kwargs):
        """h5py.Group.create_dataset for contiguous data: `data` (copied) or `shape` + `dtype` filled with `fillvalue`
        (0 when None; a NaN fill of an integer / bool dataset becomes 0 as it does in h5py's cast)."""
        g, key = self._place(path)
        if data is not None:
            if isinstance(data, (str, bytes)):
                a = np.asarray(data if isinstance(data, bytes) else data.encode("utf-8"), dtype="S")
            else:
                a = np.array(data, dtype=dtype, copy=True)
                if a.dtype == object or a.dtype.kind == "U":
                    a = np.char.encode(a.astype("U"), "utf-8")
            if shape is not None and tuple(np.atleast_1d(shape)) != a.shape and a.size == int(np.prod(shape)):
                a = a.reshape(shape)
        else:
            if shape is None:
                shape = ()
            shape = tuple(int(s) for s in np.atleast_1d(shape)) if not isinstance(shape, tuple) else tuple(int(s) for s in shape)
            dt = np.dtype(float if dtype is None else dtype)
            a = np.zeros(shape, dtype=dt)
            if fillvalue is not None:
                if dt.kind in "iub" and isinstance(fillvalue, float) and np.isnan(fillvalue):
                    fillvalue = 0
                a[...] = fillvalue
        g._children[key] = Dataset(g, key, a, fillvalue)
        return g._children[key]

    def __setitem__(self, path, value):
        self.create_dataset(path, data=value)

    def __repr__(self):
        return '<h5lite group "%s" (%d members)>' % (self.name, len(self._children))


class File(Group):
    """`h5py.File(name, mode)`: 'w' / 'x' create, 'r' / 'r+' / 'a' open an existing file (parsed into memory).  Writable
    files are serialised on `close()` / `flush()`.  `driver` / `comm` (MPI-IO in the reference) are accepted and ignored:
    one process writes a line's file here, after the end-of-run gather."""

    def __init__(self, name, mode="r", driver=None, comm=None, **kwargs):
        super().__init__(None, "")
        self.filename, self.mode = os.fspath(name), mode
        self._open = True
        if mode in ("r", "r+") or (mode == "a" and os.path.exists(self.filename)):
            from . import h5lite_format
            h5lite_format.read_into(self.filename, self)
        elif mode == "x" and os.path.exists(self.filename):
            raise FileExistsError(self.filename)
        elif mode not in ("w", "x", "a", "w-"):
            raise ValueError("invalid mode %r" % (mode,))

    def flush(self):
        if self.mode != "r":
            from . import h5lite_format
            h5lite_format.write(self.filename, self)

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __bool__(self):
        return self._open


# h5py spells the classes both ways (isinstance checks in the reference: h5py._hl.files.File, h5py._hl.group.Group)
class _NS:
    pass


_hl = _NS()
_hl.files, _hl.group, _hl.dataset = _NS(), _NS(), _NS()
_hl.files.File, _hl.group.Group, _hl.dataset.Dataset = File, Group, Dataset
