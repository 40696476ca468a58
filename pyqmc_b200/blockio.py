"""Block files: the per-block output / restart files of ``vmc`` and ``rundmc``.

Layout = the reference's (``pyqmc/method/hdftools.py:19-53``, ``mc.py:92-99``, ``dmc.py:379-391``,
``coord.py:96-112, 224-252``): every key of a block dictionary is a dataset whose leading axis is
the block index (one row appended per block), file attributes hold run constants (``tstep``), and
the walker arrays (``configs``, periodic: ``wrap``; DMC: ``weights``) are REPLACED by the current
ones at every block, so the file always restarts the run.

Two interchangeable backends behind ``open_store``:

* ``Hdf5Store`` -- real HDF5 through ``h5py`` when that module is importable: files are then readable
  and continuable by the reference (same dataset names, resizable along axis 0);
* ``NpzStore`` -- no dependency beyond numpy (h5py is absent from the GPU boxes): one ``.npz``-format
  archive rewritten atomically per block (block rows are scalars and the walkers a few hundred KB,
  so a rewrite is cheap next to a block).  ``tools``-free conversion: ``to_hdf5(path_in, path_out)``.

The backend is chosen from the file's magic bytes when it exists, otherwise h5py if available.
"""
import io
import os
import tempfile

import numpy as np

WALKER_KEYS = ("configs", "wrap", "weights")
_ATTR = "__attr__"


def _have_h5py():
    try:
        import h5py  # noqa: F401
    except Exception:
        return False
    return hasattr(h5py, "File") and isinstance(h5py.File, type) and h5py.File is not object


def exists(path):
    return path is not None and os.path.isfile(path)


def _is_hdf5(path):
    with open(path, "rb") as f:
        return f.read(8) == b"\x89HDF\r\n\x1a\n"


def _walker_arrays(walkers, extra=None):
    """Arrays replaced at every block: the container's (ours or the reference's) plus e.g. DMC weights."""
    out = {}
    if walkers is not None:
        if hasattr(walkers, "arrays"):
            out.update(walkers.arrays())
        else:
            out.update({k: getattr(walkers, k) for k in WALKER_KEYS[:2] if hasattr(walkers, k)})
    out.update(extra or {})
    return out


class NpzStore:
    """Dictionary of arrays persisted as one zip archive (numpy's ``.npz`` format, any file name)."""

    def __init__(self, path, mode="a"):
        self.path, self.mode = path, mode
        self.data, self.attrs = {}, {}
        self.dirty = False
        if os.path.isfile(path):
            with np.load(path, allow_pickle=False) as z:
                for k in z.files:
                    if k.startswith(_ATTR):
                        self.attrs[k[len(_ATTR):]] = z[k][()]
                    else:
                        self.data[k] = z[k]
        elif mode == "r":
            raise FileNotFoundError(path)

    # mapping-style read access (what restart code and analysis scripts use)
    def __contains__(self, key):
        return key in self.data

    def keys(self):
        return self.data.keys()

    def __getitem__(self, key):
        return self.data[key]

    def last(self, key):
        return self.data[key][-1]

    def append_block(self, row, attrs=None, walkers=None, extra_walker_arrays=None):
        """Appends one block: ``row`` values gain a leading block axis; walker arrays are replaced."""
        if self.mode == "r":
            raise IOError("store opened read-only")
        for k, v in (attrs or {}).items():
            self.attrs.setdefault(k, np.asarray(v))
        for k, v in row.items():
            item = np.asarray(v)[None]
            self.data[k] = np.concatenate([self.data[k], item]) if k in self.data else item
        for k, v in _walker_arrays(walkers, extra_walker_arrays).items():
            self.data[k] = np.array(v)
        self.dirty = True

    def close(self):
        if not self.dirty:
            return
        payload = dict(self.data)
        payload.update({_ATTR + k: v for k, v in self.attrs.items()})
        folder = os.path.dirname(os.path.abspath(self.path))
        fd, tmp = tempfile.mkstemp(prefix=".blockio-", dir=folder)
        try:
            with os.fdopen(fd, "wb") as f:
                buf = io.BytesIO()
                np.savez(buf, **payload)
                f.write(buf.getvalue())
            os.replace(tmp, self.path)  # a reader never sees a half-written file
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
        self.dirty = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Hdf5Store:
    """The same interface over ``h5py``: datasets extendable along the block axis, walker datasets resized
    in place -- the files the reference itself writes and continues from."""

    def __init__(self, path, mode="a"):
        import h5py

        self.f = h5py.File(path, mode)
        self.attrs = self.f.attrs

    def __contains__(self, key):
        return key in self.f

    def keys(self):
        return self.f.keys()

    def __getitem__(self, key):
        return self.f[key][()]

    def last(self, key):
        return self.f[key][-1]

    def append_block(self, row, attrs=None, walkers=None, extra_walker_arrays=None):
        for k, v in (attrs or {}).items():
            if k not in self.f.attrs:
                self.f.attrs[k] = v
        for k, v in row.items():
            item = np.asarray(v)
            if k not in self.f:
                self.f.create_dataset(k, (0,) + item.shape, maxshape=(None,) + item.shape, dtype=item.dtype)
            ds = self.f[k]
            ds.resize(ds.shape[0] + 1, axis=0)
            ds[-1] = item
        for k, v in _walker_arrays(walkers, extra_walker_arrays).items():
            v = np.asarray(v)
            if k not in self.f:
                self.f.create_dataset(k, v.shape, chunks=True, maxshape=(None,) + v.shape[1:], dtype=v.dtype)
            self.f[k].resize(v.shape)
            self.f[k][...] = v

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_store(path, mode="a"):
    if os.path.isfile(path):
        return Hdf5Store(path, mode) if _is_hdf5(path) else NpzStore(path, mode)
    return Hdf5Store(path, mode) if _have_h5py() else NpzStore(path, mode)


def load_walkers(store, configs):
    """Restores the walker arrays of a checkpoint into ``configs`` in place (any container that has
    ``load_arrays`` -- ``pyqmc_b200.Walkers`` -- or plain ``.configs`` / ``.wrap`` arrays)."""
    stored = {k: store[k] for k in WALKER_KEYS[:2] if k in store}
    if hasattr(configs, "load_arrays"):
        configs.load_arrays(stored)
        return
    for k, v in stored.items():
        if hasattr(configs, k):
            if getattr(configs, k).shape == v.shape:
                getattr(configs, k)[...] = v
            else:
                setattr(configs, k, np.array(v, dtype=float))


def to_hdf5(path_in, path_out):
    """Rewrites an ``NpzStore`` file as reference-compatible HDF5 (needs h5py)."""
    import h5py

    src = NpzStore(path_in, "r")
    with h5py.File(path_out, "w") as f:
        for k, v in src.attrs.items():
            f.attrs[k] = v
        for k, v in src.data.items():
            walker = k in WALKER_KEYS
            f.create_dataset(k, data=v, maxshape=(None,) + v.shape[1:], chunks=True if walker else None)
