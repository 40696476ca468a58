"""Host walker containers.

One class, ``Walkers``, serves open and periodic boundary conditions: it carries a tuple of
per-walker arrays (``configs (N, nelec, 3)`` and, with a lattice, the integer ``wrap`` vectors) and
applies every container operation of the driver protocol to all of them, so the periodic case is not
a second copy of the open one.  The protocol is the reference's (``pyqmc/configurations/coord.py``:
``electron, mask, make_irreducible, move, resample, split, join, copy, reshape`` and the checkpoint
hooks), which is what ``pyqmc.method.mc.vmc`` / ``dmc.rundmc`` call; the reference's own
``OpenConfigs`` / ``PeriodicConfigs`` are equally accepted by every ``pyqmc_b200`` object (only
``.configs`` / ``.wrap`` are read).  ``OpenConfigs(...)`` / ``PeriodicConfigs(...)`` below are
constructors with the reference's argument order.

Walkers stay C-contiguous float64 on the host; the device keeps its own copy.
"""
import copy as _copy

import numpy as np


def wrap_into_cell(lattice, positions):
    """Positions folded into the cell spanned by the rows of ``lattice`` and the integer number of
    lattice vectors removed -- same operations, in the same order, as ``enforce_pbc``
    (``pyqmc/pbc/pbc.py:33-47``) so that wrapped coordinates agree to the bit."""
    fractional = np.einsum("...ij,jk->...ik", positions, np.linalg.inv(lattice))
    whole, rest = np.divmod(fractional, 1)
    return np.dot(rest, lattice), whole


class ElectronView:
    """Trial positions of one electron (``(N, 3)``) or of its auxiliary points (``(N, naip, 3)``)."""

    __slots__ = ("configs", "wrap", "lvec", "dist")

    def __init__(self, positions, lattice=None, wrap=None, dist=None):
        self.configs, self.lvec, self.dist = positions, lattice, dist
        if lattice is None:
            self.wrap = None
        else:
            self.wrap = np.zeros_like(positions) if wrap is None else wrap

    def mask(self, keep):
        sub = None if self.wrap is None else self.wrap[keep]
        return ElectronView(self.configs[keep], self.lvec, sub, self.dist)

    def select_electrons(self, index):
        """Auxiliary point ``index`` of every walker (used by the reference's harness, testwf.py:86)."""
        sub = None if self.wrap is None else self.wrap[:, index]
        return ElectronView(self.configs[:, index], self.lvec, sub, self.dist)


class Walkers:
    """Walker ensemble; ``lattice=None`` means open boundary conditions."""

    def __init__(self, configs, lattice=None, wrap=None, dist=None, fold=True):
        self.lvecs = None if lattice is None else np.asarray(lattice, dtype=float)
        self.dist = dist
        if self.lvecs is None:
            self.configs = configs
            return
        if fold:
            self.configs, self.wrap = wrap_into_cell(self.lvecs, configs)
            if wrap is not None:
                self.wrap += wrap
        else:
            self.configs = configs
            self.wrap = np.zeros_like(configs) if wrap is None else wrap

    # -- the arrays every operation is applied to --------------------------------------------------
    @property
    def periodic(self):
        return self.lvecs is not None

    def _names(self):
        return ("configs", "wrap") if self.periodic else ("configs",)

    def _like(self, **arrays):
        """New container with the same lattice whose arrays are taken as given (no re-folding)."""
        return Walkers(arrays["configs"], self.lvecs, arrays.get("wrap"), self.dist, fold=False)

    def _map(self, fn):
        return self._like(**{k: fn(getattr(self, k)) for k in self._names()})

    # -- views ----------------------------------------------------------------------------------
    def electron(self, index):
        w = self.wrap[:, index] if self.periodic else None
        return ElectronView(self.configs[:, index], self.lvecs, w, self.dist)

    def select_electrons(self, indices):
        return self._map(lambda a: a[:, indices])

    def mask(self, keep):
        return self._map(lambda a: a[keep])

    def make_irreducible(self, e, vec, mask=None):
        """Trial positions ``vec`` of electron ``e`` folded into the cell; their wrap vectors continue
        electron ``e``'s (``coord.py:168-194``).  Open boundaries: the positions as they are."""
        if not self.periodic:
            return ElectronView(vec, dist=self.dist)
        base = self.wrap[:, e]
        if vec.ndim == 3:
            base = np.broadcast_to(base[:, None, :], vec.shape)
        wrap = np.array(base)
        if mask is None:
            folded, shift = wrap_into_cell(self.lvecs, vec)
            wrap += shift
        else:
            folded = vec.copy()
            folded[mask], shift = wrap_into_cell(self.lvecs, vec[mask])
            wrap[mask] += shift
        return ElectronView(folded, self.lvecs, wrap, self.dist)

    # -- in-place updates -------------------------------------------------------------------------
    def move(self, index, trial, accept):
        accept = np.asarray(accept, dtype=bool)
        for k in self._names():
            getattr(self, k)[accept, index] = getattr(trial, k)[accept]

    def resample(self, picked):
        for k in self._names():
            setattr(self, k, getattr(self, k)[picked])

    def reshape(self, shape):
        for k in self._names():
            setattr(self, k, getattr(self, k).reshape(shape))

    def join(self, parts, axis=0):
        for k in self._names():
            setattr(self, k, np.concatenate([getattr(p, k) for p in parts], axis=axis))

    def split(self, npartitions):
        pieces = {k: np.array_split(getattr(self, k), npartitions) for k in self._names()}
        return [self._like(**{k: pieces[k][i] for k in pieces}) for i in range(npartitions)]

    def copy(self):
        return _copy.deepcopy(self)

    # -- checkpoint hooks (reference dataset names: ``configs`` and, periodic, ``wrap``) ---------------
    def arrays(self):
        return {k: getattr(self, k) for k in self._names()}

    def load_arrays(self, stored):
        """In-place, keeping dtype and identity of the arrays (``coord.py:107-112``); the walker count
        follows the stored one."""
        for k in self._names():
            data = np.asarray(stored[k])
            if data.shape == getattr(self, k).shape:
                getattr(self, k)[...] = data
            else:
                setattr(self, k, np.array(data, dtype=float))

    def initialize_hdf(self, hdf):
        for k, a in self.arrays().items():
            hdf.create_dataset(k, a.shape, chunks=True, maxshape=(None,) + a.shape[1:])

    def to_hdf(self, hdf):
        for k, a in self.arrays().items():
            hdf[k].resize(a.shape)
            hdf[k][...] = a

    def load_hdf(self, hdf):
        self.load_arrays({k: hdf[k][()] for k in self._names()})


def OpenConfigs(configs, dist=None):
    return Walkers(configs, dist=dist)


def PeriodicConfigs(configs, lattice_vectors, wrap=None, dist=None):
    return Walkers(configs, lattice_vectors, wrap, dist)


OpenElectron = PeriodicElectron = ElectronView
