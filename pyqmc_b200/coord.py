"""Host walker containers (open boundary conditions).

Mirror of ``OpenConfigs`` / ``OpenElectron`` (``pyqmc/configurations/coord.py:21-88``):
``configs (N, nelec, 3)`` C-contiguous float64 on the host -- the drivers (``mc.vmc``) mutate
it in place, the device keeps its own copy.  The reference objects are accepted everywhere
these are (only ``.configs`` is read).
"""
import copy

import numpy as np


class OpenElectron:
    def __init__(self, epos, dist=None):
        self.configs = epos
        self.dist = dist

    def mask(self, mask):
        return OpenElectron(self.configs[mask], self.dist)


class OpenConfigs:
    def __init__(self, configs, dist=None):
        self.configs = configs
        self.dist = dist

    def electron(self, e):
        return OpenElectron(self.configs[:, e], self.dist)

    def select_electrons(self, es):
        return OpenConfigs(self.configs[:, es], self.dist)

    def mask(self, mask):
        return OpenConfigs(self.configs[mask], self.dist)

    def make_irreducible(self, e, vec, mask=True):
        return OpenElectron(vec, self.dist)

    def move(self, e, new, accept):
        self.configs[accept, e, :] = new.configs[accept, :]

    def resample(self, newinds):
        self.configs = self.configs[newinds]

    def split(self, npartitions):
        return [OpenConfigs(c, self.dist) for c in np.array_split(self.configs, npartitions)]

    def join(self, configslist, axis=0):
        self.configs = np.concatenate([c.configs for c in configslist], axis=axis)

    def copy(self):
        return copy.deepcopy(self)

    def reshape(self, shape):
        self.configs = self.configs.reshape(shape)


class PeriodicElectron:
    """``PeriodicElectron`` (coord.py:115-134): trial positions with their wrap vectors."""

    def __init__(self, epos, lattice_vectors, dist=None, wrap=None):
        self.configs = epos
        self.lvec = lattice_vectors
        self.wrap = wrap if wrap is not None else np.zeros_like(epos)
        self.dist = dist

    def mask(self, mask):
        return PeriodicElectron(self.configs[mask], self.lvec, self.dist, wrap=self.wrap[mask])


class PeriodicConfigs:
    """``PeriodicConfigs`` (coord.py:137-252): walkers wrapped into the simulation cell, with the
    integer wrap vectors (in units of the lattice vectors) accumulated since construction."""

    def __init__(self, configs, lattice_vectors, wrap=None, dist=None):
        from .pbc import enforce_pbc

        configs, wrap_ = enforce_pbc(lattice_vectors, configs)
        self.configs = configs
        self.wrap = wrap_
        if wrap is not None:
            self.wrap += wrap
        self.lvecs = lattice_vectors
        self.dist = dist

    def electron(self, e):
        return PeriodicElectron(self.configs[:, e], self.lvecs, self.dist, wrap=self.wrap[:, e])

    def select_electrons(self, es):
        return PeriodicConfigs(self.configs[:, es], self.lvecs, dist=self.dist, wrap=self.wrap[:, es])

    def mask(self, mask):
        return PeriodicConfigs(self.configs[mask], self.lvecs, wrap=self.wrap[mask], dist=self.dist)

    def make_irreducible(self, e, vec, mask=None):
        from .pbc import enforce_pbc

        if mask is None:
            mask = np.ones(vec.shape[0:-1], dtype=bool)
        epos_, wrap_ = enforce_pbc(self.lvecs, vec[mask])
        epos = vec.copy()
        epos[mask] = epos_
        wrap = self.wrap[:, e, :].copy()
        if len(vec.shape) == 3:
            wrap = np.repeat(self.wrap[:, e][:, np.newaxis], vec.shape[1], axis=1)
        wrap[mask] += wrap_
        return PeriodicElectron(epos, self.lvecs, wrap=wrap, dist=self.dist)

    def move(self, e, new, accept):
        self.configs[accept, e, :] = new.configs[accept, :]
        self.wrap[accept, e, :] = new.wrap[accept, :]

    def resample(self, newinds):
        self.configs = self.configs[newinds]
        self.wrap = self.wrap[newinds]

    def split(self, npartitions):
        clist = np.array_split(self.configs, npartitions)
        wlist = np.array_split(self.wrap, npartitions)
        return [PeriodicConfigs(c, self.lvecs, w, dist=self.dist) for c, w in zip(clist, wlist)]

    def join(self, configslist, axis=0):
        self.configs = np.concatenate([c.configs for c in configslist], axis=axis)
        self.wrap = np.concatenate([c.wrap for c in configslist], axis=axis)

    def copy(self):
        return copy.deepcopy(self)

    def reshape(self, shape):
        self.configs = self.configs.reshape(shape)
        self.wrap = self.wrap.reshape(shape)
