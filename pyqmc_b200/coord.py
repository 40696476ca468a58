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
