"""TEST INFRASTRUCTURE (oracle) -- host walker containers, open boundary conditions.

Restates ``OpenConfigs`` / ``OpenElectron`` of ``pyqmc/configurations/coord.py:21-88``
(only what the hot path touches: electron(e), make_irreducible, move, split, join, copy).
"""
import numpy as np


class Electron:
    def __init__(self, epos):
        self.configs = epos


class Walkers:
    def __init__(self, configs):
        self.configs = configs

    def electron(self, e):
        return Electron(self.configs[:, e])

    def make_irreducible(self, e, vec, mask=True):
        return Electron(vec)

    def move(self, e, new, accept):
        self.configs[accept, e, :] = new.configs[accept, :]

    def split(self, n):
        return [Walkers(c) for c in np.array_split(self.configs, n)]

    def join(self, parts):
        self.configs = np.concatenate([p.configs for p in parts], axis=0)

    def copy(self):
        return Walkers(self.configs.copy())
