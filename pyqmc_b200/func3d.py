"""Parameter holders for the Jastrow radial basis functions.

Same constructor signature and ``parameters`` dictionaries as the reference classes
(``pyqmc/wf/func3d.py:52-110`` PolyPadeFunction, ``112-210`` CutoffCuspFunction); the arithmetic itself
runs in the CUDA kernels (``csrc/device_common.cuh: radial_func``), which receive ``kind`` and the single
shape parameter (beta or gamma) per function plus the common cutoff radius.
"""

KIND_POLYPADE = 0
KIND_CUSP = 1


class _RadialFunction:
    kind = None
    shape_name = None

    def __init__(self, *args, **kwargs):
        names = (self.shape_name, "rcut")
        given = dict(zip(names, args), **kwargs)
        if set(given) != set(names):
            raise TypeError(f"{type(self).__name__} takes {names[0]} and rcut")
        self.parameters = {k: float(given[k]) for k in names}

    @property
    def shape_parameter(self):
        return self.parameters[self.shape_name]


class PolyPadeFunction(_RadialFunction):
    """``(1 - p(z)) / (1 + beta p(z))``, ``p = 6z^2 - 8z^3 + 3z^4``, ``z = r / rcut``; zero beyond rcut."""

    kind, shape_name = KIND_POLYPADE, "beta"


class CutoffCuspFunction(_RadialFunction):
    """``rcut (-p / (1 + gamma p) + 1 / (3 + gamma))``, ``p = y - y^2 + y^3 / 3``, ``y = r / rcut``: unit slope at
    the origin (the cusp), zero beyond rcut."""

    kind, shape_name = KIND_CUSP, "gamma"
