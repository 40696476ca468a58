"""Parameter holders for the Jastrow radial basis functions.

Same constructor signature and ``parameters`` dictionaries as the reference classes
(``pyqmc/wf/func3d.py:52-110`` PolyPadeFunction, ``112-210`` CutoffCuspFunction); the
arithmetic itself runs in the CUDA kernels (``csrc/device_common.cuh: radial_func``).
"""

KIND_POLYPADE = 0
KIND_CUSP = 1


class PolyPadeFunction:
    kind = KIND_POLYPADE

    def __init__(self, beta, rcut):
        self.parameters = {"beta": float(beta), "rcut": float(rcut)}

    @property
    def shape_parameter(self):
        return float(self.parameters["beta"])


class CutoffCuspFunction:
    kind = KIND_CUSP

    def __init__(self, gamma, rcut):
        self.parameters = {"gamma": float(gamma), "rcut": float(rcut)}

    @property
    def shape_parameter(self):
        return float(self.parameters["gamma"])
