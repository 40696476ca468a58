"""Flat namespace mirroring ``pyqmc.api`` (``pyqmc/api.py``) for the part of PyQMC this package
accelerates, so ``import pyqmc_b200.api as pyq`` reads like the reference:

    wf, to_opt = pyq.generate_wf(mol, mf)
    configs = pyq.initial_guess(mol, nconfig)
    df, configs = pyq.vmc(wf, configs, accumulators={"energy": pyq.EnergyAccumulator(mol)})
    df, configs, weights = pyq.rundmc(wf, configs, accumulators={"energy": pyq.EnergyAccumulator(mol)})
    pgrad = pyq.gradient_generator(mol, wf, to_opt)

Recipes, line minimisation, HDF5 I/O and the density-matrix accumulators are outside the accelerated
path: use the reference's own drivers on these objects for those.
"""
from .accumulators import EnergyAccumulator  # noqa: F401
from .obdm import OBDMAccumulator  # noqa: F401
from .tbdm import TBDMAccumulator  # noqa: F401
from .coord import OpenConfigs, PeriodicConfigs  # noqa: F401
from .dmc import rundmc  # noqa: F401
from .mc import initial_guess, vmc  # noqa: F401
from .pbc import get_supercell  # noqa: F401
from .sr import LinearTransform, PGradTransform, StochasticReconfiguration, gradient_generator  # noqa: F401
from .wf import JastrowSpin, MultiplyWF, Slater, ThreeBodyJastrow  # noqa: F401
from .wftools import generate_jastrow, generate_jastrow3, generate_slater, generate_wf  # noqa: F401
