"""pyqmc_b200 -- B200-native (sm_100a) backend for PyQMC's walker-batched trial-wave-function
hot path: Slater determinants with Sherman-Morrison updates, GTO orbitals, Jastrow factors and
the local-energy accumulator, behind the reference's ``pyqmc.wf`` object protocol.

Importing the package does not touch the GPU; creating a wave function's device context does,
and fails loudly without the CUDA library or a CUDA device (there is no CPU fallback).
"""
from .coord import ElectronView, OpenConfigs, OpenElectron, PeriodicConfigs, PeriodicElectron, Walkers  # noqa: F401
from .func3d import CutoffCuspFunction, PolyPadeFunction  # noqa: F401
from .wf import JastrowSpin, MultiplyWF, Slater, ThreeBodyJastrow  # noqa: F401
from .wftools import generate_jastrow, generate_jastrow3, generate_slater, generate_wf  # noqa: F401
from .accumulators import EnergyAccumulator  # noqa: F401
from .obdm import OBDMAccumulator, normalize_obdm  # noqa: F401
from .tbdm import TBDMAccumulator, normalize_tbdm  # noqa: F401
from .mc import initial_guess, vmc  # noqa: F401
from .dmc import branch, dmc_propagate, rundmc  # noqa: F401
from .sr import LinearTransform, ParameterMap, PGradTransform, StochasticReconfiguration, gradient_generator  # noqa: F401

__all__ = [
    "Walkers", "ElectronView", "OpenConfigs", "OpenElectron", "PeriodicConfigs", "PeriodicElectron", "CutoffCuspFunction",
    "PolyPadeFunction", "JastrowSpin", "MultiplyWF", "Slater", "ThreeBodyJastrow", "generate_jastrow", "generate_jastrow3",
    "generate_slater", "generate_wf", "EnergyAccumulator", "initial_guess", "vmc", "rundmc", "dmc_propagate", "branch",
    "OBDMAccumulator", "normalize_obdm", "TBDMAccumulator", "normalize_tbdm", "ParameterMap", "LinearTransform", "PGradTransform", "StochasticReconfiguration", "gradient_generator",
]
