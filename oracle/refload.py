"""Import the UNMODIFIED reference in an environment without pyscf/h5py (test infrastructure).

The package is taken from the staged copy ``oracle/_ref/pyqmc`` (``oracle/stage_reference.py``: the
reference's own files, byte for byte; it travels to the GPU box) or, failing that, from
``/root/reference``.  Used to (a) generate the committed golden vectors (``tests/golden/make_golden.py``),
(b) drive the device objects with the reference's own ``mc.vmc`` / ``dmc.rundmc`` / ``testwf`` in the
``-m gpu`` tests, and (c) as the CPU arm of ``bench.py``.  Recipe from SURVEY.md section 8c: register
empty stub modules for the pyscf/h5py names the reference imports at module scope, then select the
in-tree numba GTO evaluator (``evaluate_orbitals_with="numba"``).  ``pyqmc_b200`` never imports this.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = _STAGED if os.path.isdir(os.path.join(_STAGED, "pyqmc")) else "/root/reference"

_STUBS = [
    "pyscf",
    "pyscf.pbc",
    "pyscf.pbc.gto",
    "pyscf.pbc.gto.eval_gto",
    "pyscf.pbc.gto.cell",
    "pyscf.pbc.scf",
    "pyscf.pbc.scf.addons",
    "pyscf.mcscf",
    "pyscf.fci",
    "pyscf.hci",
    "pyscf.lib",
    "pyscf.scf",
    "pyscf.gto",
    "h5py",
]


def _estimate_rcut(cell, precision=None):
    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from pyqmc_b200 import pbc

    return pbc.estimate_rcut(cell, precision)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyqmc"))


def load():
    """Returns the imported ``pyqmc.api`` module of the reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    for name in _STUBS:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, m)
    sys.modules["h5py"].File = object
    # periodic path: the two pyscf functions that feed TABLES into the reference's numba evaluator
    # (pyqmc/wf/orbitals.py:164-166,268) are replaced by the stand-ins of pyqmc_b200.pbc; the
    # mean-field objects used here are already in k-point UHF layout
    sys.modules["pyscf.pbc.scf.addons"].convert_to_khf = lambda mf: mf
    sys.modules["pyscf.pbc.gto.cell"].estimate_rcut = _estimate_rcut
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pyqmc.api as pyq  # noqa: E402

    return pyq


class GuardedEvalGto:
    """Empty-input guard around the reference's numba ``eval_gto`` (picklable: the reference's
    parallel drivers ship the wave function to worker processes)."""

    NCOMP = {"GTOval_sph": None, "GTOval_sph_deriv1": 4, "GTOval_sph_deriv2": 5}

    def __init__(self, inner, nk, nao):
        self.inner, self.nk, self.nao = inner, tuple(nk), nao

    def __call__(self, eval_str, coords):
        import numpy as np

        if len(coords) == 0:
            nc = self.NCOMP[eval_str.replace("PBC", "")]
            return np.zeros(self.nk + (0, self.nao)) if nc is None else np.zeros(self.nk + (nc, 0, self.nao))
        return self.inner(eval_str, coords)


def build_reference_wf(mol, mf, jastrow=True, determinants=None, seed=0, na=4, nb=3,
                       three_body=False, coeff_scale=0.1):
    """Reference Slater(numba) x JastrowSpin [x ThreeBodyJastrow] with seeded coefficients."""
    import numpy as np

    load()
    import pyqmc.wf.slater
    import pyqmc.wf.multiplywf
    import pyqmc.wftools

    slater = pyqmc.wf.slater.Slater(
        mol, mf, determinants=determinants, evaluate_orbitals_with="numba"
    )
    # The numba evaluator crashes on an empty point set (all-False mask in dmc.py:175 ->
    # slater.py:279 -> gto.py:494; SURVEY.md 8c caveat 2): harness-side guard, as pyscf tolerates it.
    nao = slater.parameters["mo_coeff_alpha"].shape[0]
    nk = (len(slater.orbitals._kpts),) if hasattr(mol, "a") else ()
    slater.orbitals.eval_gto = GuardedEvalGto(slater.orbitals.eval_gto, nk, nao)
    if not jastrow:
        return slater
    jast, _ = pyqmc.wftools.generate_jastrow(mol, na=na, nb=nb)
    rng = np.random.RandomState(seed)
    ac = jast.parameters["acoeff"]
    bc = jast.parameters["bcoeff"]
    has_cusp = len(jast.a_basis) > na
    a0 = 1 if has_cusp else 0
    ac[:, a0:, :] = coeff_scale * rng.randn(*ac[:, a0:, :].shape)
    bc[1:, :] = coeff_scale * rng.randn(*bc[1:, :].shape)
    factors = [slater, jast]
    if three_body:
        j3, _ = pyqmc.wftools.generate_jastrow3(mol, na=na, nb=nb)
        j3.parameters["ccoeff"][...] = 0.2 * coeff_scale * rng.randn(*j3.parameters["ccoeff"].shape)
        factors.append(j3)
    return pyqmc.wf.multiplywf.MultiplyWF(*factors)
