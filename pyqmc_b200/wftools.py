"""Factories for the device wave functions -- the selection hook of the drop-in boundary.

The reference chooses its backend in ``pyqmc/wftools.py`` (``generate_slater`` 27-61, ``generate_jastrow``
99-152, ``generate_jastrow3`` 155-162, ``generate_wf`` 195-241; the ``jax=`` switch).  These factories
take the same arguments, apply the same defaults (basis sizes, cutoff radius, cusp constants, which
parameters are optimisable) and return ``(wf, to_opt)`` built from ``pyqmc_b200`` device objects, so
``recipes`` / ``linemin`` consume the result unchanged.  INTEGRATION.md shows the three-line stub that
makes ``pyqmc.wftools`` itself dispatch here.
"""
import numpy as np

from .func3d import CutoffCuspFunction, PolyPadeFunction
from .wf import JastrowSpin, MultiplyWF, Slater, ThreeBodyJastrow

CUSP_GAMMA = 24  # wftools.py:79-80
OPEN_RCUT = 7.5  # wftools.py:84-85
BETA_STEP = 1.6  # QWalk's geometric ladder of Pade shape parameters, wftools.py:64-73
ANTIPARALLEL_CUSP, PARALLEL_CUSP = -0.50, -0.25  # bcoeff[0] = [uu, ud, dd], wftools.py:145


def _mask(shape, value=True):
    return np.full(shape, value, dtype=bool)


def generate_slater(mol, mf, optimize_determinants=False, optimize_orbitals=False, optimize_zeros=True,
                    epsilon=1e-8, jax=False, **slater_kws):
    """Slater determinant(s) on the device.  ``to_opt``: determinant coefficients are optimised only on
    request and never the largest one (it fixes the normalisation); orbital coefficients on request,
    optionally leaving numerical zeros (< epsilon) frozen."""
    wf = Slater(mol, mf, **slater_kws)
    ci = wf.parameters["det_coeff"]
    to_opt = {"det_coeff": _mask(ci.shape, bool(optimize_determinants))}
    if optimize_determinants:
        to_opt["det_coeff"][np.argmax(np.abs(ci))] = False
    if optimize_orbitals:
        for spin in ("mo_coeff_alpha", "mo_coeff_beta"):
            c = wf.parameters[spin]
            to_opt[spin] = _mask(c.shape) if optimize_zeros else np.abs(c) >= epsilon
    return wf, to_opt


def expand_beta_qwalk(beta0, n):
    """Shape parameters ``beta_0, (beta_0 + 1.00001) e^{1.6 i} - 1`` of the n Pade functions.  Evaluated
    term by term in the reference's scalar arithmetic: these numbers parametrise every Jastrow value."""
    log_lead = np.log(beta0 + 1.00001)
    ladder = [beta0] + [np.exp(log_lead + BETA_STEP * rung) - 1 for rung in range(1, n)]
    return np.array(ladder[:n], dtype=float)


def default_jastrow_basis(mol, ion_cusp=False, na=4, nb=3, rcut=None, cusp_gamma=None, beta_a=0.2, beta_b=0.5):
    """(a_basis, b_basis): optional electron-ion cusp function + ``na`` Pade functions; the
    electron-electron cusp function + ``nb`` Pade functions.  ``rcut`` defaults to 7.5 bohr for
    molecules and to the inscribed radius of the simulation cell for solids."""
    gamma = CUSP_GAMMA if cusp_gamma is None else cusp_gamma
    if rcut is None and hasattr(mol, "a"):
        rcut = np.amin(np.pi / np.linalg.norm(mol.reciprocal_vectors(), axis=1))
    elif rcut is None:
        rcut = OPEN_RCUT

    def pade(beta0, n):
        return [PolyPadeFunction(beta=b, rcut=rcut) for b in expand_beta_qwalk(beta0, n)]

    a_basis = ([CutoffCuspFunction(gamma=gamma, rcut=rcut)] if ion_cusp else []) + pade(beta_a, na)
    b_basis = [CutoffCuspFunction(gamma=gamma, rcut=rcut)] + pade(beta_b, nb)
    return a_basis, b_basis


def _cusp_atoms(mol, ion_cusp):
    """Which atoms get the electron-ion cusp term: none (False), all (True), the all-electron ones
    with positive charge (None, the default) or the listed symbols."""
    if isinstance(ion_cusp, bool):
        return [True] * len(mol._atom) if ion_cusp else []
    if ion_cusp is None:
        z = mol.atom_charges()
        symbols = [mol.atom_symbol(i) for i in range(len(mol._atom))]
        return [s for s, zi in zip(symbols, z) if s not in mol._ecp and zi > 0]
    if not isinstance(ion_cusp, list):
        raise TypeError("ion_cusp must be True, False, None or a list of atom symbols")
    return ion_cusp


def generate_jastrow(mol, ion_cusp=None, na=4, nb=3, rcut=None, cusp_gamma=None, beta_a=0.2, beta_b=0.5,
                     jax=False):
    """One- and two-body Jastrow factor with QWalk's default basis; cusp rows fixed, the rest free."""
    cusped = _cusp_atoms(mol, ion_cusp)
    a_basis, b_basis = default_jastrow_basis(mol, len(cusped) > 0, na, nb, rcut, cusp_gamma, beta_a, beta_b)
    wf = JastrowSpin(mol, a_basis=a_basis, b_basis=b_basis)
    acoeff, bcoeff = wf.parameters["acoeff"], wf.parameters["bcoeff"]
    to_opt = {"acoeff": _mask(acoeff.shape), "bcoeff": _mask(bcoeff.shape)}
    if cusped:
        z = np.array(mol.atom_charges(), dtype=float)
        z[[atom[0] not in cusped for atom in mol._atom]] = 0.0
        acoeff[:, 0, :] = z[:, None]
        to_opt["acoeff"][:, 0] = False
    bcoeff[0, :] = [PARALLEL_CUSP, ANTIPARALLEL_CUSP, PARALLEL_CUSP]
    to_opt["bcoeff"][0, :] = False
    return wf, to_opt  # cusp rows stay fixed


def generate_jastrow3(mol, na=4, nb=3, rcut=None, jax=False):
    """Electron-electron-ion factor: default basis without the cusp function, all coefficients free."""
    if jax:
        raise NotImplementedError("there is no JAX three-body Jastrow factor (nor in the reference)")
    j3 = ThreeBodyJastrow(mol, *default_jastrow_basis(mol, False, na, nb, rcut))
    return j3, {"ccoeff": _mask(j3.parameters["ccoeff"].shape)}


def _shared_supercell(mol):
    """Every factor must see ONE cell object: a plain primitive cell is promoted to its 1x1x1 supercell
    here, once, instead of inside each factor (``pyscftools.py:159``)."""
    if hasattr(mol, "a") and not hasattr(mol, "original_cell"):
        from . import pbc

        return pbc.get_supercell(mol, np.eye(3, dtype=int))
    return mol


def generate_wf(mol, mf, jastrow=generate_jastrow, jastrow_kws=None, slater_kws=None, mc=None, jax=False):
    """Slater x Jastrow factor(s), fused into one device context; ``to_opt`` keys carry the ``wfN``
    prefix of the factor they belong to (``multiplywf.py:18-68``)."""
    builders = jastrow if isinstance(jastrow, list) else [jastrow]
    kws = jastrow_kws if isinstance(jastrow, list) else [jastrow_kws]
    kws = [{} if k is None else k for k in (kws if kws is not None else [None] * len(builders))]
    mol = _shared_supercell(mol)
    built = [generate_slater(mol, mf, mc=mc, **(slater_kws or {}))]
    built += [make(mol, **kw) for make, kw in zip(builders, kws)]
    wf = MultiplyWF(*[factor for factor, _ in built])
    to_opt = {f"wf{n}{name}": flags for n, (_, opt) in enumerate(built, start=1) for name, flags in opt.items()}
    return wf, to_opt
