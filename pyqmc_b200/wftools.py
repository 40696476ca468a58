"""Factory functions with the reference's names and defaults (``pyqmc/wftools.py``).

``generate_slater`` 27-61, ``expand_beta_qwalk`` 64-73, ``default_jastrow_basis`` 76-96,
``generate_jastrow`` 99-152, ``generate_wf`` 195-241 -- building B200 device objects.
"""
import numpy as np

from . import func3d
from .wf import JastrowSpin, MultiplyWF, Slater, ThreeBodyJastrow


def generate_slater(mol, mf, optimize_determinants=False, optimize_orbitals=False, optimize_zeros=True,
                    epsilon=1e-8, **kwargs):
    kwargs.pop("jax", None)
    wf = Slater(mol, mf, **kwargs)
    to_opt = {"det_coeff": np.zeros_like(wf.parameters["det_coeff"], dtype=bool)}
    if optimize_determinants:
        to_opt["det_coeff"] = np.ones_like(wf.parameters["det_coeff"], dtype=bool)
        to_opt["det_coeff"][np.argmax(np.abs(wf.parameters["det_coeff"]))] = False
    if optimize_orbitals:
        for k in ["mo_coeff_alpha", "mo_coeff_beta"]:
            to_opt[k] = np.ones(wf.parameters[k].shape, dtype=bool)
            if not optimize_zeros:
                to_opt[k][np.abs(wf.parameters[k]) < epsilon] = False
    return wf, to_opt


def expand_beta_qwalk(beta0, n):
    if n == 0:
        return np.zeros(0)
    beta = np.zeros(n)
    beta[0] = beta0
    beta1 = np.log(beta0 + 1.00001)
    for i in range(1, n):
        beta[i] = np.exp(beta1 + 1.6 * i) - 1
    return beta


def default_jastrow_basis(mol, ion_cusp=False, na=4, nb=3, rcut=None, cusp_gamma=None, beta_a=0.2, beta_b=0.5):
    if cusp_gamma is None:
        cusp_gamma = 24
    if rcut is None:
        if hasattr(mol, "a"):  # inscribed radius of the simulation cell (wftools.py:82-83)
            rcut = np.amin(np.pi / np.linalg.norm(mol.reciprocal_vectors(), axis=1))
        else:
            rcut = 7.5
    abasis = [func3d.CutoffCuspFunction(gamma=cusp_gamma, rcut=rcut)] if ion_cusp else []
    abasis += [func3d.PolyPadeFunction(beta=b, rcut=rcut) for b in expand_beta_qwalk(beta_a, na)]
    bbasis = [func3d.CutoffCuspFunction(gamma=cusp_gamma, rcut=rcut)]
    bbasis += [func3d.PolyPadeFunction(beta=b, rcut=rcut) for b in expand_beta_qwalk(beta_b, nb)]
    return abasis, bbasis


def generate_jastrow(mol, ion_cusp=None, na=4, nb=3, rcut=None, cusp_gamma=None, beta_a=0.2, beta_b=0.5,
                     jax=False):
    if ion_cusp is False:
        ion_cusp = []
    elif ion_cusp is True:
        ion_cusp = [True] * len(mol._atom)
    elif ion_cusp is None:
        charges = mol.atom_charges()
        ion_cusp = [mol.atom_symbol(i) for i in range(len(mol._atom))
                    if mol.atom_symbol(i) not in mol._ecp.keys() and charges[i] > 0]
    else:
        assert isinstance(ion_cusp, list)
    abasis, bbasis = default_jastrow_basis(mol, len(ion_cusp) > 0, na, nb, rcut, cusp_gamma, beta_a, beta_b)
    jastrow = JastrowSpin(mol, a_basis=abasis, b_basis=bbasis)
    if len(ion_cusp) > 0:
        coefs = np.array(mol.atom_charges(), dtype=float)
        coefs[[atom[0] not in ion_cusp for atom in mol._atom]] = 0.0
        jastrow.parameters["acoeff"][:, 0, :] = coefs[:, None]
    jastrow.parameters["bcoeff"][0, [0, 1, 2]] = np.array([-0.25, -0.50, -0.25])
    to_opt = {"acoeff": np.ones(jastrow.parameters["acoeff"].shape).astype(bool)}
    if len(ion_cusp) > 0:
        to_opt["acoeff"][:, 0, :] = False
    to_opt["bcoeff"] = np.ones(jastrow.parameters["bcoeff"].shape).astype(bool)
    to_opt["bcoeff"][0, [0, 1, 2]] = False
    return jastrow, to_opt


def generate_jastrow3(mol, na=4, nb=3, rcut=None, jax=False):
    """wftools.py:155-162: default basis without the electron-ion cusp, zero coefficients."""
    if jax is True:
        raise NotImplementedError("JAX 3-body Jastrow not yet implemented")
    abasis, bbasis = default_jastrow_basis(mol, False, na, nb, rcut)
    wf = ThreeBodyJastrow(mol, abasis, bbasis)
    to_opt = {"ccoeff": np.ones(wf.parameters["ccoeff"].shape).astype(bool)}
    return wf, to_opt


def generate_wf(mol, mf, jastrow=generate_jastrow, jastrow_kws=None, slater_kws=None, mc=None, jax=False):
    jastrow_kws = {} if jastrow_kws is None else jastrow_kws
    slater_kws = {} if slater_kws is None else slater_kws
    if not isinstance(jastrow, list):
        jastrow, jastrow_kws = [jastrow], [jastrow_kws]
    wf1, to_opt1 = generate_slater(mol, mf, mc=mc, **slater_kws)
    pack = [jast(mol, **kw) for jast, kw in zip(jastrow, jastrow_kws)]
    wf = MultiplyWF(wf1, *[p[0] for p in pack])
    to_opt = {"wf1" + k: v for k, v in to_opt1.items()}
    for i, (_, to_opt2) in enumerate(pack):
        to_opt.update({f"wf{i + 2}" + k: v for k, v in to_opt2.items()})
    return wf, to_opt
