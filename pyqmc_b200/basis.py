"""Flattened, normalised GTO shell tables for the device (input of ``qmcb_set_basis``).

Content and ordering are those of the reference's in-tree evaluator
(``AtomicOrbitalEvaluator.__init__``, ``pyqmc/wf/numba/gto.py:435-488``): atoms in
``mol._atom`` order, shells in ``mol._basis[symbol]`` order, one contraction column per
shell, coefficients normalised as ``normalize_basis_coeffs`` (gto.py:375-405).
"""
import numpy as np
from scipy.special import gamma


def normalized_coefficients(l, exps, coefs):
    m = l + 1.5
    cs = coefs * np.sqrt(2.0 * (2.0 * exps) ** m / gamma(m))
    pair = exps[:, None] + exps[None, :]
    norm = cs @ (gamma(m) / (2.0 * pair**m)) @ cs
    return cs / np.sqrt(norm)


def shell_tables(mol):
    shell_atom, shell_l, prim_off, exps, coefs = [], [], [0], [], []
    for a in range(len(mol._atom)):
        sym = mol.atom_pure_symbol(a) if hasattr(mol, "atom_pure_symbol") else mol._atom[a][0]
        for shell in mol._basis[sym]:
            l = int(shell[0])
            start = 1
            if not hasattr(shell[1], "__len__"):  # pyscf allows an optional kappa entry
                start = 2
            prim = np.asarray(shell[start:], dtype=float)
            if prim.ndim != 2 or prim.shape[1] != 2:
                raise NotImplementedError("general contractions: one contraction column per shell is required "
                                          "(as in the reference numba evaluator, gto.py:441-456)")
            if l > 5:  # the reference's evaluator dispatches l <= 5 (gto.py:107-118)
                raise NotImplementedError("angular momentum l > 5")
            shell_atom.append(a)
            shell_l.append(l)
            exps.extend(prim[:, 0])
            coefs.extend(normalized_coefficients(l, prim[:, 0], prim[:, 1]))
            prim_off.append(len(exps))
    nao = int(sum(2 * l + 1 for l in shell_l))
    return dict(
        shell_atom=np.asarray(shell_atom, dtype=np.int32),
        shell_l=np.asarray(shell_l, dtype=np.int32),
        prim_off=np.asarray(prim_off, dtype=np.int32),
        exps=np.asarray(exps, dtype=np.float64),
        coefs=np.asarray(coefs, dtype=np.float64),
        nao=nao,
    )
