"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference's in-tree GTO evaluator.

Follows ``pyqmc/wf/numba/gto.py``:
  * contraction normalisation ``normalize_basis_coeffs`` (gto.py:375-405);
  * table flattening of ``AtomicOrbitalEvaluator.__init__`` (gto.py:435-488): atoms in
    ``mol._atom`` order, shells in ``mol._basis[symbol]`` order, 2l+1 functions per shell;
  * value / gradient / Laplacian kernels ``mol_eval_gto{,_grad,_lap}`` (gto.py:89-254) with
    the radial sums of ``radial_gto{,_grad,_lap}`` (gto.py:257-321):
        chi = S_lm R,  R = sum_p c_p exp(-a_p r^2)
        d_i chi = d_i S R + S d_i R,   d_i R = sum_p -2 a_p x_i c_p exp(-a_p r^2)
        lap chi = S sum_p c_p 2 a_p (2 a_p r^2 - 3) exp(-a_p r^2) + 2 grad S . grad R
  * AO -> MO contraction ``MoleculeOrbitalEvaluator.mos`` (pyqmc/wf/orbitals.py:95-96).

The reference's default evaluator is pyscf's libcgto (``pyqmc/wf/orbitals.py:47``; pyscf
>=2.8,<3 per pyproject.toml:12), which is not vendored and not installable here; the
reference itself pins numba == pyscf only to 3e-5 (tests/unit/test_gto.py:114-160).
At the pyscf boundary: PARITY UNPINNED.  Against the in-tree numba evaluator this file
is pinned by ``tests/test_oracle_vs_reference.py`` and the committed golden vectors.
"""
import numpy as np
from scipy.special import gamma

from . import solid_harmonics as sh


def normalize_shell(l, exps, coefs):
    """gto.py:375-405 : primitive normalisation followed by contraction normalisation."""
    exps = np.asarray(exps, dtype=float)
    coefs = np.asarray(coefs, dtype=float)
    m = l + 1.5
    prim = np.sqrt(2.0 * (2.0 * exps) ** m / gamma(m))
    cs = coefs * prim
    pair = exps[:, None] + exps[None, :]
    overlap = gamma(m) / (2.0 * pair**m)
    norm = cs @ overlap @ cs
    return cs / np.sqrt(norm)


class BasisTable:
    """Flattened shell tables for a molecule (same content as gto.py:435-488)."""

    def __init__(self, mol):
        self.atom_coords = np.asarray(mol.atom_coords(), dtype=float)
        natom = len(self.atom_coords)
        shell_atom, shell_l, offs, exps, coefs, ao_off = [], [], [0], [], [], [0]
        self.max_l = np.zeros(natom, dtype=int)
        for a in range(natom):
            sym = mol.atom_pure_symbol(a)
            for shell in mol._basis[sym]:
                l = int(shell[0])
                prim = np.asarray(shell[1:], dtype=float)
                if prim.shape[1] != 2:
                    raise NotImplementedError("one contraction column per shell (gto.py:441-456)")
                shell_atom.append(a)
                shell_l.append(l)
                exps.extend(prim[:, 0])
                coefs.extend(normalize_shell(l, prim[:, 0], prim[:, 1]))
                offs.append(len(exps))
                ao_off.append(ao_off[-1] + 2 * l + 1)
                self.max_l[a] = max(self.max_l[a], l)
        self.shell_atom = np.asarray(shell_atom)
        self.shell_l = np.asarray(shell_l)
        self.prim_off = np.asarray(offs)
        self.exps = np.asarray(exps)
        self.coefs = np.asarray(coefs)
        self.ao_off = np.asarray(ao_off)
        self.nao = int(ao_off[-1])
        self.nshell = len(shell_l)

    def eval(self, deriv, points):
        """AO values at points (P,3).

        deriv = 0 -> (P, A); 1 -> (4, P, A) [val, dx, dy, dz]; 2 -> (5, P, A) [.., lap].
        """
        points = np.asarray(points, dtype=float).reshape(-1, 3)
        P = points.shape[0]
        ncomp = (1, 4, 5)[deriv]
        out = np.zeros((ncomp, P, self.nao))
        for a, center in enumerate(self.atom_coords):
            rv = points - center
            x, y, z = rv[:, 0], rv[:, 1], rv[:, 2]
            r2 = x * x + y * y + z * z
            if deriv == 0:
                S = sh.evaluate(self.max_l[a], x, y, z)
            else:
                S, dS = sh.evaluate(self.max_l[a], x, y, z, deriv=True)
            for s in np.nonzero(self.shell_atom == a)[0]:
                l = self.shell_l[s]
                al = self.exps[self.prim_off[s] : self.prim_off[s + 1]]
                cf = self.coefs[self.prim_off[s] : self.prim_off[s + 1]]
                g = np.exp(-r2[:, None] * al[None, :]) * cf[None, :]  # (P, nprim)
                R = g.sum(axis=1)
                lo, hi = self.ao_off[s], self.ao_off[s + 1]
                Sl = S[l * l : (l + 1) * (l + 1)].T  # (P, 2l+1)
                out[0, :, lo:hi] = Sl * R[:, None]
                if deriv == 0:
                    continue
                Rp = -(g * (2.0 * al)[None, :]).sum(axis=1)  # dR/dx_i = Rp * x_i
                dSl = dS[:, l * l : (l + 1) * (l + 1)]  # (3, 2l+1, P)
                cross = np.zeros((P, hi - lo))
                for i in range(3):
                    dRi = Rp * rv[:, i]
                    out[1 + i, :, lo:hi] = dSl[i].T * R[:, None] + Sl * dRi[:, None]
                    cross += dSl[i].T * dRi[:, None]
                if deriv == 2:
                    Rl = (g * (2.0 * al * (2.0 * al[None, :] * r2[:, None] - 3.0))).sum(axis=1)
                    out[4, :, lo:hi] = Sl * Rl[:, None] + 2.0 * cross
        return out[0] if deriv == 0 else out
