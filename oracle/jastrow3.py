"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference three-body Jastrow factor.

Follows ``pyqmc/wf/three_body_jastrow.py``:  U = 1/2 sum_i P_i,
    P_i = sum_{j != i} P_ij,   P_ij = sum_{I k l m} C[I,k,l,m,sp(i,j)] a_k(r_iI) a_l(r_jI) b_m(r_ij)
with C = (ccoeff + ccoeff^T(k,l)) / 2 (recompute 66-104), the spin-pair index of electron e (spin s)
with an up / down partner being s / s+1 (single_e_partial 197-262), caches ``a_values (ne, N, I, na)``
and ``P_i (ne, N)`` (updateinternals 149-189), ``testvalue`` 328-347, ``testvalue_many`` 349-375,
``gradient_value`` 454-539, ``gradient_laplacian`` 541-655 (Laplacian of the factor returned as
lap U + |grad U|^2), ``pgradient`` 657-719.  Basis construction: ``wftools.generate_jastrow3``
(wftools.py:155-162) = default_jastrow_basis without the electron-ion cusp.
"""
import numpy as np

from .jastrow2 import RadialBasis, expand_beta


class Jastrow3Oracle:
    def __init__(self, mol, a_funcs, b_funcs, rcut=7.5):
        self.a_basis = RadialBasis(a_funcs, rcut)
        self.b_basis = RadialBasis(b_funcs, rcut)
        self._mol = mol
        self._nup = int(mol.nelec[0])
        self._ne = int(np.sum(mol.nelec))
        self.atoms = np.asarray(mol.atom_coords(), dtype=float)
        na, nb = len(a_funcs), len(b_funcs)
        self.parameters = {"ccoeff": np.zeros((len(self.atoms), na, na, nb, 3))}
        self.dtype = float
        self._minimal = None
        if hasattr(mol, "a"):
            from .pbc import MinimalImage

            self._minimal = MinimalImage(mol.lattice_vectors())

    def _mi(self, d):
        return d if self._minimal is None else self._minimal(d)

    @classmethod
    def default(cls, mol, na=4, nb=3, rcut=None, gamma=24.0, beta_a=0.2, beta_b=0.5):
        if rcut is None:  # wftools.py:81-85
            rcut = 7.5
            if hasattr(mol, "a"):
                rcut = np.amin(np.pi / np.linalg.norm(mol.reciprocal_vectors(), axis=1))
        a_funcs = [("pade", b) for b in expand_beta(beta_a, na)]
        b_funcs = [("cusp", gamma)] + [("pade", b) for b in expand_beta(beta_b, nb)]
        return cls(mol, a_funcs, b_funcs, rcut)

    # ---- helpers ---------------------------------------------------------------------------
    def _sym(self):
        c = self.parameters["ccoeff"]
        return (c + c.swapaxes(1, 2)) / 2

    def _pair_index(self, e, j):
        return int(e >= self._nup) + int(j >= self._nup)

    def _a_at(self, pos, want):
        d = self._mi(pos[..., None, :] - self.atoms)  # (..., I, 3)
        r = np.linalg.norm(d, axis=-1)
        return (d,) + self.a_basis.eval(r, want)

    def _pairs(self, e, pos, cur, avals, want):
        """Per-partner terms of electron e at pos (M,[aip,]3): list of (j, P_ej, extras)."""
        C = self._sym()
        da, av, ag, al = self._a_at(pos, want)  # (M,[aip,]I,na)
        out = []
        for j in range(self._ne):
            if j == e:
                continue
            oth = cur[:, j]
            dv = self._mi(pos - (oth[:, None, :] if pos.ndim == 3 else oth))
            r = np.linalg.norm(dv, axis=-1)
            bv, bg, bl = self.b_basis.eval(r, want)
            aj = avals[j]  # (M, I, na)
            if pos.ndim == 3:
                aj = aj[:, None]
            Csp = C[..., self._pair_index(e, j)]  # (I, k, l, m)
            S0 = np.einsum("Iklm,...Ik,...Il->...Im", Csp, av, aj)
            P = np.einsum("...Im,...m->...", S0, bv)
            extra = None
            if want >= 1:
                S1 = np.einsum("Iklm,...Ik,...Il->...Im", Csp, ag, aj)
                grad = np.einsum("...Im,...m,...Id->...d", S1, bv, da) + \
                    np.einsum("...Im,...m->...", S0, bg)[..., None] * dv
                extra = [grad]
                if want >= 2:
                    S2 = np.einsum("Iklm,...Ik,...Il->...Im", Csp, al, aj)
                    dot = np.einsum("...Id,...d->...I", da, dv)
                    lap = np.einsum("...Im,...m->...", S2, bv) + \
                        2.0 * np.einsum("...Im,...m,...I->...", S1, bg, dot) + \
                        np.einsum("...Im,...m->...", S0, bl)
                    extra.append(lap)
            out.append((j, P, extra))
        return out, av

    # ---- state -------------------------------------------------------------------------------
    def recompute(self, configs):
        c = np.array(configs.configs, dtype=float)
        self._cur = c
        N, ne, _ = c.shape
        self.a_values = np.zeros((ne, N, len(self.atoms), len(self.a_basis)))
        for e in range(ne):
            self.a_values[e] = self._a_at(c[:, e], 0)[1]
        self.P_i = np.zeros((ne, N))
        for e in range(ne):
            pairs, _ = self._pairs(e, c[:, e], c, self.a_values, 0)
            for _, P, _ in pairs:
                self.P_i[e] += P
        self.val = 0.5 * self.P_i.sum(axis=0)
        return self.value()

    def value(self):
        return np.ones(len(self.val)), self.val.copy()

    def updateinternals(self, e, epos, configs, mask=None, saved_values=None):
        N = self._cur.shape[0]
        mask = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        cur = self._cur[mask]
        av = self.a_values[:, mask]
        new, ae = self._pairs(e, epos.configs[mask], cur, av, 0)
        old, _ = self._pairs(e, cur[:, e], cur, av, 0)
        newval = sum(P for _, P, _ in new)
        self.val[mask] += newval - self.P_i[e, mask]
        self.P_i[e, mask] = newval
        for (j, Pn, _), (_, Po, _) in zip(new, old):
            self.P_i[j, mask] += Pn - Po
        self.a_values[e, mask] = ae
        self._cur[mask, e, :] = epos.configs[mask]

    # ---- queries -------------------------------------------------------------------------------
    def testvalue(self, e, epos, mask=None):
        N = self._cur.shape[0]
        mask = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        pairs, _ = self._pairs(e, epos.configs[mask], self._cur[mask], self.a_values[:, mask], 0)
        Pnew = sum(P for _, P, _ in pairs)
        old = self.P_i[e, mask]
        if epos.configs.ndim == 3:
            old = old[:, None]
        return np.exp(Pnew - old), None

    def testvalue_many(self, e, epos, mask=None):
        N = self._cur.shape[0]
        mask = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        out = np.zeros((int(mask.sum()), len(e)))
        for k, el in enumerate(np.asarray(e)):
            out[:, k] = self.testvalue(int(el), epos, mask)[0]
        return out

    def gradient(self, e, epos):
        pairs, _ = self._pairs(e, epos.configs, self._cur, self.a_values, 1)
        return sum(x[0] for _, _, x in pairs).T

    def gradient_value(self, e, epos):
        pairs, _ = self._pairs(e, epos.configs, self._cur, self.a_values, 1)
        grad = sum(x[0] for _, _, x in pairs).T
        Pnew = sum(P for _, P, _ in pairs)
        return grad, np.exp(Pnew - self.P_i[e]), None

    def gradient_laplacian(self, e, epos):
        pairs, _ = self._pairs(e, epos.configs, self._cur, self.a_values, 2)
        grad = sum(x[0] for _, _, x in pairs).T
        lap = sum(x[1] for _, _, x in pairs)
        return grad, lap + np.sum(grad**2, axis=0)

    def pgradient(self):
        c = self._cur
        N, ne, _ = c.shape
        na, nb = len(self.a_basis), len(self.b_basis)
        ders = np.zeros((N, len(self.atoms), na, na, nb, 3))
        for i in range(ne):
            for j in range(i + 1, ne):
                r = np.linalg.norm(self._mi(c[:, i] - c[:, j]), axis=-1)
                bv, _, _ = self.b_basis.eval(r, 0)
                ders[..., self._pair_index(i, j)] += np.einsum(
                    "nIk,nIl,nm->nIklm", self.a_values[i], self.a_values[j], bv)
        ders = ders + ders.swapaxes(2, 3)
        return {"ccoeff": 0.5 * ders}
