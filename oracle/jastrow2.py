"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference one-/two-body Jastrow factor.

Follows ``pyqmc/wf/jastrowspin.py`` (state arrays and update rules: recompute 56-109,
updateinternals 111-137, _a_update 139-159, _b_update 161-191, _update_b_partial 221-249,
gradient 258-294, gradient_value 296-340, gradient_laplacian 342-385, testvalue 387-419,
testvalue_many 421-455, pgradient 457-464) and the radial functions of ``pyqmc/wf/func3d.py``
(PolyPade 25-49; CutoffCusp 112-181; the ``r < rcut`` selection of CutoffFunc3dEvaluator
299-333).  Default basis construction mirrors ``pyqmc/wftools.py:64-96,99-152``.

Open boundary conditions only (distance convention ``dist_i(a, b) = b - a``,
``pyqmc/configurations/distance.py:25-32``).
"""
import numpy as np


# --- radial functions: each returns (value, g, lap) with grad = g * rvec ---------------
def polypade(r, beta, rcut, want):
    z1 = r / rcut - 1.0
    z12 = z1 * z1
    p = (3.0 * z12 + 4.0 * z1) * z12 + 1.0
    obp = 1.0 / (1.0 + beta * p)
    val = (1.0 - p) * obp
    if want == 0:
        return val, None, None
    g = -(1.0 + beta) * 12.0 / rcut**2 * obp * obp * z12
    if want == 1:
        return val, g, None
    with np.errstate(divide="ignore", invalid="ignore"):
        lap = g * (5.0 + 2.0 / z1 - 24.0 * beta * (z1 + 1.0) ** 2 * z12 * obp)
    return val, g, lap


def cutoffcusp(r, gamma, rcut, want):
    y = r / rcut
    y1 = y - 1.0
    a = y1 * y1
    b = (a * y1 + 1.0) / 3.0
    ogb = 1.0 / (1.0 + gamma * b)
    val = (-b * ogb + 1.0 / (3.0 + gamma)) * rcut
    if want == 0:
        return val, None, None
    with np.errstate(divide="ignore", invalid="ignore"):
        c = ogb * ogb / r
        g = -a * c
        if want == 1:
            return val, g, None
        lap = -c * 2.0 * ((y1 - a * a * gamma * ogb) * y + a)
    return val, g, lap


class RadialBasis:
    """List of (kind, parameter) sharing one cutoff; everything is exactly 0 for r >= rcut."""

    def __init__(self, funcs, rcut):
        self.funcs = list(funcs)  # [("cusp", gamma) | ("pade", beta)]
        self.rcut = float(rcut)

    def __len__(self):
        return len(self.funcs)

    def eval(self, r, want):
        """r (...,) -> val (..., nb), g (..., nb), lap (..., nb) (g/lap None when not wanted)."""
        inside = r < self.rcut
        nb = len(self.funcs)
        val = np.zeros(r.shape + (nb,))
        g = np.zeros(r.shape + (nb,)) if want >= 1 else None
        lap = np.zeros(r.shape + (nb,)) if want >= 2 else None
        rs = np.where(inside, r, 0.5 * self.rcut)
        for k, (kind, par) in enumerate(self.funcs):
            f = cutoffcusp if kind == "cusp" else polypade
            v, gg, ll = f(rs, par, self.rcut, want)
            val[..., k] = np.where(inside, v, 0.0)
            if want >= 1:
                g[..., k] = np.where(inside, gg, 0.0)
            if want >= 2:
                lap[..., k] = np.where(inside, ll, 0.0)
        return val, g, lap


def expand_beta(beta0, n):
    """wftools.py:64-73."""
    beta = np.zeros(n)
    if n == 0:
        return beta
    beta[0] = beta0
    b1 = np.log(beta0 + 1.00001)
    for i in range(1, n):
        beta[i] = np.exp(b1 + 1.6 * i) - 1.0
    return beta


def default_basis(mol, na=4, nb=3, rcut=7.5, gamma=24.0, beta_a=0.2, beta_b=0.5):
    """wftools.py:76-96 + the ion_cusp=None rule of 118-126.  Returns (a_funcs, b_funcs, cusp_atoms)."""
    charges = mol.atom_charges()
    cusp_atoms = [
        i for i in range(len(mol._atom))
        if mol.atom_symbol(i) not in mol._ecp.keys() and charges[i] > 0
    ]
    a_funcs = [("cusp", gamma)] if len(cusp_atoms) > 0 else []
    a_funcs += [("pade", b) for b in expand_beta(beta_a, na)]
    b_funcs = [("cusp", gamma)] + [("pade", b) for b in expand_beta(beta_b, nb)]
    return a_funcs, b_funcs, cusp_atoms


class JastrowOracle:
    def __init__(self, mol, a_funcs, b_funcs, rcut=7.5):
        self.a_basis = RadialBasis(a_funcs, rcut)
        self.b_basis = RadialBasis(b_funcs, rcut)
        self._mol = mol
        self._nup = int(mol.nelec[0])
        self._ne = int(np.sum(mol.nelec))
        self.atoms = np.asarray(mol.atom_coords(), dtype=float)
        self.parameters = {
            "bcoeff": np.zeros((len(b_funcs), 3)),
            "acoeff": np.zeros((len(self.atoms), len(a_funcs), 2)),
        }
        self.dtype = float
        # periodic systems: every displacement goes through the minimal-image convention of the
        # walkers' ``dist`` object (jastrowspin.py uses ``configs.dist`` throughout)
        self._minimal = None
        if hasattr(mol, "a"):
            from .pbc import MinimalImage

            self._minimal = MinimalImage(mol.lattice_vectors())

    def _mi(self, d):
        return d if self._minimal is None else self._minimal(d)

    @classmethod
    def default(cls, mol, na=4, nb=3, rcut=None):
        """Same parameter defaults as wftools.generate_jastrow (wftools.py:99-152)."""
        if rcut is None:  # wftools.py:81-85
            rcut = 7.5
            if hasattr(mol, "a"):
                rcut = np.amin(np.pi / np.linalg.norm(mol.reciprocal_vectors(), axis=1))
        a_funcs, b_funcs, cusp_atoms = default_basis(mol, na, nb, rcut)
        j = cls(mol, a_funcs, b_funcs, rcut)
        if len(cusp_atoms) > 0:
            coefs = mol.atom_charges().copy()
            for i in range(len(coefs)):
                if i not in cusp_atoms:
                    coefs[i] = 0.0
            j.parameters["acoeff"][:, 0, :] = coefs[:, None]
        j.parameters["bcoeff"][0, :] = [-0.25, -0.50, -0.25]
        return j

    # --- helpers -------------------------------------------------------------------------
    def _edown(self, e):
        return int(e >= self._nup)

    def _others(self, e):
        return np.arange(self._ne) != e

    def _a_terms(self, pos, want):
        """pos (..., 3) -> displacement to atoms (..., I, 3) and basis values."""
        d = self._mi(pos[..., None, :] - self.atoms)
        r = np.linalg.norm(d, axis=-1)
        return (d,) + self.a_basis.eval(r, want)

    def _b_terms(self, e, pos, current, want):
        """pos (M, [aip,] 3); current (M, ne, 3) -> terms against the other electrons."""
        oth = current[:, self._others(e)]  # (M, ne-1, 3)
        if pos.ndim == 3:
            d = pos[:, :, None, :] - oth[:, None, :, :]
        else:
            d = pos[:, None, :] - oth
        d = self._mi(d)
        r = np.linalg.norm(d, axis=-1)
        return (d,) + self.b_basis.eval(r, want)

    # --- state ---------------------------------------------------------------------------
    def recompute(self, configs):
        c = np.array(configs.configs, dtype=float)
        self._cur = c
        N, ne, _ = c.shape
        na, nb = len(self.a_basis), len(self.b_basis)
        I = len(self.atoms)
        self._a_partial = np.zeros((ne, N, I, na))
        self._b_partial = np.zeros((ne, N, nb, 2))
        for e in range(ne):
            _, av, _, _ = self._a_terms(c[:, e], 0)
            self._a_partial[e] = av
            _, bv, _, _ = self._b_terms(e, c[:, e], c, 0)
            sep = self._nup - int(e < self._nup)
            self._b_partial[e, :, :, 0] = bv[:, :sep].sum(axis=1)
            self._b_partial[e, :, :, 1] = bv[:, sep:].sum(axis=1)
        nup = self._nup
        self._avalues = np.zeros((N, I, na, 2))
        self._avalues[..., 0] = self._a_partial[:nup].sum(axis=0)
        self._avalues[..., 1] = self._a_partial[nup:].sum(axis=0)
        # pair sums in the reference order: (i<j) up-up, all up-down, (i<j) down-down
        self._bvalues = np.zeros((N, nb, 3))
        for i in range(ne):
            for j in range(i + 1, ne):
                r = np.linalg.norm(self._mi(c[:, i] - c[:, j]), axis=-1)
                v, _, _ = self.b_basis.eval(r, 0)
                self._bvalues[:, :, int(i >= nup) + int(j >= nup)] += v
        return self.value()

    def value(self):
        u = np.sum(self._bvalues * self.parameters["bcoeff"], axis=(1, 2))
        u = u + np.einsum("nIks,Iks->n", self._avalues, self.parameters["acoeff"])
        return np.ones(len(u)), u

    def _partials_at(self, e, pos, current):
        """New a/b partial sums for electron e at pos (M,[aip,]3); also per-partner b values."""
        _, av, _, _ = self._a_terms(pos, 0)
        _, bv, _, _ = self._b_terms(e, pos, current, 0)
        sep = self._nup - int(e < self._nup)
        bp = np.stack([bv[..., :sep, :].sum(axis=-2), bv[..., sep:, :].sum(axis=-2)], axis=-1)
        return av, bp, bv

    def updateinternals(self, e, epos, configs, mask=None, saved_values=None):
        N = self._cur.shape[0]
        mask = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        s = self._edown(e)
        if saved_values is None:
            av, bp, bv = self._partials_at(e, epos.configs[mask], self._cur[mask])
        else:
            av, bp, bv = [x[mask] for x in saved_values]
        self._avalues[mask, :, :, s] += av - self._a_partial[e][mask]
        self._bvalues[mask, :, s : s + 2] += bp - self._b_partial[e][mask]
        self._a_partial[e][mask] = av
        # patch the partial sums of every other electron: b(new) - b(old)
        oth = self._others(e)
        old = self._mi(self._cur[mask][:, e][:, None, :] - self._cur[mask][:, oth])
        ov, _, _ = self.b_basis.eval(np.linalg.norm(old, axis=-1), 0)
        diff = np.moveaxis(bv - ov, 1, 0)  # (ne-1, Nm, nb)
        idx = np.nonzero(oth)[0]
        for k, i in enumerate(idx):
            self._b_partial[i, mask, :, s] += diff[k]
        self._b_partial[e][mask] = bp
        self._cur[mask, e, :] = epos.configs[mask]

    # --- single-electron queries -------------------------------------------------------------
    def _grad_terms(self, e, pos, want):
        s = self._edown(e)
        sep = self._nup - int(e < self._nup)
        bc = self.parameters["bcoeff"]
        ac = self.parameters["acoeff"][:, :, s]
        da, av, ag, al = self._a_terms(pos, want)
        db, bv, bg, bl = self._b_terms(e, pos, self._cur, want)
        gb = bg[..., None] * db[:, :, None, :]  # (N, ne-1, nb, 3)
        ga = ag[..., None] * da[:, :, None, :]  # (N, I, na, 3)
        grad = np.einsum("b,nbx->xn", bc[:, s], gb[:, :sep].sum(axis=1))
        grad = grad + np.einsum("b,nbx->xn", bc[:, s + 1], gb[:, sep:].sum(axis=1))
        grad = grad + np.einsum("Ik,nIkx->xn", ac, ga)
        return grad, (av, bv, al, bl, sep, s, ac, bc)

    def gradient(self, e, epos):
        return self._grad_terms(e, epos.configs, 1)[0]

    def gradient_value(self, e, epos):
        grad, (av, bv, _, _, sep, s, ac, bc) = self._grad_terms(e, epos.configs, 1)
        bp = np.stack([bv[:, :sep].sum(axis=1), bv[:, sep:].sum(axis=1)], axis=-1)
        da = av - self._a_partial[e]
        dbp = bp - self._b_partial[e]
        u = np.einsum("nIk,Ik->n", da, ac) + np.einsum("nbs,bs->n", dbp, bc[:, s : s + 2])
        return grad, np.exp(u), (av, bp, bv)

    def gradient_laplacian(self, e, epos):
        grad, (_, _, al, bl, sep, s, ac, bc) = self._grad_terms(e, epos.configs, 2)
        lap = np.einsum("Ik,nIk->n", ac, al)
        lap = lap + np.einsum("b,nb->n", bc[:, s], bl[:, :sep].sum(axis=1))
        lap = lap + np.einsum("b,nb->n", bc[:, s + 1], bl[:, sep:].sum(axis=1))
        return grad, lap + np.sum(grad**2, axis=0)

    def testvalue(self, e, epos, mask=None):
        N = self._cur.shape[0]
        mask = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        s = self._edown(e)
        av, bp, bv = self._partials_at(e, epos.configs[mask], self._cur[mask])
        aux = epos.configs.ndim == 3
        apart = self._a_partial[e][mask]
        bpart = self._b_partial[e][mask]
        if aux:
            apart, bpart = apart[:, None], bpart[:, None]
        u = np.einsum("...Ik,Ik->...", av - apart, self.parameters["acoeff"][..., s])
        u = u + np.einsum("...bs,bs->...", bp - bpart, self.parameters["bcoeff"][:, s : s + 2])
        if aux:  # reference saved-value layout has the aip axis first (jastrowspin.py:155,185)
            saved = (np.moveaxis(av, 1, 0), np.moveaxis(bp, 1, 0), np.moveaxis(bv, 1, 0))
        else:
            saved = (av, bp, bv)
        return np.exp(u), saved

    def testvalue_many(self, e, epos, mask=None):
        e = np.asarray(e)
        N = self._cur.shape[0]
        mask = np.ones(N, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        pos = epos.configs[mask]
        cur = self._cur[mask]
        out = np.zeros((pos.shape[0], len(e)))
        _, av, _, _ = self._a_terms(pos, 0)  # (M, I, na)
        d = self._mi(pos[:, None, :] - cur)
        bv, _, _ = self.b_basis.eval(np.linalg.norm(d, axis=-1), 0)  # (M, ne, nb)
        nup = self._nup
        tot = np.stack([bv[:, :nup].sum(axis=1), bv[:, nup:].sum(axis=1)], axis=-1)  # (M, nb, 2)
        for k, el in enumerate(e):
            s = self._edown(el)
            bp = tot.copy()
            bp[:, :, s] -= bv[:, el]
            u = np.einsum("nIk,Ik->n", av - self._a_partial[el][mask], self.parameters["acoeff"][..., s])
            u = u + np.einsum("nbs,bs->n", bp - self._b_partial[el][mask],
                              self.parameters["bcoeff"][:, s : s + 2])
            out[:, k] = np.exp(u)
        return out

    def pgradient(self):
        return {"bcoeff": self._bvalues.copy(), "acoeff": self._avalues.copy()}
