"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference multi-determinant Slater
wave function (open boundary conditions, real orbitals).

Follows ``pyqmc/wf/slater.py``:
  * ``sherman_morrison_ms`` (slater.py:88-94) -> :func:`rank1_row_update`;
  * ``recompute`` (227-260), ``updateinternals`` (262-291), ``value`` (293-299);
  * ``_testrow`` (301-340) / ``_testrowderiv`` (342-380) with the overflow-safe
    determinant sum of ``determinant_tools.compute_value`` (determinant_tools.py:74-88);
  * ``gradient`` (390-401), ``gradient_value`` (403-418), ``gradient_laplacian`` (420-427),
    ``testvalue`` (429-446), ``testvalue_many`` (448-460), ``pgradient`` (462-542);
  * determinant bookkeeping ``create_packed_objects`` (determinant_tools.py:39-71) and MO
    truncation of ``orbital_evaluator_from_pyscf`` (pyscftools.py:176-186).

State layout is the reference's: ``inverse[s] (N, D_s, n_s, n_s)`` indexed [orbital, electron],
``dets[s] (2, N, D_s)`` = [sign, log|det|].
"""
import warnings

import numpy as np

from .gto import BasisTable


def rank1_row_update(e, inv, vec):
    """Replace electron-row ``e`` of every matrix whose inverse is ``inv (..., n, n)`` by ``vec (..., n)``.

    Returns (det ratio (...,), updated inverse).  Same arithmetic as slater.py:88-94.
    """
    t = np.matmul(vec[..., None, :], inv)[..., 0, :]  # t_j = sum_k vec_k inv[k, j]
    ratio = t[..., e]
    col = inv[..., :, e] / ratio[..., None]
    new = inv - col[..., :, None] * t[..., None, :]
    new[..., :, e] = col
    return ratio, new


def pack_determinants(determinants, tol):
    """determinant_tools.py:39-71: unique spin-determinants + map from full list."""
    weights, occ, dmap = [], [[], []], [[], []]
    for w, spin_occ in determinants:
        if abs(w) <= tol:
            continue
        weights.append(w)
        for s in (0, 1):
            o = [int(i) for i in spin_occ[s]]
            if o in occ[s]:
                dmap[s].append(occ[s].index(o))
            else:
                dmap[s].append(len(occ[s]))
                occ[s].append(o)
    weights = np.array(weights)
    return (weights if np.iscomplexobj(weights) else weights.astype(float)), occ, np.array(dmap, dtype=int)


def determinants_from_mf(mf):
    """pyscftools.py:206-219 (single determinant from mo_occ)."""
    mf = mf.to_uhf()
    return [(1.0, [list(np.nonzero(np.asarray(o) > 0.5)[0]) for o in mf.mo_occ])]


class SlaterOracle:
    def __init__(self, mol, mf, determinants=None, tol=None):
        tol = -1 if tol is None else tol
        self._mol = mol
        self._nelec = tuple(mol.nelec)
        mfu = mf.to_uhf()
        if determinants is None:
            determinants = determinants_from_mf(mfu)
        top = [0, 0]
        for _, d in determinants:
            for s in (0, 1):
                if len(d[s]) > 0:
                    top[s] = max(top[s], int(np.max(d[s])) + 1)
        coeff, self._det_occup, self._det_map = pack_determinants(determinants, tol)
        self.parameters = {
            "det_coeff": coeff,
            "mo_coeff_alpha": np.array(mfu.mo_coeff[0][:, : top[0]]),
            "mo_coeff_beta": np.array(mfu.mo_coeff[1][:, : top[1]]),
        }
        self.basis = BasisTable(mol)
        # slater.py:212-216: complex when any parameter is complex; the phase of a determinant then replaces its sign
        self.dtype = complex if any(np.iscomplexobj(v) for v in self.parameters.values()) else float

    # --- orbital evaluation -------------------------------------------------------
    def _mo_coeff(self, s):
        return self.parameters["mo_coeff_alpha" if s == 0 else "mo_coeff_beta"]

    def _mos(self, ao, s):
        return ao @ self._mo_coeff(s)

    def _ao_for_mo(self, ao, s, i):
        return ao

    def _spin(self, e):
        s = int(e >= self._nelec[0])
        return s, e - s * self._nelec[0]

    def _ao(self, deriv, epos, mask=None):
        pts = epos.configs if mask is None else epos.configs[mask]
        shape = pts.shape[:-1]
        ao = self.basis.eval(deriv, pts.reshape(-1, 3))
        nao = self.basis.nao  # explicit: reshape(-1) is ambiguous for an empty selection
        if deriv == 0:
            return ao.reshape(*shape, nao)
        return ao.reshape(ao.shape[0], *shape, nao)

    # --- internal state -----------------------------------------------------------
    def recompute(self, configs):
        N, ne, _ = configs.configs.shape
        self._aovals = self._ao(0, configs)  # (N, ne, A)
        self._dets, self._inverse = [], []
        for s in (0, 1):
            lo = self._nelec[0] * s
            hi = self._nelec[0] + self._nelec[1] * s
            mo = self._mos(self._aovals[:, lo:hi], s)  # (N, n_s, nmo)
            mats = np.swapaxes(mo[:, :, self._det_occup[s]], 1, 2)  # (N, D_s, n_s, n_s)
            assert mats.shape[-1] == mats.shape[-2]
            sign, logdet = np.linalg.slogdet(mats)
            self._dets.append(np.array([sign, logdet]))
            ok = np.isfinite(logdet)
            if np.any(np.abs(sign) < 1e-16):
                warnings.warn("A wave function is zero.")
            inv = np.zeros_like(mats)
            inv[ok] = np.linalg.inv(mats[ok])
            self._inverse.append(inv)
        return self.value()

    def updateinternals(self, e, epos, configs, mask=None, saved_values=None):
        s, eeff = self._spin(e)
        N = epos.configs.shape[0]
        if mask is None:
            mask = np.ones(N, dtype=bool)
        mask = np.asarray(mask, dtype=bool)
        if np.any(np.isinf(self._dets[s][1])):
            warnings.warn("Found a zero in the wave function. Recomputing everything.")
            self.recompute(configs)
            return
        if saved_values is None:
            ao = self._ao(0, epos, mask)
            mo = self._mos(ao, s)
        else:
            ao_all, mo_all = saved_values
            ao, mo = ao_all[mask], mo_all[mask]
        self._aovals[mask, e, :] = ao
        rows = mo[:, self._det_occup[s]]  # (Nm, D_s, n_s)
        ratio, self._inverse[s][mask] = rank1_row_update(eeff, self._inverse[s][mask], rows)
        self._dets[s][0, mask] *= (ratio / np.abs(ratio)) if self.dtype == complex else np.sign(ratio)
        self._dets[s][1, mask] += np.log(np.abs(ratio))

    def _det_weights(self, mask=None):
        """(N[m], D) array of c_D * sign * exp(log - global refs) and its row sums' factors."""
        sel = slice(None) if mask is None else mask
        upref = np.amax(self._dets[0][1]).real
        dnref = np.amax(self._dets[1][1]).real
        m0, m1 = self._det_map
        up, dn = self._dets[0][:, sel], self._dets[1][:, sel]
        amp = up[0][:, m0] * dn[0][:, m1] * np.exp(up[1][:, m0] + dn[1][:, m1] - upref - dnref)
        return amp, upref, dnref

    def value(self):
        amp, upref, dnref = self._det_weights()
        val = amp @ self.parameters["det_coeff"]
        with np.errstate(divide="ignore", invalid="ignore"):
            sign = np.nan_to_num(val / np.abs(val))
            logv = np.nan_to_num(np.log(np.abs(val)) + upref + dnref)
        return sign, logv

    def _combine(self, per_det, s, mask=None):
        """per_det: (C, Nm, X, D_s) single-determinant ratios -> (C, Nm, X) combined ratio."""
        amp, _, _ = self._det_weights(mask)  # (Nm, D)
        w = amp * self.parameters["det_coeff"][None, :]
        num = np.einsum("cnxd,nd->cnx", per_det[..., self._det_map[s]], w)
        den = w.sum(axis=1)
        return num / den[None, :, None]

    def _ratios(self, e, mo, mask=None):
        """mo: (C, Nm, X, nmo) orbital values (C components, X aux points) for electron e."""
        s, eeff = self._spin(e)
        sel = slice(None) if mask is None else mask
        rows = mo[..., self._det_occup[s]]  # (C, Nm, X, D_s, n_s)
        invcol = self._inverse[s][sel][..., eeff]  # (Nm, D_s, n_s)
        per_det = np.einsum("cnxdj,ndj->cnxd", rows, invcol)
        return self._combine(per_det, s, mask)

    # --- single-electron queries ----------------------------------------------------
    def gradient(self, e, epos):
        s, _ = self._spin(e)
        mo = self._mos(self._ao(1, epos), s)  # (4, N, nmo)
        r = self._ratios(e, mo[:, :, None, :])[:, :, 0]
        return r[1:] / r[0]

    def gradient_value(self, e, epos):
        s, _ = self._spin(e)
        ao = self._ao(1, epos)
        mo = self._mos(ao, s)
        r = self._ratios(e, mo[:, :, None, :])[:, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            g = r[1:] / r[0]
        g[~np.isfinite(g)] = 0.0
        v = r[0].copy()
        v[~np.isfinite(v)] = 1.0
        return g, v, (ao[0], mo[0])

    def gradient_laplacian(self, e, epos):
        s, _ = self._spin(e)
        mo = self._mos(self._ao(2, epos), s)  # (5, N, nmo)
        r = self._ratios(e, mo[:, :, None, :])[:, :, 0]
        r = r / r[:1]
        return r[1:4], r[4]

    def testvalue(self, e, epos, mask=None):
        s, _ = self._spin(e)
        if mask is not None:
            mask = np.asarray(mask, dtype=bool)
        ao = self._ao(0, epos, mask)  # (Nm, [aip,] A)
        mo = self._mos(ao, s)
        aux = mo.ndim == 3
        mo4 = mo[None] if aux else mo[None, :, None, :]
        r = self._ratios(e, mo4, mask)[0]
        return (r if aux else r[:, 0]), (ao, mo)

    def testvalue_many(self, e, epos, mask=None):
        e = np.asarray(e)
        spins = (e >= self._nelec[0]).astype(int)
        ao = self._ao(0, epos, mask)  # (Nm, A)
        out = np.zeros((ao.shape[0], len(e)), dtype=self.dtype)  # ao.shape[0] = Nm in both layouts
        for s in (0, 1):
            idx = np.nonzero(spins == s)[0]
            if len(idx) == 0:
                continue
            mo = self._mos(ao, s)
            rows = mo[:, self._det_occup[s]]  # (Nm, D_s, n_s)
            sel = slice(None) if mask is None else mask
            inv = self._inverse[s][sel]  # (Nm, D_s, n_s, n_s)
            eeff = e[idx] - s * self._nelec[0]
            per_det = np.einsum("ndj,ndje->ned", rows, inv[..., eeff])  # (Nm, len(idx), D_s)
            out[:, idx] = self._combine(per_det[None], s, mask)[0]
        return out

    # --- parameter gradient -----------------------------------------------------------
    def pgradient(self):
        sign, logv = self.value()
        coeff = self.parameters["det_coeff"]
        m0, m1 = self._det_map
        nz = sign != 0.0
        N = len(sign)
        dcoef = np.zeros((N, len(coeff)), dtype=self.dtype)
        up, dn = self._dets
        dcoef[nz] = (
            up[0][nz][:, m0]
            * dn[0][nz][:, m1]
            * np.exp(up[1][nz][:, m0] + dn[1][nz][:, m1] - logv[nz, None])
            / sign[nz, None]
        )
        out = {"det_coeff": dcoef}
        for s, name in ((0, "mo_coeff_alpha"), (1, "mo_coeff_beta")):
            lo = s * self._nelec[0]
            ao_all = self._aovals[:, lo : lo + self._nelec[s]]  # (N, n_s, [nk,] A)
            nmo = self._mo_coeff(s).shape[1]
            A = ao_all.shape[-1]
            # d ln D_d / d C[a, i] = sum_e ao[e, a] inv[d, col(i), e] if orbital i is occupied in d
            per_det = np.zeros((len(self._det_occup[s]), N, A, nmo), dtype=self.dtype)
            for d, occ in enumerate(self._det_occup[s]):
                for col, i in enumerate(occ):
                    ao = self._ao_for_mo(ao_all, s, i)  # (N, n_s, A): the AO set MO i is expanded in
                    per_det[d, :, :, i] = np.einsum("nea,ne->na", ao, self._inverse[s][:, d, col, :])
            g = np.zeros((N, A, nmo), dtype=self.dtype)
            for D, c in enumerate(coeff):
                g += per_det[self._det_map[s][D]] * c * dcoef[:, D, None, None]
            out[name] = g
        return {k: v for k, v in out.items() if v.size > 0}
