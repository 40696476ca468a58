"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference's periodic-boundary path.

Follows (relative to /root/reference):
  * ``enforce_pbc``                              pyqmc/pbc/pbc.py:17-49
  * ``PeriodicConfigs`` / ``PeriodicElectron``   pyqmc/configurations/coord.py:115-252
  * ``MinimalImageDistance``                     pyqmc/configurations/distance.py:83-159
  * lattice-summed GTOs ``_pbc_eval_gto{,_grad,_lap}``   pyqmc/wf/numba/pbcgto.py:98-515
    (image loop ``Ls[:num_Ls[a]]``, per-atom and per-shell r^2 cutoffs, phase table) with the
    set-up of ``PeriodicAtomicOrbitalEvaluator.__init__`` (594-621) and ``max_Ls`` (551-591)
  * ``PBCOrbitalEvaluatorKpoints.aos/mos``       pyqmc/wf/orbitals.py:192-239 (primitive-cell wrap,
    wrap phase ``(-1)**round(k.R/pi)``, per-k MO contraction)
  * determinant flattening over k-points          pyqmc/pyscftools.py:153-186,
                                                  pyqmc/wf/determinant_tools.py:92-106
  * ``Ewald``                                     pyqmc/observables/ewald.py:93-379

Real orbitals only (k-points that fold onto a real twist).  The candidate image list and the
starting radius come from pyscf in the reference (absent here); they are TABLE INPUTS built by
``pyqmc_b200.pbc`` and consumed unchanged by the reference when the golden vectors are generated.
"""
import numpy as np
from scipy.special import erfc

from . import solid_harmonics as sh
from .gto import normalize_shell
from .slater_det import SlaterOracle, pack_determinants


def enforce_pbc(lattvecs, epos):
    recpvecs = np.linalg.inv(lattvecs)
    frac = np.einsum("...ij,jk->...ik", epos, recpvecs)
    wrap, rem = np.divmod(frac, 1)
    return np.dot(rem, lattvecs), wrap


class MinimalImage:
    """distance.py:83-159; ``dist_i(a, b) = minimal(b[:, None] - a)``."""

    def __init__(self, latvec):
        latvec = np.asarray(latvec, dtype=float)
        tol = 1e-10

        def is_diag(M):
            return np.all(np.abs(M - np.diag(np.diagonal(M))) < tol)

        self.mode = "diagonal" if is_diag(latvec) else ("orthogonal" if is_diag(latvec @ latvec.T) else "general")
        self.latvec, self.invvec = latvec, np.linalg.inv(latvec)
        mesh = np.meshgrid(*[np.array(range(3)) for _ in range(3)])
        self.shifts = np.dot(np.stack([m.ravel() for m in mesh], axis=0).T - 1, latvec)

    def minimal(self, d):
        d = np.array(d, dtype=float)
        if self.mode == "diagonal":
            for i in range(3):
                L = self.latvec[i, i]
                d[..., i] = (d[..., i] + L / 2) % L - L / 2
            return d
        if self.mode == "orthogonal":
            frac = np.einsum("...j,jk->...k", d, self.invvec)
            frac = (frac + 0.5) % 1 - 0.5
            return np.einsum("...j,jk->...k", frac, self.latvec)
        allv = d[None] + self.shifts.reshape((-1,) + (1,) * (d.ndim - 1) + (3,))
        idx = np.argmin(np.sum(allv**2, axis=-1), axis=0)
        return np.take_along_axis(allv, idx[None, ..., None], axis=0)[0]

    def __call__(self, d):
        return self.minimal(d)


class PeriodicElectron:
    def __init__(self, epos, lvecs, dist, wrap=None):
        self.configs = epos
        self.lvec = lvecs
        self.wrap = wrap if wrap is not None else np.zeros_like(epos)
        self.dist = dist


class PeriodicWalkers:
    def __init__(self, configs, lvecs, wrap=None, dist=None):
        configs, wrap_ = enforce_pbc(lvecs, configs)
        self.configs = configs
        self.wrap = wrap_
        if wrap is not None:
            self.wrap += wrap
        self.lvecs = np.asarray(lvecs, dtype=float)
        self.dist = dist if dist is not None else MinimalImage(lvecs)

    def electron(self, e):
        return PeriodicElectron(self.configs[:, e], self.lvecs, self.dist, wrap=self.wrap[:, e])

    def make_irreducible(self, e, vec, mask=None):
        if mask is None:
            mask = np.ones(vec.shape[0:-1], dtype=bool)
        epos_, wrap_ = enforce_pbc(self.lvecs, vec[mask])
        epos = vec.copy()
        epos[mask] = epos_
        wrap = self.wrap[:, e, :].copy()
        if vec.ndim == 3:
            wrap = np.repeat(self.wrap[:, e][:, None], vec.shape[1], axis=1)
        wrap[mask] += wrap_
        return PeriodicElectron(epos, self.lvecs, self.dist, wrap=wrap)

    def move(self, e, new, accept):
        self.configs[accept, e, :] = new.configs[accept, :]
        self.wrap[accept, e, :] = new.wrap[accept, :]

    def split(self, n):
        return [PeriodicWalkers(c, self.lvecs, w, self.dist)
                for c, w in zip(np.array_split(self.configs, n), np.array_split(self.wrap, n))]

    def join(self, parts):
        self.configs = np.concatenate([p.configs for p in parts], axis=0)
        self.wrap = np.concatenate([p.wrap for p in parts], axis=0)

    def copy(self):
        return PeriodicWalkers(self.configs.copy(), self.lvecs, self.wrap.copy() - enforce_pbc(self.lvecs, self.configs)[1], self.dist)


class PbcBasis:
    """Lattice-summed AO evaluation on the PRIMITIVE cell (pbcgto.py:98-515)."""

    def __init__(self, cell, kpts, tables):
        self.atom_coords = np.asarray(cell.atom_coords(), dtype=float)
        self.kpts = np.asarray(kpts, dtype=float).reshape(-1, 3)
        self.shells = []  # (atom, l, exps, normalised coefs, ao offset)
        off = 0
        self.max_l = np.zeros(len(self.atom_coords), dtype=int)
        for a in range(len(self.atom_coords)):
            for shell in cell._basis[cell.atom_pure_symbol(a)]:
                l = int(shell[0])
                prim = np.asarray(shell[1:], dtype=float)
                self.shells.append((a, l, prim[:, 0], normalize_shell(l, prim[:, 0], prim[:, 1]), off))
                off += 2 * l + 1
                self.max_l[a] = max(self.max_l[a], l)
        self.nao = off
        self.Ls = np.asarray(tables["Ls"], dtype=float)
        self.num_Ls = np.asarray(tables["num_Ls"])
        self.atom_cutoff = np.asarray(tables["atom_cutoff"], dtype=float)
        self.l_cutoff = np.asarray(tables["l_cutoff"], dtype=float)
        self.phases = np.asarray(tables["phases"])  # complex for general twists (pbcgto.py:620-621)

    def eval(self, deriv, points):
        """points (P,3) inside the primitive cell -> (nk, P, A) or (nk, nc, P, A)."""
        points = np.asarray(points, dtype=float).reshape(-1, 3)
        P, nk = len(points), len(self.kpts)
        nc = (1, 4, 5)[deriv]
        out = np.zeros((nk, nc, P, self.nao), dtype=self.phases.dtype)
        for a, center in enumerate(self.atom_coords):
            rv0 = points - center
            for j in range(int(self.num_Ls[a])):
                rv = rv0 - self.Ls[j]
                r2 = np.sum(rv**2, axis=-1)
                near = r2 <= self.atom_cutoff[a]  # skipped when r2 > cut (pbcgto.py:207)
                if not np.any(near):
                    continue
                x, y, z = rv[:, 0], rv[:, 1], rv[:, 2]
                if deriv == 0:
                    S = sh.evaluate(self.max_l[a], x, y, z)
                else:
                    S, dS = sh.evaluate(self.max_l[a], x, y, z, deriv=True)
                ph = self.phases[j]  # (nk,)
                for ish, (sa, l, al, cf, lo) in enumerate(self.shells):
                    if sa != a:
                        continue
                    # value kernel: r2 < cutoff (pbcgto.py:215); gradient/Laplacian kernels: not r2 > cutoff (348, 490)
                    sel = near & ((r2 < self.l_cutoff[ish]) if deriv == 0 else (r2 <= self.l_cutoff[ish]))
                    if not np.any(sel):
                        continue
                    g = np.exp(-r2[:, None] * al[None, :]) * cf[None, :]
                    R = g.sum(axis=1)
                    Sl = S[l * l:(l + 1) * (l + 1)].T  # (P, 2l+1)
                    comp = np.zeros((nc, P, 2 * l + 1))
                    comp[0] = Sl * R[:, None]
                    if deriv >= 1:
                        Rp = -(g * (2.0 * al)[None, :]).sum(axis=1)
                        dSl = dS[:, l * l:(l + 1) * (l + 1)]
                        cross = np.zeros((P, 2 * l + 1))
                        for i in range(3):
                            dRi = Rp * rv[:, i]
                            comp[1 + i] = dSl[i].T * R[:, None] + Sl * dRi[:, None]
                            cross += dSl[i].T * dRi[:, None]
                        if deriv == 2:
                            Rl = (g * (2.0 * al * (2.0 * al[None, :] * r2[:, None] - 3.0))).sum(axis=1)
                            comp[4] = Sl * Rl[:, None] + 2.0 * cross
                    comp = np.where(sel[None, :, None], comp, 0.0)
                    out[:, :, :, lo:lo + 2 * l + 1] += ph[:, None, None, None] * comp[None]
        return out[:, 0] if deriv == 0 else out


class PbcOrbitals:
    """orbitals.py:118-239 (numba evaluator selected); complex coefficients select the complex wrap phase
    exp(i k.R) (orbitals.py:38-39, 160-165)."""

    def __init__(self, supercell, mo_coeff, kpts, tables):
        self.cell = supercell.original_cell
        self.S = np.asarray(supercell.S, dtype=float)
        self.Lprim = self.cell.lattice_vectors()
        self.kpts = np.asarray(kpts, dtype=float).reshape(-1, 3)
        self.isgamma = np.abs(self.kpts).sum() < 1e-9
        self.basis = PbcBasis(self.cell, self.kpts, tables)
        nper = [np.asarray([m.shape[1] for m in mo]) for mo in mo_coeff]
        self.param_split = [np.cumsum(nper[s]) for s in (0, 1)]
        self.parameters = {
            "mo_coeff_alpha": np.concatenate(mo_coeff[0], axis=1),
            "mo_coeff_beta": np.concatenate(mo_coeff[1], axis=1),
        }
        self.iscomplex = any(np.iscomplexobj(v) for v in self.parameters.values())

    def aos(self, deriv, epos, mask=None):
        """-> ([nc,] shape..., nk, A): k axis next to the AO axis."""
        coords = epos.configs if mask is None else epos.configs[mask]
        shape = coords.shape[:-1]
        flat = coords.reshape(-1, 3)
        prim, primwrap = enforce_pbc(self.Lprim, flat)
        if len(flat) == 0:
            ao = np.zeros((len(self.kpts), self.basis.nao)) if deriv == 0 else np.zeros((len(self.kpts), (1, 4, 5)[deriv], 0, self.basis.nao))
            ao = np.zeros((len(self.kpts), 0, self.basis.nao)) if deriv == 0 else ao
        else:
            ao = self.basis.eval(deriv, prim)
        if not self.isgamma:
            wrap = epos.wrap if mask is None else epos.wrap[mask]
            wrap = np.dot(wrap, self.S).reshape(-1, 3) + primwrap
            kdotR = np.linalg.multi_dot((self.kpts, self.Lprim.T, wrap.T))  # (nk, P)
            phase = np.exp(1j * kdotR) if self.iscomplex else (-1.0) ** np.round(kdotR / np.pi)
            ao = np.einsum("k...,k...a->k...a", phase, ao) if deriv == 0 else np.einsum("kp,kcpa->kcpa", phase, ao)
        if deriv == 0:
            return np.moveaxis(ao, 0, -2).reshape(*shape, len(self.kpts), self.basis.nao)
        ao = np.moveaxis(ao, 0, -2)  # (nc, P, nk, A)
        return ao.reshape(ao.shape[0], *shape, len(self.kpts), self.basis.nao)

    def mos(self, ao, s):
        C = self.parameters["mo_coeff_alpha" if s == 0 else "mo_coeff_beta"]
        ps = [0] + list(self.param_split[s])
        out = np.zeros(ao.shape[:-2] + (C.shape[1],), dtype=complex if self.iscomplex else float)
        for k in range(len(ps) - 1):
            out[..., ps[k]:ps[k + 1]] = ao[..., k, :] @ C[:, ps[k]:ps[k + 1]]
        return out


def kpoint_determinants(supercell, mf, determinants=None, twist=0):
    """pyscftools.py:140-186: primitive k indices of the twist, truncated MO blocks, flattened
    determinant list."""
    from pyqmc_b200 import pbc as hostpbc  # table builders shared with the product (inputs, not arithmetic)

    kinds = hostpbc.create_supercell_twists(supercell, mf)["primitive_ks"][twist]
    if len(kinds) != supercell.scale:
        raise ValueError(f"Found {len(kinds)} k-points but should have found {supercell.scale}.")
    if determinants is None:
        determinants = [(1.0, [[list(np.nonzero(k > 0.5)[0]) for k in s] for s in mf.mo_occ])]

    def f_max_orb(a):
        return int(np.max(a, initial=0)) + 1 if len(a) > 0 else 0

    max_orb = np.amax([[[f_max_orb(k) for k in s] for s in det] for wt, det in determinants], axis=0)
    mo_coeff = [[mf.mo_coeff[s][k][:, 0:max_orb[s][k]] for k in kinds] for s in (0, 1)]
    offs = np.cumsum(max_orb[:, kinds], axis=1)
    offs = np.pad(offs[:, :-1], ((0, 0), (1, 0)))
    flat = []
    for wt, det in determinants:
        fd = []
        for det_s, off_s in zip(det, offs):
            fd.append(list(np.concatenate([np.asarray(det_s[k]) + off_s[ki] for ki, k in enumerate(kinds)]).astype(int)))
        flat.append((wt, fd))
    return mf.kpts[kinds], mo_coeff, flat


class SlaterPbcOracle(SlaterOracle):
    """Slater determinant(s) of Bloch orbitals on a supercell (slater.py with the PBC evaluator)."""

    def __init__(self, supercell, mf, determinants=None, tol=None, twist=0, eval_gto_precision=None):
        from pyqmc_b200 import pbc as hostpbc

        tol = -1 if tol is None else tol
        self._mol = supercell
        self._nelec = tuple(supercell.nelec)
        kpts, mo_coeff, flat = kpoint_determinants(supercell, mf, determinants, twist)
        tables = hostpbc.image_tables(supercell.original_cell, kpts, eval_gto_precision)
        self.orbitals = PbcOrbitals(supercell, mo_coeff, kpts, tables)
        coeff, self._det_occup, self._det_map = pack_determinants(flat, tol)
        self.parameters = {"det_coeff": coeff}
        self.parameters.update(self.orbitals.parameters)
        self.dtype = complex if any(np.iscomplexobj(v) for v in self.parameters.values()) else float

    def _ao(self, deriv, epos, mask=None):
        return self.orbitals.aos(deriv, epos, mask)

    def _mos(self, ao, s):
        return self.orbitals.mos(ao, s)

    def _ao_for_mo(self, ao, s, i):
        """slater.py:509-522 with PBCOrbitalEvaluatorKpoints.pgradient (orbitals.py:241-255): MO i is a
        combination of the AOs of its own k-point only."""
        k = int(np.searchsorted(self.orbitals.param_split[s], i, side="right"))
        return ao[:, :, k, :]


class EwaldOracle:
    """ewald.py:93-379 (energy(): ee, ei, ii)."""

    def __init__(self, cell, ewald_gmax=200, nlatvec=1):
        from pyqmc_b200 import pbc as hostpbc

        t = hostpbc.ewald_tables(cell, ewald_gmax, nlatvec)
        self.t = t
        self.atom_coords = cell.atom_coords()
        self.atom_charges = np.asarray(cell.atom_charges(), dtype=float)
        self.alpha = t["alpha"]

    def _real_cij(self, d):
        cij = np.zeros(d.shape[:-1])
        for ld in self.t["disp"]:
            r = np.linalg.norm(d + ld, axis=-1)
            cij += erfc(self.alpha * r) / r
        return cij

    def energy(self, configs):
        c = configs.configs
        N, ne, _ = c.shape
        t = self.t
        # dist.pairwise(atoms, configs)[m, a, e] = minimal(configs[m, e] - atom[a])
        ei_d = configs.dist(c[:, None, :, :] - self.atom_coords[None, :, None, :])
        ei_real = np.einsum("a,cae->c", -self.atom_charges, self._real_cij(ei_d))
        ee_real = np.zeros(N)
        for i in range(ne):
            for j in range(i + 1, ne):
                ee_real += self._real_cij(configs.dist(c[:, i] - c[:, j]))
        GdotR = np.einsum("hik,jk->hij", c, t["gpoints"])
        ssin, scos = np.sin(GdotR).sum(axis=1), np.cos(GdotR).sum(axis=1)
        ee_rec = np.dot(ssin**2 + scos**2, t["gweight"])
        ei_rec = 2 * np.dot(-t["ion_exp"].real * scos - t["ion_exp"].imag * ssin, t["gweight"])
        ee = ee_real + ee_rec + (ne * (ne - 1) / 2 * t["ijconst"] + ne * t["squareconst"])
        ei = ei_real + ei_rec + (-ne * t["i_sum"] * t["ijconst"])
        return ee, ei, t["ii"]
