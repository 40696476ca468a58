"""Two-body density matrix accumulator on the device objects.

Interface and estimator of the reference's ``TBDMAccumulator`` (``pyqmc/observables/tbdm.py:26-282``; Wagner,
J. Chem. Phys. 138, 094106, Eq. 10), one spin sector (s1, s2) per accumulator, PySCF index order:

    rho_ijkl = < sum_{a != b} Psi(R'_ab)/Psi(R) * phi_i(r_a) phi_k(r_b) phi_j(r_a') phi_l(r_b') / (f1(r_a') f2(r_b')) >

with electron a of spin s1 moved to an auxiliary point r_a' ~ f1 = sum_i |phi^{s1}_i|^2 and electron b of spin s2 to
r_b' ~ f2.  The ratio factorises into ``testvalue(a, r_a')`` and, after ``updateinternals`` has put electron a at
r_a', ``testvalue_many(b's, r_b')`` -- all three on the device (``qmcb_testvalue``, ``qmcb_updateinternals``,
``qmcb_testvalue_many``) -- and the orbitals at the auxiliary points and at the walkers' electrons come from
``qmcb_orbitals_at_points``; the moved electron is put back before the next one is tried.  The two auxiliary walks
consume the global legacy ``np.random`` stream in the reference's order (per spin: ``initial_guess`` + warm-up walk;
per call and spin: the walk's ``randn`` / ``rand`` per sweep, then the ``randint`` assignments), so seeded results
equal the reference's (tests/test_gpu_tbdm.py: golden from the reference).  Open boundary conditions, real orbitals.
"""
import numpy as np

from . import _lib
from .accumulators import _device_context
from .mc import initial_guess


class TBDMAccumulator:
    def __init__(self, mol, orb_coeff, spin, nsweeps=4, tstep=0.50, warmup=200, naux=None, ijkl=None, kpts=None,
                 eval_gto_precision=None):
        if kpts is not None or hasattr(mol, "a"):
            raise NotImplementedError("the B200 TBDM accumulator covers open boundary conditions; use "
                                      "pyqmc.observables.tbdm.TBDMAccumulator on these wave functions for solids")
        if any(np.iscomplexobj(c) for c in orb_coeff):
            raise NotImplementedError("complex orbitals")
        self._coeff = [np.ascontiguousarray(c, dtype=np.float64) for c in orb_coeff]  # [spin] (nao, norb)
        self._sector = tuple(int(s) for s in spin)
        nup, ndn = (int(x) for x in mol.nelec)
        first = (0, nup)
        count = (nup, ndn)
        self._electrons = [np.arange(first[s], first[s] + count[s]) for s in self._sector]
        self._norb = [self._coeff[s].shape[1] for s in (0, 1)]  # the aux walks use spin 0 / spin 1 orbitals (tbdm.py:154-161)
        if ijkl is None:  # the full sector, i, j over the up orbitals and k, l over the down ones (tbdm.py:106-114)
            grid = np.indices((self._norb[0], self._norb[0], self._norb[1], self._norb[1]))
            ijkl = grid.reshape(4, -1).T
        self._ijkl = np.asarray(ijkl, dtype=int).T  # (4, M)
        self._mol, self._tstep, self._nsweeps, self._warmup, self._naux = mol, tstep, nsweeps, warmup, naux
        self.dtype = float
        self._aux = None  # [walk] (naux, 3) positions of the two auxiliary walks
        self._ctx = None

    # ---- orbitals and the auxiliary walks ------------------------------------------------------------------
    def _orbitals(self, points, s):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(pts), self._coeff[s].shape[1]))
        _lib.check(self._ctx.lib.qmcb_orbitals_at_points(self._ctx.h, len(pts), _lib.dptr(pts), self._coeff[s].shape[1],
                                                         _lib.dptr(self._coeff[s]), _lib.dptr(out)))
        return out

    def _walk(self, pos, s, nsamples):
        """``sample_onebody`` (obdm.py:217-247) for the orbitals of spin ``s``: positions and orbital values after
        every Metropolis step of the walk in f(r) = sum_i phi_i(r)^2."""
        orb = self._orbitals(pos, s)
        f = np.sum(orb**2, axis=1)
        history = []
        for _ in range(nsamples):
            trial = pos + np.sqrt(self._tstep) * np.random.randn(len(pos), 1, 3)[:, 0]
            orb_t = self._orbitals(trial, s)
            f_t = np.sum(orb_t**2, axis=1)
            take = f_t / f > np.random.rand(len(pos))
            pos = np.where(take[:, None], trial, pos)
            orb = np.where(take[:, None], orb_t, orb)
            f = np.where(take, f_t, f)
            history.append((pos, orb))
        return history

    def _warm_up(self, naux):
        nelec = int(sum(self._mol.nelec))
        self._aux = []
        for _ in (0, 1):  # both warm-up walks sample the spin-0 orbitals (tbdm.py:128-130)
            start = initial_guess(self._mol, int(naux / nelec) + 1).configs.reshape(-1, 3)[:naux].copy()
            walk = self._walk(start, 0, self._warmup)
            self._aux.append(walk[-1][0] if walk else start)

    # ---- accumulator protocol -----------------------------------------------------------------------------
    def __call__(self, configs, wf):
        self._ctx = _device_context(wf)
        if self._ctx is None or self._ctx.nconf == 0:
            raise RuntimeError("wf.recompute(configs) must be called before the TBDM accumulator")
        nconf = configs.configs.shape[0]
        if self._aux is None:
            self._warm_up(nconf if self._naux is None else self._naux)
        # per walk: nsweeps steps, then which auxiliary walker every configuration uses in every sweep
        walks, picks = [], []
        for s in (0, 1):
            walks.append(self._walk(self._aux[s], s, self._nsweeps))
            picks.append(np.random.randint(0, len(self._aux[s]), size=(self._nsweeps, nconf)))
            self._aux[s] = walks[s][-1][0]
        i_, j_, k_, l_ = self._ijkl
        ea_list, eb_all = self._electrons
        # orbitals of the sector's electrons at their own positions, already gathered to the requested index lists
        phi_a = self._orbitals(configs.configs[:, ea_list], 0).reshape(nconf, len(ea_list), -1)[:, :, i_]
        phi_b = self._orbitals(configs.configs[:, eb_all], 1).reshape(nconf, len(eb_all), -1)[:, :, k_]
        value = np.zeros((nconf, self._ijkl.shape[1]))
        norm = [np.zeros((nconf, self._norb[0])), np.zeros((nconf, self._norb[1]))]
        for sweep in range(self._nsweeps):
            pos, orb, dens = [], [], []
            for s in (0, 1):
                p, o = walks[s][sweep]
                pos.append(p[picks[s][sweep]])
                orb.append(o[picks[s][sweep]])
                dens.append(np.sum(orb[s]**2, axis=1))
                norm[s] += orb[s]**2 / dens[s][:, None]
            aux_a = configs.make_irreducible(0, pos[0])
            aux_b = configs.make_irreducible(0, pos[1])
            primed = orb[0][:, j_] * orb[1][:, l_] / (dens[0] * dens[1])[:, None]  # phi_j(r_a') phi_l(r_b') / (f1 f2)
            for ia, ea in enumerate(ea_list):
                others = eb_all != ea  # the same electron is not moved twice
                ratio_a, saved = wf.testvalue(ea, aux_a)
                wf.updateinternals(ea, aux_a, configs, saved_values=saved)
                ratio_b = wf.testvalue_many(eb_all[others], aux_b)
                wf.updateinternals(ea, configs.electron(ea), configs)  # electron a back where the walkers have it
                partner = np.einsum("nb,nbo->no", ratio_b, phi_b[:, others, :])
                value += (ratio_a[:, None] * phi_a[:, ia, :]) * partner * primed
        scale = 1.0 / self._nsweeps
        return {"value": value * scale, "norm_a": norm[0] * scale, "norm_b": norm[1] * scale}

    def avg(self, configs, wf):
        return {k: np.mean(v, axis=0) for k, v in self(configs, wf).items()}

    def keys(self):
        return {"value", "norm_a", "norm_b"}

    def shapes(self):
        return {"value": (self._ijkl.shape[1],), "norm_a": (self._norb[self._sector[0]],),
                "norm_b": (self._norb[self._sector[1]],)}


def normalize_tbdm(tbdm, norm_a, norm_b):
    """rho_ijkl / sqrt(n_i n_j n_k n_l) in PySCF's index order (tbdm.py:293-297)."""
    return tbdm / np.sqrt(np.einsum("i,j,k,l->ijkl", norm_a, norm_a, norm_b, norm_b))
