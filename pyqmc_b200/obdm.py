"""One-body density matrix accumulator on the device objects.

Interface and estimator of the reference's ``OBDMAccumulator`` (``pyqmc/observables/obdm.py:25-214``; Wagner,
J. Chem. Phys. 138, 094106, Eq. 9): an auxiliary position r' is sampled from f(r) = sum_i |phi_i(r)|^2 by its own
Metropolis walk, and

    rho_ij = < Psi(R')/Psi(R) * phi_i(r') phi_j(r_e) / f(r') >      R' = R with electron e moved to r',

summed over the chosen electrons e.  Three things are evaluated per sweep, all on the GPU: the orbitals at the
auxiliary points and at every electron (``qmcb_orbitals_at_points``), and the wave-function ratios for moving EACH
electron to the same auxiliary point (``wf.testvalue_many`` -> ``qmcb_testvalue_many``); the (N, norb, norb)
contraction is a host einsum.  The global legacy ``np.random`` stream is consumed in the reference's order
(warm-up ``initial_guess`` + walk; per call ``randint`` assignments, then per sweep ``randn`` shifts and ``rand``
acceptances), so seeded results equal the reference's (tests/test_gpu_obdm.py: golden from the reference).
Open boundary conditions, real orbitals.
"""
import numpy as np

from . import _lib
from .accumulators import _device_context
from .mc import initial_guess


class OBDMAccumulator:
    def __init__(self, mol, orb_coeff, nsweeps=5, tstep=0.50, warmup=10000, naux=None, spin=None, electrons=None,
                 kpts=None, eval_gto_precision=None):
        if kpts is not None or hasattr(mol, "a"):
            raise NotImplementedError("the B200 OBDM accumulator covers open boundary conditions; use "
                                      "pyqmc.observables.obdm.OBDMAccumulator on these wave functions for solids")
        nup, ndn = (int(x) for x in mol.nelec)
        if spin is not None:
            if spin not in (0, 1):
                raise ValueError("Spin not equal to 0 or 1")
            electrons = np.arange(0, nup) if spin == 0 else np.arange(nup, nup + ndn)
        elif electrons is None:
            electrons = np.arange(nup + ndn)
        self._electrons = np.asarray(electrons, dtype=int)
        self._coeff = np.ascontiguousarray(orb_coeff, dtype=np.float64)
        if np.iscomplexobj(orb_coeff):
            raise NotImplementedError("complex orbitals")
        self.norb = self._coeff.shape[1]
        self.nelec = len(self._electrons)
        self.dtype = float
        self._mol, self._tstep, self._nsweeps, self._warmup, self._naux = mol, tstep, nsweeps, warmup, naux
        self._aux = None  # (naux, 3) positions of the auxiliary walk
        self._ctx = None

    # ---- orbital evaluation on the device ------------------------------------------------------------
    def _orbitals(self, points):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(pts), self.norb))
        _lib.check(self._ctx.lib.qmcb_orbitals_at_points(self._ctx.h, len(pts), _lib.dptr(pts), self.norb,
                                                         _lib.dptr(self._coeff), _lib.dptr(out)))
        return out

    def _walk(self, nsamples):
        """Metropolis walk of the auxiliary points in f(r) = sum_i phi_i(r)^2 (``sample_onebody``, obdm.py:217-247):
        returns the positions and orbital values after every step."""
        pos = self._aux
        orb = self._orbitals(pos)
        f = np.sum(orb**2, axis=1)
        history = []
        for _ in range(nsamples):
            trial = pos + np.sqrt(self._tstep) * np.random.randn(len(pos), 1, 3)[:, 0]
            orb_t = self._orbitals(trial)
            f_t = np.sum(orb_t**2, axis=1)
            take = f_t / f > np.random.rand(len(pos))
            pos = np.where(take[:, None], trial, pos)
            orb = np.where(take[:, None], orb_t, orb)
            f = np.where(take, f_t, f)
            history.append((pos, orb))
        self._aux = pos
        return history

    def _warm_up(self, naux):
        start = initial_guess(self._mol, int(naux / self.nelec) + 1)
        self._aux = start.configs.reshape(-1, 3)[:naux].copy()
        self._walk(self._warmup)

    # ---- accumulator protocol -----------------------------------------------------------------------------
    def __call__(self, configs, wf):
        self._ctx = _device_context(wf)
        if self._ctx is None or self._ctx.nconf == 0:
            raise RuntimeError("wf.recompute(configs) must be called before the OBDM accumulator")
        nconf = len(configs.configs)
        if self._aux is None:
            self._warm_up(nconf if self._naux is None else self._naux)
        naux = len(self._aux)
        assign = np.random.randint(0, naux, size=(self._nsweeps, nconf))
        history = self._walk(self._nsweeps)
        orb_e = self._orbitals(configs.configs[:, self._electrons]).reshape(nconf, self.nelec, self.norb)
        value = np.zeros((nconf, self.norb, self.norb))
        norm = np.zeros((nconf, self.norb))
        for (pos, orb), pick in zip(history, assign):
            aux_pos, aux_orb = pos[pick], orb[pick]
            weight = aux_orb**2
            fsum = np.sum(weight, axis=-1, keepdims=True) / self.norb
            ratios = wf.testvalue_many(self._electrons, configs.make_irreducible(0, aux_pos))
            value += np.einsum("ie,ij,iek->ijk", ratios, aux_orb / fsum, orb_e, optimize=True)
            norm += weight / fsum
        # the reference keeps, as the walk's next starting point, the LAST sweep's points AFTER re-assigning them
        # to the walkers (obdm.py:143-146 resamples the very object stored in _extra_config)
        self._aux = history[-1][0][assign[-1]]
        return {"value": value / self._nsweeps, "norm": norm / self._nsweeps}

    def avg(self, configs, wf):
        per_walker = self(configs, wf)
        return {name: per_walker[name].mean(axis=0) for name in per_walker}

    def keys(self):
        return {"value", "norm"}

    def shapes(self):
        return {"value": (self.norb, self.norb), "norm": (self.norb,)}


def normalize_obdm(obdm, norm):
    """rho_ij / sqrt(norm_i norm_j) (obdm.py:250-251)."""
    return obdm / np.sqrt(np.outer(norm, norm))
