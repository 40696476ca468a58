"""Local-energy accumulator for the B200 wave functions.

Same interface as the reference ``EnergyAccumulator`` (``pyqmc/observables/accumulators.py:45-95``;
contract tested at ``tests/unit/test_accumulators.py:72-83``): ``__call__(configs, wf)`` returns
per-walker ``ke, ee, ei, ecp, grad2, total``; ``avg`` their walker means; ``keys``/``shapes``;
``nonlocal_tmoves`` / ``has_nonlocal_moves`` for DMC.

The whole evaluation (open-boundary Coulomb ``energy.py:28-44``, kinetic ``energy.py:57-65``,
semi-local ECP ``eval_ecp.py:21-146``) is one device pipeline operating on the walker state the
wave function already holds.  Random variates are drawn on the host from the global legacy
``np.random`` stream in the reference's order -- per (electron, ECP atom): ``np.random.random(N)``
(eval_ecp.py:145) then one ``scipy Rotation.random()`` (eval_ecp.py:263) -- so seeded runs
reproduce the reference's stochastic ECP masks and rotations.
"""
import itertools

import numpy as np
import scipy.spatial.transform

from . import _lib, quadrature
from .wf import MultiplyWF, _DeviceFactor

KEYS = ("ke", "ee", "ei", "ecp", "grad2", "total")
_SERIAL = itertools.count(1)  # identity of an accumulator's device tables (id() is recycled after garbage collection)


def flatten_ecp(mol, naip=None):
    """mol._ecp -> arrays for qmcb_set_ecp (channel columns l = 0..lmax, then local l = -1)."""
    ecp_atom, chan_off, term_off, power, alpha, coef, naips, quad = [], [0], [0], [], [], [], [], []
    for i, (sym, _) in enumerate(mol._atom):
        if sym not in mol._ecp:
            continue
        chans = {}
        for l, expand in mol._ecp[sym][1]:  # eval_ecp.py:160-179
            terms = []
            for n, lines in enumerate(expand):
                for a, c in lines:
                    terms.append((n - 2, float(a), float(c)))
            chans[int(l)] = terms
        nl = len(chans)
        if sorted(chans) != list(range(-1, nl - 1)):
            raise ValueError(f"ECP channels of {sym} must be l = -1, 0, ..., lmax; got {sorted(chans)}")
        ecp_atom.append(i)
        for l in list(range(nl - 1)) + [-1]:
            for n, a, c in chans[l]:
                power.append(n)
                alpha.append(a)
                coef.append(c)
            term_off.append(len(power))
        chan_off.append(chan_off[-1] + nl)
        this_naip = naip if naip is not None else (6 if nl <= 2 else 12)  # eval_ecp.py:239-240
        naips.append(this_naip)
        pts, wts = quadrature.grid(this_naip)
        quad.extend(pts.reshape(-1))
        quad.extend(wts)
    return dict(ecp_atom=_lib.i32(ecp_atom), chan_off=_lib.i32(chan_off), term_off=_lib.i32(term_off),
                power=_lib.i32(power), alpha=_lib.f64(alpha), coef=_lib.f64(coef), naip=_lib.i32(naips),
                quad=_lib.f64(quad))


def _device_context(wf):
    if isinstance(wf, MultiplyWF) and wf._fused:
        return wf._ctx
    if isinstance(wf, _DeviceFactor):
        return wf._ctx
    raise TypeError("pyqmc_b200.EnergyAccumulator needs a pyqmc_b200 wave function whose state lives on "
                    "the device (Slater, JastrowSpin or a fused MultiplyWF of both); got " + type(wf).__name__)


class EnergyAccumulator:
    """Returns local energy of each configuration in a dictionary."""

    def __init__(self, mol, threshold=10, naip=None, use_old_ecp=True, **kwargs):
        if not use_old_ecp:
            raise NotImplementedError("only the default ECP path (use_old_ecp=True) is implemented")
        self.mol, self.threshold, self.naip = mol, threshold, naip
        self._serial = next(_SERIAL)
        self._ecp = flatten_ecp(mol, naip)
        # compute_tmoves is called without naip by the reference's nonlocal_tmoves (accumulators.py:80-81): T-move tables
        # always use the default quadrature sizes (6 / 12 points), whatever the energy evaluation uses
        self._ecp_tmoves = self._ecp if naip is None else flatten_ecp(mol, None)
        self.necp = len(self._ecp["ecp_atom"])
        self._ewald = None
        if hasattr(mol, "a"):  # accumulators.py:52-55: Ewald replaces the open-boundary Coulomb sums
            from . import pbc

            self._ewald = pbc.ewald_tables(mol, **kwargs)

    def _attach(self, wf, tmoves=False):
        ctx = _device_context(wf)
        if ctx is None or ctx.nconf == 0:
            raise RuntimeError("wf.recompute(configs) must be called before the energy accumulator")
        tables = self._ecp_tmoves if tmoves else self._ecp
        key = (self._serial, self.threshold, self.naip, tables is self._ecp)
        if ctx.ecp_key != key:
            t = tables
            _lib.check(ctx.lib.qmcb_set_ecp(ctx.h, self.necp, _lib.iptr(t["ecp_atom"]), _lib.iptr(t["chan_off"]),
                                            _lib.iptr(t["term_off"]), _lib.iptr(t["power"]), _lib.dptr(t["alpha"]),
                                            _lib.dptr(t["coef"]), _lib.iptr(t["naip"]), _lib.dptr(t["quad"]),
                                            float(self.threshold)))
            if self._ewald is not None:
                t = self._ewald
                disp, gp, gw = _lib.f64(t["disp"]), _lib.f64(t["gpoints"]), _lib.f64(t["gweight"])
                ire, iim = _lib.f64(np.real(t["ion_exp"])), _lib.f64(np.imag(t["ion_exp"]))
                _lib.check(ctx.lib.qmcb_set_ewald(ctx.h, t["alpha"], len(disp), _lib.dptr(disp), len(gw), _lib.dptr(gp),
                                                  _lib.dptr(gw), _lib.dptr(ire), _lib.dptr(iim), t["ijconst"],
                                                  t["squareconst"], t["i_sum"], t["ii"]))
            ctx.ecp_key = key
        return ctx

    def draw_ecp_variates(self, nconf, nelec):
        """(u [ne][necp][N], rot [ne][necp][3][3]) from the global legacy RNG, reference order."""
        u = np.empty((nelec, self.necp, nconf))
        rot = np.empty((nelec, self.necp, 3, 3))
        for e in range(nelec):
            for a in range(self.necp):
                u[e, a] = np.random.random(size=nconf)
                rot[e, a] = scipy.spatial.transform.Rotation.random().as_matrix()
        return u, rot

    def __call__(self, configs, wf):
        ctx = self._attach(wf)
        nconf, nelec = configs.configs.shape[:2]
        if nconf != ctx.nconf:
            raise ValueError("configs and the wave function's internal state disagree on the walker count")
        u, rot = self.draw_ecp_variates(nconf, nelec)
        out = np.empty((8 if ctx.cplx else 6, nconf))
        _lib.check(ctx.lib.qmcb_energy(ctx.h, _lib.dptr(u), _lib.dptr(rot), _lib.dptr(out)))
        res = {k: out[i] for i, k in enumerate(KEYS)}
        if ctx.cplx:  # the ECP values (and with them the total) carry wf.dtype (eval_ecp.py:26)
            res["ecp"] = out[3] + 1j * out[6]
            res["total"] = out[5] + 1j * out[7]
        return res

    def avg(self, configs, wf):
        per_walker = self(configs, wf)
        return {k: np.mean(per_walker[k], axis=0) for k in KEYS}

    def nonlocal_tmoves(self, configs, wf, e, tau):
        """T-move candidates of electron e (``compute_tmoves``, eval_ecp.py:43-80): ratio (N, M),
        weight (N, M) and the candidate positions (N, M, 3), M = sum of the quadrature sizes of the
        ECP atoms.  Random variates as in the reference: per ECP atom ``random(N)`` then a rotation."""
        ctx = self._attach(wf, tmoves=True)
        nconf = configs.configs.shape[0]
        if self.necp == 0:
            return {"ratio": np.ones((nconf, 0)), "weight": np.zeros((nconf, 0))}
        u = np.empty((self.necp, nconf))
        rot = np.empty((self.necp, 3, 3))
        for a in range(self.necp):
            u[a] = np.random.random(size=nconf)
            rot[a] = scipy.spatial.transform.Rotation.random().as_matrix()
        M = int(np.sum(self._ecp_tmoves["naip"]))
        ratio = np.empty((nconf, M), dtype=complex if ctx.cplx else float)
        weight, epos = np.empty((nconf, M)), np.empty((nconf, M, 3))
        _lib.check(ctx.lib.qmcb_tmoves(ctx.h, int(e), float(tau), _lib.dptr(u), _lib.dptr(rot), _lib.dptr(ratio),
                                       _lib.dptr(weight), _lib.dptr(epos)))
        return {"ratio": ratio, "weight": weight, "configs": configs.make_irreducible(e, epos)}

    def has_nonlocal_moves(self):
        return len(self.mol._ecp) > 0

    def keys(self):
        return set(KEYS)

    def shapes(self):
        return {k: () for k in KEYS}
