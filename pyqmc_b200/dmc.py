"""DMC driver with the reference's signatures (``pyqmc/method/dmc.py``).

``limdrift`` (22-35), ``get_V2`` (38-46), ``propose_drift_diffusion`` (49-70), ``propose_tmoves``
(73-120), ``dmc_propagate`` (123-221), ``compute_S`` (224-235), ``branch`` (342-376) and ``rundmc``
(412-586, without the HDF5 restart file) keep the reference's arguments, RNG consumption order and
output dictionaries.

When the wave function is a fused single-determinant Slater-Jastrow on an open-boundary system and
the only accumulator is a ``pyqmc_b200.EnergyAccumulator``, ``dmc_propagate`` runs DEVICE-RESIDENT:
every random variate of the block is drawn up front from the global legacy ``np.random`` stream in
exactly the order the reference loop consumes it, shipped once, and ``qmcb_dmc_block`` executes the
T-moves, drift-diffusion sweeps, local energies and weight updates without host round trips.  Any
other combination goes through the generic loop over the wave-function protocol calls.
"""
import logging

import numpy as np
import scipy.spatial.transform

from . import _lib, mc
from .accumulators import KEYS, EnergyAccumulator, _device_context
from .wf import JASTROW, SLATER


def limdrift(g, tau, acyrus=0.5):
    v2 = np.sum(g**2, axis=1)
    mask = v2 > 1e-8
    taueff = np.ones(v2.shape) * tau
    taueff[mask] = (np.sqrt(1 + 2 * tau * acyrus * v2[mask]) - 1) / (acyrus * v2[mask])
    return g * taueff[:, np.newaxis]


def get_V2(configs, wf, acc_out):
    if "grad2" in acc_out.keys():
        return acc_out["grad2"]
    nconfig, nelec = configs.configs.shape[0:2]
    v2 = np.zeros(nconfig)
    for e in range(nelec):
        v2 += np.sum(np.abs(wf.gradient(e, configs.electron(e))).T ** 2, axis=1)
    return v2


def propose_drift_diffusion(wf, configs, tstep, e):
    nconfig = configs.configs.shape[0]
    grad = limdrift(np.real(wf.gradient(e, configs.electron(e)).T), tstep)
    gauss = np.random.normal(scale=np.sqrt(tstep), size=(nconfig, 3))
    eposnew = configs.configs[:, e, :] + gauss + grad
    newepos = configs.make_irreducible(e, eposnew)
    g, wfratio, saved = wf.gradient_value(e, newepos)
    new_grad = limdrift(np.real(g.T), tstep)
    forward = np.sum(gauss**2, axis=1)
    backward = np.sum((gauss + grad + new_grad) ** 2, axis=1)
    t_prob = np.exp(1 / (2 * tstep) * (forward - backward))
    ratio = np.abs(wfratio) ** 2 * t_prob
    if wf.dtype == float:
        ratio *= np.sign(wfratio)
    accept = ratio > np.random.rand(nconfig)
    r2 = np.sum((gauss + grad) ** 2, axis=1)
    return newepos, accept, r2, saved


def propose_tmoves(wf, configs, energy_accumulator, tstep, e):
    moves = energy_accumulator.nonlocal_tmoves(configs, wf, e, tstep)
    t_amplitudes = moves["ratio"] * moves["weight"]
    forward_probability = np.zeros_like(t_amplitudes)
    forward_probability[t_amplitudes > 0] = t_amplitudes[t_amplitudes > 0]
    norm = 1.0 + np.sum(forward_probability, axis=1)
    cdf = np.cumsum(forward_probability / norm[:, np.newaxis], axis=1)
    selected_moves = np.array([np.searchsorted(row, np.random.rand()) for row in cdf], dtype=int).reshape(len(cdf))
    move_selected = selected_moves < t_amplitudes.shape[1]
    newpos = np.zeros((norm.shape[0], 3))
    reverse_ratio = np.zeros((norm.shape[0]))
    backward_amplitudes = t_amplitudes.copy()
    for walker, move in enumerate(selected_moves):
        if move_selected[walker]:
            newpos[walker, :] = moves["configs"].configs[walker, move, :]
            reverse_ratio[walker] = 1.0 / moves["ratio"][walker, move]
            backward_amplitudes[walker, :] *= reverse_ratio[walker]
            backward_amplitudes[walker, move] = reverse_ratio[walker] * moves["weight"][walker, move]
        else:
            newpos[walker, :] = configs.configs[walker, e, :]
            reverse_ratio[walker] = 0.0
    newpos = configs.make_irreducible(e, newpos)
    backward_amplitudes[backward_amplitudes < 0] = 0.0
    back_norm = 1.0 + np.sum(backward_amplitudes, axis=1)
    acceptance = norm / back_norm
    acceptance[move_selected == False] = 0.0  # noqa: E712
    return newpos, move_selected, acceptance, np.sum(t_amplitudes)


def compute_S(e_trial, e_est, branchcut, v2, tau, eloc, nelec):
    e_cut = e_est - eloc
    mask = np.abs(e_cut) > branchcut
    e_cut[mask] = branchcut * np.sign(e_cut[mask])
    denominator = np.sqrt(1 + (v2 * tau / nelec) ** 2)
    return e_trial - e_est + e_cut / denominator


def _device_dmc_path(wf, accumulators, ekey):
    """Device-resident propagation: fused single-determinant Slater x JastrowSpin, open boundary
    conditions, one EnergyAccumulator under ``ekey[0]``."""
    try:
        ctx = _device_context(wf)
    except TypeError:
        return False
    if len(accumulators) != 1 or not isinstance(accumulators.get(ekey[0]), EnergyAccumulator) or ekey[1] != "total":
        return False
    which = getattr(wf, "_which", 0)
    if which & ~(SLATER | JASTROW) or not (which & SLATER):
        return False
    factors = getattr(wf, "wf_factors", [wf])
    if len(factors[0].parameters["det_coeff"]) != 1:
        return False
    mol = factors[0]._mol
    del ctx
    return not hasattr(mol, "a")


def _dmc_buffers(shapes, pinned_owner, slot=0):
    """Arrays for one block's variates; page-locked (true async H2D) and cached on the device context
    (one set per ``slot``: the prefetcher fills one set while the device reads the other) when an owner
    is given."""
    if pinned_owner is None:
        return {k: np.empty(s) for k, s in shapes.items()}
    key = tuple(sorted((k, tuple(s)) for k, s in shapes.items()))
    cache = pinned_owner.__dict__.setdefault("_dmc_buffers", {}).setdefault(slot, {})
    if cache.get("key") != key:
        cache["key"] = key
        cache["own"] = {k: _lib.PinnedArray(s) for k, s in shapes.items()}
    return {k: v.array for k, v in cache["own"].items()}


def draw_dmc_block_variates(nconf, nelec, tstep, nsteps, accumulator, native=True, pinned_owner=None, slot=0):
    """Every random number of one ``dmc_propagate`` call in the reference's order: the energy
    evaluation before the first step; then per step, for every electron the T-move draws
    (``nonlocal_tmoves``: per ECP atom ``random(N)`` + a rotation; ``select_walker``: one ``rand()``
    per walker; acceptance ``rand(N)``), for every electron ``normal(N, 3)`` + ``rand(N)``, and the
    energy evaluation."""
    necp = accumulator.necp
    b = _dmc_buffers(dict(ecp_u=(nsteps + 1, nelec, necp, nconf), ecp_rot=(nsteps + 1, nelec, necp, 3, 3),
                          tm_u=(nsteps, nelec, necp, nconf), tm_rot=(nsteps, nelec, necp, 3, 3),
                          tm_sel=(nsteps, nelec, nconf), tm_acc=(nsteps, nelec, nconf),
                          gauss=(nsteps, nelec, nconf, 3), unif=(nsteps, nelec, nconf)), pinned_owner, slot)
    ecp_u, ecp_rot, tm_u, tm_rot = b["ecp_u"], b["ecp_rot"], b["tm_u"], b["tm_rot"]
    tm_sel, tm_acc, gauss, unif = b["tm_sel"], b["tm_acc"], b["gauss"], b["unif"]
    tmoves = accumulator.has_nonlocal_moves()
    # the draw program, in consumption order: (kind, destination array view, scale)
    ops = []

    def energy_draws(k):
        for e in range(nelec):
            for a in range(necp):
                ops.append((0, ecp_u[k, e, a], 1.0))
                ops.append((2, ecp_rot[k, e, a], 1.0))

    energy_draws(0)
    for step in range(nsteps):
        if tmoves:
            for e in range(nelec):
                for a in range(necp):
                    ops.append((0, tm_u[step, e, a], 1.0))
                    ops.append((2, tm_rot[step, e, a], 1.0))
                ops.append((0, tm_sel[step, e], 1.0))  # == nconf successive scalar rand() calls
                ops.append((0, tm_acc[step, e], 1.0))
        for e in range(nelec):
            ops.append((1, gauss[step, e], float(np.sqrt(tstep))))
            ops.append((0, unif[step, e], 1.0))
        energy_draws(step + 1)
    if not (native and _run_draw_program_native(ops)):
        for kind, dst, scale in ops:
            if kind == 0:
                dst[...] = np.random.random(size=dst.shape)
            elif kind == 1:
                dst[...] = np.random.normal(scale=scale, size=dst.shape)
            else:
                dst[...] = scipy.spatial.transform.Rotation.random().as_matrix()
    return dict(ecp_u=ecp_u, ecp_rot=ecp_rot, tm_u=tm_u, tm_rot=tm_rot, tm_sel=tm_sel, tm_acc=tm_acc, gauss=gauss,
                unif=unif)


def _run_draw_program_native(ops):
    """Runs the draw program in csrc/legacy_rng.cpp on the state of the global legacy generator
    (bit-identical to the numpy / scipy calls, several times faster); False if not applicable."""
    import ctypes

    state = np.random.get_state()
    if state[0] != "MT19937":
        return False
    for _, dst, _ in ops:
        if not (dst.flags["C_CONTIGUOUS"] and dst.dtype == np.float64):
            return False
    lib = _lib.load()
    kind = np.array([o[0] for o in ops], dtype=np.int32)
    count = np.array([o[1].size for o in ops], dtype=np.int64)
    dst = np.array([o[1].ctypes.data for o in ops], dtype=np.uint64)
    scale = np.array([o[2] for o in ops], dtype=np.float64)
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = ctypes.c_int32(int(state[2]))
    has_gauss = ctypes.c_int32(int(state[3]))
    cached = ctypes.c_double(float(state[4]))
    rc = lib.qmcb_rng_program(key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(pos),
                              ctypes.byref(has_gauss), ctypes.byref(cached), len(ops), _lib.iptr(kind),
                              count.ctypes.data_as(_lib.c_i64_p), dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                              _lib.dptr(scale), mc.rng_threads())
    if rc != 0:
        return False
    np.random.set_state(("MT19937", key, pos.value, has_gauss.value, cached.value))
    return True


class DmcPrefetcher:
    """Draws the variates of block b+1 on a host thread while block b runs on the device.  The global
    legacy stream is consumed in the reference's order -- block variates, then the one ``rand()`` of
    ``branch`` (dmc.py:361), then the next block's variates -- so the branching draw of block b is
    taken by the same thread just before it draws block b+1; nothing else touches ``np.random`` while
    the thread runs."""

    def __init__(self, wf, configs, tstep, nsteps, accumulator, nblocks, with_branch=True):
        from concurrent.futures import ThreadPoolExecutor

        self.shape = configs.configs.shape[:2]
        self.args = (tstep, nsteps, accumulator)
        if _device_context(wf) is None:
            wf.recompute(configs)
        self.ctx = _device_context(wf)
        self.remaining = nblocks
        self.with_branch = with_branch
        self.issued = 0
        self.pool = ThreadPoolExecutor(max_workers=1)
        self.future = self.pool.submit(self._draw, False)
        self.remaining -= 1

    def _draw(self, branch_first):
        base = np.random.rand() if branch_first else None
        tstep, nsteps, accumulator = self.args
        slot = self.issued % 2
        self.issued += 1
        v = draw_dmc_block_variates(self.shape[0], self.shape[1], tstep, nsteps, accumulator, pinned_owner=self.ctx, slot=slot)
        return base, v

    def next(self):
        """Variates of the next block; also starts drawing the one after (preceded by this block's
        branching draw, returned by ``branch_draw``)."""
        self._prev_base, v = self.future.result()
        if self.remaining > 0:
            self.future = self.pool.submit(self._draw, self.with_branch)
            self.remaining -= 1
        else:
            self.future = None
        return v

    def branch_draw(self):
        """The ``np.random.rand()`` of this block's ``branch`` call."""
        if self.future is None:
            self.pool.shutdown(wait=True)
            return np.random.rand() if self.with_branch else None
        base, v = self.future.result()  # waits for the thread: it drew the branching number first
        self.future = _Done((None, v))
        return base


class _Done:
    def __init__(self, value):
        self.value = value

    def result(self):
        return self.value


def dmc_propagate_device(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps, accumulators, ekey,
                         variates=None):
    nconf, nelec, _ = configs.configs.shape
    wf.recompute(configs)
    ctx = _device_context(wf)
    accumulator = accumulators[ekey[0]]
    accumulator._attach(wf)
    v = variates if variates is not None else draw_dmc_block_variates(nconf, nelec, tstep, nsteps, accumulator,
                                                                        pinned_owner=ctx)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    newconf = np.empty((nconf, nelec, 3))
    wsums = np.zeros((nsteps, 8))
    nacc = np.zeros((nsteps, nelec), dtype=np.int64)
    ntacc = np.zeros((nsteps, nelec), dtype=np.int64)
    d = _lib.dptr
    _lib.check(ctx.lib.qmcb_dmc_block(
        ctx.h, nsteps, float(tstep), float(branchcut_start), float(e_trial), float(e_est), d(v["gauss"]), d(v["unif"]),
        d(v["ecp_u"]), d(v["ecp_rot"]), d(v["tm_u"]), d(v["tm_rot"]), d(v["tm_sel"]), d(v["tm_acc"]), d(w), d(newconf),
        d(wsums), nacc.ctypes.data_as(_lib.c_i64_p), ntacc.ctypes.data_as(_lib.c_i64_p)))
    ctx.epoch += 1  # the block moved the walkers on the device
    configs.configs[...] = newconf
    if w is not weights:
        weights[...] = w
    rows = []
    for step in range(nsteps):
        wavg = wsums[step, 6] / nconf
        avg = {ekey[0] + k: wsums[step, i] / (nconf * wavg) for i, k in enumerate(KEYS)}
        avg["weight"] = wavg
        avg["acceptance"] = np.mean(nacc[step] / nconf)
        avg["tmove_acceptance"] = np.mean(ntacc[step] / nconf)
        rows.append(avg)
    return _collect(rows), configs, weights


def _collect(df):
    weight = np.asarray([d["weight"] for d in df])
    avg_weight = weight / np.mean(weight)
    df_ret = {k: np.mean([d[k] * w for d, w in zip(df, avg_weight)], axis=0) for k in df[0].keys()}
    df_ret["weight"] = np.mean(weight)
    return df_ret


def dmc_propagate(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps=5, accumulators=None,
                  ekey=("energy", "total"), variates=None):
    """Propagate DMC without branching (dmc.py:123-221).  ``variates``: pre-drawn random numbers of the
    block (device-resident path only; see ``DmcPrefetcher``)."""
    assert accumulators is not None, "Need an energy accumulator for DMC"
    if _device_dmc_path(wf, accumulators, ekey):
        return dmc_propagate_device(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps, accumulators,
                                    ekey, variates=variates)
    nconfig, nelec = configs.configs.shape[0:2]
    wf.recompute(configs)
    energy_acc = accumulators[ekey[0]](configs, wf)
    eloc = energy_acc[ekey[1]].real
    v2 = get_V2(configs, wf, energy_acc)
    df = []
    for _ in range(nsteps):
        r2_accepted = np.zeros(nconfig)
        r2_proposed = np.zeros(nconfig)
        prob_acceptance = np.zeros(nconfig)
        tmove_acceptance = np.zeros(nconfig)
        if accumulators[ekey[0]].has_nonlocal_moves():
            for e in range(nelec):
                newepos, mask, probability, _ = propose_tmoves(wf, configs, accumulators[ekey[0]], tstep, e)
                accept = mask & (probability > np.random.rand(nconfig))
                configs.move(e, newepos, accept)
                wf.updateinternals(e, newepos, configs, mask=accept)
                tmove_acceptance += accept / nelec
        for e in range(nelec):
            newepos, accept, r2, saved = propose_drift_diffusion(wf, configs, tstep, e)
            configs.move(e, newepos, accept)
            wf.updateinternals(e, newepos, configs, mask=accept, saved_values=saved)
            r2_proposed += r2
            r2_accepted[accept] += r2[accept]
            prob_acceptance += accept / nelec
        elocold = eloc.copy()
        v2old = v2.copy()
        energydat = accumulators[ekey[0]](configs, wf)
        eloc = energydat[ekey[1]].real
        tdamp = r2_accepted / r2_proposed
        v2 = get_V2(configs, wf, energydat)
        Snew = compute_S(e_trial, e_est, branchcut_start, v2, tstep, eloc, nelec)
        Sold = compute_S(e_trial, e_est, branchcut_start, v2old, tstep, elocold, nelec)
        wmult = np.exp(tstep * tdamp * (0.5 * Snew + 0.5 * Sold))
        weights *= wmult
        wavg = np.mean(weights)
        avg = {}
        for k, accumulator in accumulators.items():
            dat = accumulator(configs, wf) if k != ekey[0] else energydat
            for m, res in dat.items():
                avg[k + m] = np.einsum("...i,i...->...", weights, res) / (nconfig * wavg)
        avg["weight"] = wavg
        avg["acceptance"] = np.mean(prob_acceptance)
        avg["tmove_acceptance"] = np.mean(tmove_acceptance)
        df.append(avg)
    return _collect(df), configs, weights


def branch(configs, weights, base_draw=None):
    """Stochastic-comb branching (dmc.py:342-376).  ``base_draw``: the ``np.random.rand()`` of line 361
    when it was drawn ahead of time (``DmcPrefetcher``)."""
    nconfig = configs.configs.shape[0]
    if np.any(weights > 2.0):
        logging.warning("Some weights are larger than 2")
    probability = np.cumsum(weights)
    wtot = probability[-1]
    base = (np.random.rand() if base_draw is None else base_draw) * wtot
    newinds = np.searchsorted(probability, (base + np.linspace(0, wtot, nconfig, endpoint=False)) % wtot)
    unique, counts = np.unique(newinds, return_counts=True)
    configs.resample(newinds)
    weights.fill(wtot / nconfig)
    return configs, weights, {"max branches": np.max(counts), "Number of walkers killed": nconfig - unique.shape[0]}


def estimate_energy(df, ekey):
    en = np.asarray([d[ekey[0] + ekey[1]] for d in df])
    wt = np.asarray([d["weight"] for d in df])
    warmup = int(len(en) / 4)
    return np.average(en[warmup:], weights=wt[warmup:]).real


def rundmc(wf, configs, weights=None, tstep=0.01, nblocks=200, nsteps_per_block=None, blockoffset=0, accumulators=None,
           verbose=False, hdf_file=None, continue_from=None, client=None, npartitions=None, ekey=("energy", "total"),
           vmc_warmup=10, branchcut_start=10, feedback=1.0):
    """Same arguments and return value as ``pyqmc.method.dmc.rundmc`` (dmc.py:412-586); the HDF5
    restart file and the futures client are outside the accelerated path."""
    if hdf_file is not None or continue_from is not None:
        raise NotImplementedError("HDF5 checkpointing is outside the accelerated path; pass hdf_file=None "
                                  "or drive these wave functions with pyqmc.method.dmc.rundmc")
    if client is not None:
        raise NotImplementedError("walker partitions are sharded one process per GPU (pyqmc_b200.parallel)")
    if nsteps_per_block is None:
        nsteps_per_block = max(1, int(0.1 / tstep))
    df, configs = mc.vmc(wf, configs, verbose=verbose, nblocks=vmc_warmup)
    wf.recompute(configs)
    en = accumulators[ekey[0]](configs, wf)[ekey[1]]
    eref = np.mean(en).real
    e_trial = eref
    e_est = eref
    esigma = np.std(en)
    if verbose:
        print("eref start", eref, "esigma", esigma)
    nconfig = configs.configs.shape[0]
    if weights is None:
        weights = np.ones(nconfig)
    df = []
    if blockoffset >= nblocks:
        logging.warning(f"blockoffset {blockoffset} >= nblocks {nblocks}; no steps will be run.")
    prefetch = None
    if nblocks > blockoffset and _device_dmc_path(wf, accumulators, ekey):
        prefetch = DmcPrefetcher(wf, configs, tstep, nsteps_per_block, accumulators[ekey[0]], nblocks - blockoffset)
    for block in range(blockoffset, nblocks):
        df_, configs, weights = dmc_propagate(wf, configs, weights, tstep, branchcut_start * esigma, e_trial=e_trial,
                                              e_est=e_est, nsteps=nsteps_per_block, accumulators=accumulators,
                                              ekey=ekey, variates=prefetch.next() if prefetch else None)
        df_["e_trial"] = e_trial
        df_["e_est"] = e_est
        df_["block"] = block
        df_["esigma"] = esigma
        df_["tstep"] = tstep
        df_["weight_std"] = np.std(weights)
        df_["nsteps_per_block"] = nsteps_per_block
        configs, weights, branch_info = branch(configs, weights, prefetch.branch_draw() if prefetch else None)
        df_.update(branch_info)
        df.append(df_)
        e_est = estimate_energy(df, ekey)
        e_trial = e_est - feedback * np.log(np.mean(weights)).real
        if verbose:
            print("energy", df_[ekey[0] + ekey[1]], "e_trial", e_trial, "e_est", e_est, "sigma(w)", df_["weight_std"])
    df_ret = {k: np.asarray([d[k] for d in df]) for k in df[0].keys()} if len(df) > 0 else {}
    return df_ret, configs, weights
