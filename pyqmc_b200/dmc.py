"""DMC driver with the reference's signatures (``pyqmc/method/dmc.py``): ``dmc_propagate`` (123-221),
``branch`` (342-376) and ``rundmc`` (412-586) keep the reference's arguments, RNG consumption order,
output dictionaries and restart-file layout.

When the wave function is a fused Slater-Jastrow (single- or multi-determinant, with or without a
three-body factor; open boundaries or a periodic cell) and the only accumulator is a ``pyqmc_b200.EnergyAccumulator``, ``dmc_propagate`` runs DEVICE-RESIDENT:
every random variate of the block is drawn up front from the global legacy ``np.random`` stream in
exactly the order the reference loop consumes it, shipped once, and ``qmcb_dmc_block`` executes the
T-moves, drift-diffusion sweeps, local energies and weight updates without host round trips.

The per-electron propagation loop over the wave-function protocol (``propose_drift_diffusion``,
``propose_tmoves``, ``compute_S`` ...) is NOT restated on the host: the reference's own
``pyqmc.method.dmc`` drives these objects unchanged (tests/test_gpu_reference_drivers.py) and is what
``dmc_propagate`` delegates to outside the device-resident path.
"""
import logging

import numpy as np
import scipy.spatial.transform

from . import _lib, mc
from .accumulators import KEYS, EnergyAccumulator, _device_context
from .wf import JASTROW, JASTROW3, SLATER


def _device_dmc_path(wf, accumulators, ekey):
    """Device-resident propagation: fused real Slater x JastrowSpin [x ThreeBodyJastrow] (any determinant expansion,
    open or periodic boundary conditions), one EnergyAccumulator under ``ekey[0]``."""
    try:
        ctx = _device_context(wf)
    except TypeError:
        return False
    if len(accumulators) != 1 or not isinstance(accumulators.get(ekey[0]), EnergyAccumulator) or ekey[1] != "total":
        return False
    if accumulators[ekey[0]].naip is not None:
        # the fused block uses ONE quadrature table for the energy and the T-moves; the reference's T-moves keep the
        # default sizes when naip is passed (accumulators.py:80-81): that combination runs through the protocol calls
        return False
    which = getattr(wf, "_which", 0)
    if which & ~(SLATER | JASTROW | JASTROW3) or not (which & SLATER) or wf.dtype == complex:
        return False
    del ctx
    return True  # open boundaries: sweep kernel or k_vmc_move_coop<16, true>; periodic: k_pbc_move_general<16, true>


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


    def shutdown(self):
        self.pool.shutdown(wait=True)


class _Done:
    def __init__(self, value):
        self.value = value

    def result(self):
        return self.value


class DeviceDmcVariates:
    """``DmcPrefetcher`` with the generator on the GPU (csrc/device_rng.cuh): the global legacy ``np.random`` state is
    handed to the device once, each block's draw program -- block variates, then the one ``rand()`` of ``branch``
    (dmc.py:361) -- is enqueued one block ahead into one of two device slots, and ``shutdown()`` writes the advanced
    state back to ``np.random``.  Same interface as ``DmcPrefetcher``."""

    def __init__(self, wf, configs, tstep, nsteps, accumulator, nblocks, with_branch=True):
        import ctypes

        self.shape = configs.configs.shape[:2]
        self.args = (float(np.sqrt(tstep)), nsteps, accumulator)
        if _device_context(wf) is None:
            wf.recompute(configs)
        self.ctx = _device_context(wf)
        self.remaining, self.issued, self.with_branch = nblocks, 0, with_branch
        self.queue, self.current, self.open = [], None, True
        st = np.random.get_state()
        key = np.ascontiguousarray(st[1], dtype=np.uint32)
        _lib.check(self.ctx.lib.qmcb_devrng_set_state(self.ctx.h, key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                      int(st[2]), int(st[3]), float(st[4])))
        self._submit()

    def _submit(self):
        if self.remaining <= 0:
            return
        sigma, nsteps, accumulator = self.args
        slot = self.issued % 2
        self.issued += 1
        self.remaining -= 1
        _lib.check(self.ctx.lib.qmcb_devrng_dmc_block(self.ctx.h, slot, nsteps, self.shape[1], self.shape[0], accumulator.necp,
                                                      sigma, 1 if accumulator.has_nonlocal_moves() else 0,
                                                      1 if self.with_branch else 0))
        self.queue.append({"device_slot": slot})

    def next(self):
        self.current = self.queue.pop(0)
        self._submit()  # the following block's program: generated while this block runs
        return self.current

    def branch_draw(self):
        return self.current.get("branch") if self.with_branch else None

    def shutdown(self):
        import ctypes

        if not self.open:
            return
        self.open = False
        key = np.empty(624, dtype=np.uint32)
        pos, has_gauss, cached = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_double(0.0)
        _lib.check(self.ctx.lib.qmcb_devrng_get_state(self.ctx.h, key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                      ctypes.byref(pos), ctypes.byref(has_gauss), ctypes.byref(cached)))
        np.random.set_state(("MT19937", key, pos.value, has_gauss.value, cached.value))


def dmc_variate_source(wf, configs, tstep, nsteps, accumulator, nblocks, with_branch=True):
    """Device generator when it reproduces this host's libm (``mc.device_rng_usable``), else the host thread."""
    cls = DeviceDmcVariates if mc.device_rng_usable() else DmcPrefetcher
    return cls(wf, configs, tstep, nsteps, accumulator, nblocks, with_branch)


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
    if "device_slot" in v:  # variates generated on the device (DeviceDmcVariates)
        import ctypes

        branch = ctypes.c_double(0.0)
        _lib.check(ctx.lib.qmcb_dmc_block_slot(
            ctx.h, v["device_slot"], nsteps, float(tstep), float(branchcut_start), float(e_trial), float(e_est), d(w),
            d(newconf), d(wsums), nacc.ctypes.data_as(_lib.c_i64_p), ntacc.ctypes.data_as(_lib.c_i64_p), ctypes.byref(branch)))
        v["branch"] = branch.value
    else:
        _lib.check(ctx.lib.qmcb_dmc_block(
            ctx.h, nsteps, float(tstep), float(branchcut_start), float(e_trial), float(e_est), d(v["gauss"]), d(v["unif"]),
            d(v["ecp_u"]), d(v["ecp_rot"]), d(v["tm_u"]), d(v["tm_rot"]), d(v["tm_sel"]), d(v["tm_acc"]), d(w), d(newconf),
            d(wsums), nacc.ctypes.data_as(_lib.c_i64_p), ntacc.ctypes.data_as(_lib.c_i64_p)))
    ctx.epoch += 1  # the block moved the walkers on the device
    configs.configs[...] = newconf
    if ctx.periodic:
        configs.wrap[...] = ctx.get_state("wrap", configs.wrap.shape)
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


def _reference_dmc():
    try:
        import pyqmc.method.dmc as refdmc
    except ImportError:
        return None
    return refdmc


def dmc_propagate(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps=5, accumulators=None,
                  ekey=("energy", "total"), variates=None):
    """``nsteps`` of DMC propagation without branching (dmc.py:123-221).  ``variates``: pre-drawn random
    numbers of the block (``DmcPrefetcher``)."""
    assert accumulators is not None, "Need an energy accumulator for DMC"
    if _device_dmc_path(wf, accumulators, ekey):
        return dmc_propagate_device(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps, accumulators,
                                    ekey, variates=variates)
    refdmc = _reference_dmc()
    if refdmc is None:
        raise TypeError("pyqmc_b200.dmc_propagate runs device-resident blocks (fused "
                        "real Slater-Jastrow, one pyqmc_b200.EnergyAccumulator); drive other "
                        "combinations with pyqmc.method.dmc, which accepts these objects unchanged")
    return refdmc.dmc_propagate(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps=nsteps,
                                accumulators=accumulators, ekey=ekey)


def comb_indices(weights, offset):
    """Stochastic comb: ``n`` equally spaced teeth, shifted by ``offset`` (a uniform variate in [0, 1)) times the
    total weight and folded back into [0, W), select walkers through the cumulative weights.  Tooth ``i`` may
    land anywhere, so the result is NOT sorted: slot ``i`` of the new population is a copy of walker
    ``result[i]``.  Arithmetic as in ``branch`` (dmc.py:358-366), so equal draws give equal populations."""
    import ctypes

    w = np.ascontiguousarray(weights, dtype=np.float64)
    if len(w) == 0 or not np.all(np.isfinite(w)) or np.any(w < 0):
        return comb_indices_numpy(w, offset)
    picked, total = np.empty(len(w), dtype=np.int64), ctypes.c_double(0.0)
    _lib.check(_lib.load().qmcb_comb_indices(len(w), _lib.dptr(w), float(offset), picked.ctypes.data_as(_lib.c_i64_p),
                                             ctypes.byref(total)))
    return picked, total.value


def comb_indices_numpy(weights, offset):
    """The same comb in numpy, written as the reference writes it: the definition ``qmcb_comb_indices`` (one linear pass
    in native code, csrc/legacy_rng.cpp) is checked against (tests/test_parallel_gloo.py)."""
    ladder = np.cumsum(weights)
    total = ladder[-1]
    teeth = (offset * total + np.linspace(0, total, len(weights), endpoint=False)) % total
    return np.searchsorted(ladder, teeth), total


def branch(configs, weights, base_draw=None):
    """Branching step (dmc.py:342-376): resample by the comb, reset every weight to the mean.
    ``base_draw``: the uniform variate when it was drawn ahead of time (``DmcPrefetcher``)."""
    if np.any(weights > 2.0):
        logging.warning("Some weights are larger than 2")
    offset = np.random.rand() if base_draw is None else base_draw
    picked, total = comb_indices(weights, offset)
    survivors, copies = np.unique(picked, return_counts=True)
    configs.resample(picked)
    weights.fill(total / len(weights))
    return configs, weights, {"max branches": np.max(copies),
                              "Number of walkers killed": len(weights) - len(survivors)}


class _EnergyHistory:
    """Running estimate of the DMC energy: weighted mean of the block energies after dropping the first
    quarter as warm-up (dmc.py:592-603)."""

    def __init__(self, energies=(), weights=()):
        self.energies, self.weights = list(energies), list(weights)

    def add(self, energy, weight):
        self.energies.append(energy)
        self.weights.append(weight)

    def estimate(self):
        skip = int(len(self.energies) / 4)
        return np.average(np.asarray(self.energies)[skip:], weights=np.asarray(self.weights)[skip:]).real


def rundmc(wf, configs, weights=None, tstep=0.01, nblocks=200, nsteps_per_block=None, blockoffset=0, accumulators=None,
           verbose=False, hdf_file=None, continue_from=None, client=None, npartitions=None, ekey=("energy", "total"),
           vmc_warmup=10, branchcut_start=10, feedback=1.0):
    """Same arguments, return value and restart file as ``pyqmc.method.dmc.rundmc`` (dmc.py:412-586).

    A futures ``client`` selects the reference's host-parallel propagation, which is delegated to the
    reference; walkers are sharded one process per GPU instead (``pyqmc_b200.parallel``)."""
    import os

    from . import blockio

    if client is not None:
        refdmc = _reference_dmc()
        if refdmc is None:
            raise TypeError("client= selects the reference's futures-parallel driver; PyQMC is not importable "
                            "(multi-GPU runs shard walkers with pyqmc_b200.parallel)")
        return refdmc.rundmc(wf, configs, weights=weights, tstep=tstep, nblocks=nblocks, nsteps_per_block=nsteps_per_block,
                             blockoffset=blockoffset, accumulators=accumulators, verbose=verbose, hdf_file=hdf_file,
                             continue_from=continue_from, client=client, npartitions=npartitions, ekey=ekey,
                             vmc_warmup=vmc_warmup, branchcut_start=branchcut_start, feedback=feedback)
    if nsteps_per_block is None:
        nsteps_per_block = max(1, int(0.1 / tstep))  # branch every 0.1 time units
    if continue_from is not None and hdf_file is not None and os.path.isfile(hdf_file):
        raise RuntimeError(f"continue_from is set but hdf_file={hdf_file} already exists! "
                           f"Delete or rename {hdf_file} and try again.")
    if continue_from is None and blockio.exists(hdf_file):
        continue_from = hdf_file
    ename = ekey[0] + ekey[1]
    history = _EnergyHistory()
    if continue_from is not None:
        with blockio.open_store(continue_from, "r") as store:
            if "e_trial" not in store:
                raise ValueError("Did not find e_trial in the restart file. This may mean that you are trying to "
                                 "restart from a different version of DMC")
            blockoffset = int(store.last("block")) + 1
            blockio.load_walkers(store, configs)
            weights = np.array(store["weights"])
            e_trial, e_est, esigma = (store.last(k) for k in ("e_trial", "e_est", "esigma"))
            if continue_from == hdf_file:  # the estimate keeps averaging over the blocks already on file
                history = _EnergyHistory(store[ename], store["weight"])
        if verbose:
            print(f"Restarting calculation {continue_from} from block {blockoffset}")
    else:
        _, configs = mc.vmc(wf, configs, verbose=verbose, nblocks=vmc_warmup)
        wf.recompute(configs)
        en = accumulators[ekey[0]](configs, wf)[ekey[1]]
        e_trial = e_est = np.mean(en).real
        esigma = np.std(en)
        if verbose:
            print("eref start", e_trial, "esigma", esigma)
    if weights is None:
        weights = np.ones(configs.configs.shape[0])
    if blockoffset >= nblocks:
        logging.warning(f"blockoffset {blockoffset} >= nblocks {nblocks}; no steps will be run.")
    todo = max(0, nblocks - blockoffset)
    prefetch = None
    if todo and _device_dmc_path(wf, accumulators, ekey):
        prefetch = dmc_variate_source(wf, configs, tstep, nsteps_per_block, accumulators[ekey[0]], todo)
    rows = []
    try:
        for block in range(blockoffset, nblocks):
            row, configs, weights = dmc_propagate(
                wf, configs, weights, tstep, branchcut_start * esigma, e_trial=e_trial, e_est=e_est,
                nsteps=nsteps_per_block, accumulators=accumulators, ekey=ekey,
                variates=prefetch.next() if prefetch else None)
            row.update({"e_trial": e_trial, "e_est": e_est, "block": block, "esigma": esigma, "tstep": tstep,
                        "weight_std": np.std(weights), "nsteps_per_block": nsteps_per_block})
            configs, weights, info = branch(configs, weights, prefetch.branch_draw() if prefetch else None)
            row.update(info)
            rows.append(row)
            if hdf_file is not None:
                with blockio.open_store(hdf_file, "a") as store:
                    store.append_block(row, walkers=configs, extra_walker_arrays={"weights": weights})
            history.add(row[ename], row["weight"])
            e_est = history.estimate()
            e_trial = e_est - feedback * np.log(np.mean(weights)).real
            if verbose:
                print("energy", row[ename], "e_trial", e_trial, "e_est", e_est, "sigma(w)", row["weight_std"])
                print(info)
    finally:
        if prefetch is not None:
            prefetch.shutdown()
    return ({k: np.asarray([r[k] for r in rows]) for k in rows[0]} if rows else {}), configs, weights
