"""VMC driver with the reference's signature (``pyqmc/method/mc.py``).

``initial_guess`` (mc.py:25-73), ``limdrift`` (76-89) and ``vmc`` (176-274) keep the
reference's arguments and output dictionary.  When the wave function is a fused device
``MultiplyWF`` (or a single device factor) and every accumulator is a
``pyqmc_b200.EnergyAccumulator``, a block runs DEVICE-RESIDENT: the random variates of the
whole block are drawn up front from the global legacy ``np.random`` stream in exactly the
order ``vmc_worker`` (mc.py:102-153) and ``eval_ecp`` would consume them, shipped once, and
``qmcb_vmc_block`` executes all sweeps and energy evaluations without host round trips.
Any other wave function / accumulator goes through the generic per-electron loop, which is
the reference's loop verbatim in behaviour (wf protocol calls, host RNG).
"""
import logging
import time

import numpy as np

from . import _lib
from .accumulators import KEYS, EnergyAccumulator, _device_context
from .coord import OpenConfigs, PeriodicConfigs


def initial_guess(mol, nconfig, r=1.0):
    """mc.py:25-73: electrons near atoms proportionally to charge; same RNG consumption."""
    nelec = int(np.sum(mol.nelec))
    epos = np.zeros((nconfig, nelec, 3))
    wts = mol.atom_charges()
    wts = wts / np.sum(wts)
    coords = mol.atom_coords()
    for s in [0, 1]:
        neach = np.array(np.floor(mol.nelec[s] * wts), dtype=int)
        nassigned = int(np.sum(neach))
        totleft = int(mol.nelec[s] - nassigned)
        ind0 = s * mol.nelec[0]
        epos[:, ind0 : ind0 + nassigned, :] = np.repeat(coords, neach, axis=0)
        if totleft > 0:
            inds = np.argpartition(np.random.random((nconfig, len(wts))), totleft, axis=1)[:, :totleft]
            epos[:, ind0 + nassigned : ind0 + mol.nelec[s], :] = coords[inds]
    epos += r * np.random.randn(*epos.shape)
    if hasattr(mol, "a"):
        return PeriodicConfigs(epos, mol.lattice_vectors())
    return OpenConfigs(epos)


def limdrift(g, cutoff=1):
    tot = np.linalg.norm(g, axis=1)
    mask = tot > cutoff
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(mask[:, np.newaxis], cutoff * g / tot[:, np.newaxis], g)


def _device_path(wf, accumulators):
    try:
        _device_context(wf)
    except TypeError:
        return False
    return all(isinstance(a, EnergyAccumulator) for a in accumulators.values()) and len(accumulators) <= 1


class BlockBuffers:
    """Page-locked host buffers for the variates and results of one device-resident block."""

    def __init__(self, nconf, nelec, nsteps, necp, with_energy):
        P = _lib.PinnedArray
        self.key = (nconf, nelec, nsteps, necp, with_energy)
        self._own = [P((nsteps, nelec, nconf, 3)), P((nsteps, nelec, nconf))]
        self.gauss, self.unif = self._own[0].array, self._own[1].array
        self.ecp_u = self.ecp_rot = self.energy = None
        if with_energy:
            self._own += [P((nsteps, nelec, necp, nconf)), P((nsteps, nelec, necp, 3, 3)), P((nsteps, 6, nconf))]
            self.ecp_u, self.ecp_rot, self.energy = (x.array for x in self._own[2:5])
        self._own += [P((nconf, nelec, 3)), P((nsteps, nelec), np.int64)]
        self.newconf, self.nacc = self._own[-2].array, self._own[-1].array

    def variates(self):
        return self.gauss, self.unif, self.ecp_u, self.ecp_rot


def draw_block_variates(nconf, nelec, tstep, nsteps, accumulator, native=True, out=None):
    """All random numbers of one block in the reference's consumption order.

    native=True runs the draws in ``qmcb_rng_vmc_block`` (csrc/legacy_rng.cpp) on the state of the
    global legacy ``np.random`` generator and writes the advanced state back: same numbers, same
    final stream position as the numpy calls below (tests/test_host_cabi.py), several times faster.
    ``out``: optional (gauss, unif, ecp_u, ecp_rot) arrays to fill (e.g. pinned BlockBuffers)."""
    necp = accumulator.necp if accumulator is not None else 0
    if out is None:
        out = (np.empty((nsteps, nelec, nconf, 3)), np.empty((nsteps, nelec, nconf)),
               np.empty((nsteps, nelec, necp, nconf)) if accumulator is not None else None,
               np.empty((nsteps, nelec, necp, 3, 3)) if accumulator is not None else None)
    gauss, unif, ecp_u, ecp_rot = out
    if native and _draw_block_variates_native(nconf, nelec, tstep, nsteps, necp, out):
        return out
    for step in range(nsteps):
        for e in range(nelec):
            gauss[step, e] = np.random.normal(scale=np.sqrt(tstep), size=(nconf, 3))
            unif[step, e] = np.random.rand(nconf)
        if accumulator is not None:
            ecp_u[step], ecp_rot[step] = accumulator.draw_ecp_variates(nconf, nelec)
    return out


def rng_threads():
    """Host threads for the log/sqrt phase of the native generator (QMCB_RNG_THREADS overrides)."""
    import os

    if "QMCB_RNG_THREADS" in os.environ:
        return max(1, int(os.environ["QMCB_RNG_THREADS"]))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    cores = (os.cpu_count() or 1) // max(local_world, 1)
    # leave a core for the sequential stream walk (phase A of the next block) when cores are scarce
    return max(1, min(8, cores - 1 if cores <= 6 else cores))


def _draw_block_variates_native(nconf, nelec, tstep, nsteps, necp, out):
    import ctypes

    state = np.random.get_state()
    if state[0] != "MT19937":
        return False
    gauss, unif, ecp_u, ecp_rot = out
    for a in out:
        if a is not None and not (a.flags["C_CONTIGUOUS"] and a.dtype == np.float64):
            return False
    lib = _lib.load()
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = ctypes.c_int32(int(state[2]))
    has_gauss = ctypes.c_int32(int(state[3]))
    cached = ctypes.c_double(float(state[4]))
    rc = lib.qmcb_rng_vmc_block(key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(pos),
                                ctypes.byref(has_gauss), ctypes.byref(cached), nsteps, nelec, nconf, necp,
                                float(np.sqrt(tstep)), _lib.dptr(gauss), _lib.dptr(unif), _lib.dptr(ecp_u),
                                _lib.dptr(ecp_rot), rng_threads())
    if rc != 0:
        return False
    np.random.set_state(("MT19937", key, pos.value, has_gauss.value, cached.value))
    return True


def _block_buffers(ctx, nconf, nelec, nsteps, accumulator, slot=0):
    """Two cached sets of pinned buffers per context (double buffering for the RNG prefetch)."""
    key = (nconf, nelec, nsteps, accumulator.necp if accumulator is not None else 0, accumulator is not None)
    cache = ctx.__dict__.setdefault("_block_buffers", {})
    if cache.get(slot) is None or cache[slot].key != key:
        cache[slot] = BlockBuffers(*key)
    return cache[slot]


def _recompute_resident(wf, configs):
    """``wf.recompute(configs)`` without the host round trip when ``configs`` are the walkers the device
    already holds: the coordinates the previous device-resident block returned, untouched since (no
    protocol call changed the device state, the host array still equals the returned one).  The recompute
    kernels then run from the resident coordinates (qmcb_recompute_resident); parameters are pushed as
    ``recompute`` would.  False = not applicable, the caller does the ordinary recompute."""
    import os

    if os.environ.get("QMCB_NO_RESIDENT_RECOMPUTE"):
        return False
    try:
        ctx = _device_context(wf)
    except TypeError:
        return False
    res = getattr(ctx, "_resident", None) if ctx is not None else None
    if res is None or ctx.periodic or res[1] != ctx.epoch or res[0].newconf.shape != configs.configs.shape:
        return False
    if not np.array_equal(res[0].newconf, configs.configs):
        return False
    factors = getattr(wf, "wf_factors", None)
    if factors is not None:
        if not getattr(wf, "_fused", False):
            return False
        for f in factors:
            f._push_parameters(ctx)
    else:
        wf._push_parameters(ctx)
    ctx.recompute_resident(wf._which)
    return True


def vmc_block_device(wf, configs, tstep, nsteps, accumulators, variates=None, return_walker_data=False,
                     buffers=None):
    """One device-resident block; equivalent of ``vmc_worker`` (mc.py:102-153).

    ``variates``: pre-drawn (gauss, unif, ecp_u, ecp_rot) (must have been drawn in stream order);
    ``buffers``: BlockBuffers whose variates were already filled (RNG prefetch of ``vmc``)."""
    nconf, nelec, _ = configs.configs.shape
    if not _recompute_resident(wf, configs):
        wf.recompute(configs)
    ctx = _device_context(wf)
    acc_name, accumulator = (next(iter(accumulators.items())) if accumulators else (None, None))
    if accumulator is not None:
        accumulator._attach(wf)
    if buffers is None:
        buffers = _block_buffers(ctx, nconf, nelec, nsteps, accumulator)
        if variates is None:
            draw_block_variates(nconf, nelec, tstep, nsteps, accumulator, out=buffers.variates())
    if variates is None:
        variates = buffers.variates()
    gauss, unif, ecp_u, ecp_rot = variates
    start = time.perf_counter()
    accept = np.empty((nsteps, nelec, nconf), dtype=np.uint8) if return_walker_data else None
    energy = buffers.energy
    nacc = buffers.nacc
    slot = getattr(buffers, "uploaded_slot", None)
    if slot is not None:  # variates already on their way to the device (RNG prefetch thread)
        buffers.uploaded_slot = None
        _lib.check(ctx.lib.qmcb_vmc_block_slot(
            ctx.h, slot, nsteps, float(tstep), 1 if accumulator is not None else 0,
            _lib.dptr(buffers.newconf), _lib.u8ptr(accept), _lib.dptr(energy), None,
            nacc.ctypes.data_as(_lib.c_i64_p)))
    else:
        _lib.check(ctx.lib.qmcb_vmc_block(
            ctx.h, nsteps, float(tstep), 1 if accumulator is not None else 0,
            _lib.dptr(gauss), _lib.dptr(unif), _lib.dptr(ecp_u), _lib.dptr(ecp_rot),
            _lib.dptr(buffers.newconf), _lib.u8ptr(accept), _lib.dptr(energy), None,
            nacc.ctypes.data_as(_lib.c_i64_p)))
    end = time.perf_counter()
    configs.configs[...] = buffers.newconf
    if ctx.periodic:
        configs.wrap[...] = ctx.get_state("wrap", configs.wrap.shape)
    else:  # the device holds exactly these walkers: the next block may recompute from them in place
        ctx._resident = (buffers, ctx.epoch)  # the BlockBuffers object keeps the pinned memory alive
    block_avg = {}
    if accumulator is not None:
        # per-step walker means (one pairwise-summed reduction per (step, key) row, as np.mean of the row),
        # accumulated over the steps in order as the reference loop does (mc.py:139-147)
        means = np.mean(energy[:nsteps], axis=2)
        for i, m in enumerate(KEYS):
            tot = means[0, i] / nsteps
            for step in range(1, nsteps):
                tot += means[step, i] / nsteps
            block_avg[acc_name + m] = tot
    acc = 0.0
    for e in range(nelec):
        acc += (nacc[nsteps - 1, e] / nconf) / nelec
    block_avg["acceptance"] = acc
    block_avg["move time"] = end - start
    block_avg["accumulator time"] = 0.0
    if return_walker_data:
        return block_avg, configs, {"accept": accept.astype(bool), "energy": np.array(energy)}
    return block_avg, configs


def vmc_worker(wf, configs, tstep, nsteps, accumulators):
    """Generic block: per-electron wf protocol calls, as mc.py:102-153."""
    if _device_path(wf, accumulators):
        return vmc_block_device(wf, configs, tstep, nsteps, accumulators)
    nconf, nelec, _ = configs.configs.shape
    block_avg = {}
    wf.recompute(configs)
    for _ in range(nsteps):
        acc = 0.0
        start_move = time.perf_counter()
        for e in range(nelec):
            g, _, _ = wf.gradient_value(e, configs.electron(e))
            grad = limdrift(np.real(g.T))
            gauss = np.random.normal(scale=np.sqrt(tstep), size=(nconf, 3))
            newcoorde = configs.configs[:, e, :] + gauss + grad * tstep
            newcoorde = configs.make_irreducible(e, newcoorde)
            g, new_val, saved = wf.gradient_value(e, newcoorde)
            new_grad = limdrift(np.real(g.T))
            forward = np.sum(gauss**2, axis=1)
            backward = np.sum((gauss + tstep * (grad + new_grad)) ** 2, axis=1)
            t_prob = np.exp(1 / (2 * tstep) * (forward - backward))
            ratio = np.abs(new_val) ** 2 * t_prob
            accept = ratio > np.random.rand(nconf)
            configs.move(e, newcoorde, accept)
            wf.updateinternals(e, newcoorde, configs, mask=accept, saved_values=saved)
            acc += np.mean(accept) / nelec
        end_move = time.perf_counter()
        start_average = time.perf_counter()
        for k, accumulator in accumulators.items():
            dat = accumulator.avg(configs, wf)
            for m, res in dat.items():
                if k + m not in block_avg:
                    block_avg[k + m] = res / nsteps
                else:
                    block_avg[k + m] += res / nsteps
        end_average = time.perf_counter()
        block_avg["acceptance"] = acc
        block_avg["move time"] = end_move - start_move
        block_avg["accumulator time"] = end_average - start_average
    return block_avg, configs


class _VariatePrefetcher:
    """Three-stage host/device pipeline for the device-resident driver.

    Stage 1 (one thread, blocks strictly in order): phase A of the native generator -- the
    sequential walk of the global legacy MT19937 stream for block b (np.random state read before,
    written back after).  Stage 2 (one thread): phase B (log/sqrt of the accepted pairs) and the
    asynchronous host->device copy into one of three device slots.  The caller's thread runs the
    GPU block.  Nothing else in this driver consumes ``np.random``, so the numbers are exactly
    those the reference loop would draw."""

    NSLOT = 3

    def __init__(self, wf, configs, tstep, nsteps, accumulators, nblocks):
        from concurrent.futures import ThreadPoolExecutor

        nconf, nelec, _ = configs.configs.shape
        self.args = (nconf, nelec, tstep, nsteps)
        self.accumulator = next(iter(accumulators.values())) if accumulators else None
        if _device_context(wf) is None:  # context is created lazily by the first recompute
            wf.recompute(configs)
        self.ctx = _device_context(wf)
        self.lib = self.ctx.lib
        self.remaining = nblocks
        self.issued = 0
        self.stage1 = ThreadPoolExecutor(max_workers=1)
        self.stage2 = ThreadPoolExecutor(max_workers=1)
        self.plans = [self.lib.qmcb_rng_plan_create() for _ in range(self.NSLOT)]
        self.queue = []
        for _ in range(self.NSLOT - 1):
            self._submit()

    def _submit(self):
        if self.remaining <= 0:
            return
        import ctypes

        nconf, nelec, tstep, nsteps = self.args
        slot = self.issued % self.NSLOT
        buf = _block_buffers(self.ctx, nconf, nelec, nsteps, self.accumulator, slot=slot)
        plan = ctypes.c_void_p(self.plans[slot])
        self.issued += 1
        self.remaining -= 1
        ctx, lib = self.ctx, self.lib
        necp = self.accumulator.necp if self.accumulator is not None else 0
        g, u, eu, er = buf.variates()
        nthreads = rng_threads()

        def phase_a():
            state = np.random.get_state()
            if state[0] != "MT19937":
                return False
            key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
            pos, has_gauss = ctypes.c_int32(int(state[2])), ctypes.c_int32(int(state[3]))
            cached = ctypes.c_double(float(state[4]))
            rc = lib.qmcb_rng_phase_a(plan, key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(pos),
                                      ctypes.byref(has_gauss), ctypes.byref(cached), nsteps, nelec, nconf, necp,
                                      float(np.sqrt(tstep)), _lib.dptr(g), _lib.dptr(u), _lib.dptr(eu), _lib.dptr(er),
                                      nthreads)
            if rc != 0:
                raise RuntimeError("native RNG failed")
            np.random.set_state(("MT19937", key, pos.value, has_gauss.value, cached.value))
            return True

        fa = self.stage1.submit(phase_a)

        def phase_b():
            if fa.result():
                lib.qmcb_rng_phase_b(plan, nthreads)
            else:  # non-MT19937 global generator: plain numpy draws (still in order: stage 1 is idle)
                draw_block_variates(nconf, nelec, tstep, nsteps, self.accumulator, native=False, out=buf.variates())
            _lib.check(lib.qmcb_vmc_upload(ctx.h, slot, nsteps, nelec, nconf, necp, _lib.dptr(g), _lib.dptr(u),
                                           _lib.dptr(eu), _lib.dptr(er)))
            buf.uploaded_slot = slot
            return buf

        self.queue.append(self.stage2.submit(phase_b))

    def next(self):
        buf = self.queue.pop(0).result()
        self._submit()
        if not self.queue and self.remaining <= 0:
            self.close()
        return buf

    def close(self):
        if self.plans:
            self.stage1.shutdown(wait=True)
            self.stage2.shutdown(wait=True)
            for p in self.plans:
                self.lib.qmcb_rng_plan_destroy(ctypes_void(p))
            self.plans = []


def ctypes_void(p):
    import ctypes

    return ctypes.c_void_p(p)


def vmc_parallel(wf, configs, tstep, nsteps_per_block, accumulators, client, npartitions):
    """mc.py:156-173: walker partitions on a futures client, weighted average of the blocks."""
    config = configs.split(npartitions)
    runs = [client.submit(vmc_worker, wf, conf, tstep, nsteps_per_block, accumulators) for conf in config]
    allresults = list(zip(*[r.result() for r in runs]))
    configs.join(allresults[1])
    confweight = np.array([len(c.configs) for c in config], dtype=float)
    confweight /= np.mean(confweight) * npartitions
    block_avg = {}
    for k in allresults[0][0].keys():
        block_avg[k] = np.sum([res[k] * w for res, w in zip(allresults[0], confweight)], axis=0)
    return block_avg, configs


def vmc(wf, configs, tstep=0.5, nblocks=10, nsteps_per_block=10, nsteps=None, blockoffset=0,
        accumulators=None, verbose=False, hdf_file=None, continue_from=None, client=None, npartitions=None):
    """Same arguments and return value as ``pyqmc.method.mc.vmc`` (mc.py:176-274)."""
    if nsteps is not None:
        nblocks, nsteps_per_block = nsteps, 1
    if accumulators is None:
        accumulators = {}
        if verbose:
            print("WARNING: running VMC with no accumulators")
    if hdf_file is not None or continue_from is not None:
        raise NotImplementedError("HDF5 checkpointing is outside the accelerated path; pass hdf_file=None "
                                  "or drive these wave functions with pyqmc.method.mc.vmc")
    df = []
    if blockoffset >= nblocks:
        logging.warning(f"blockoffset {blockoffset} >= nblocks {nblocks}; no steps will be run.")
    prefetch = None
    if client is None and _device_path(wf, accumulators) and nblocks > blockoffset:
        prefetch = _VariatePrefetcher(wf, configs, tstep, nsteps_per_block, accumulators, nblocks - blockoffset)
    for block in range(blockoffset, nblocks):
        if verbose:
            print("-", end="", flush=True)
        if prefetch is not None:
            block_avg, configs = vmc_block_device(wf, configs, tstep, nsteps_per_block, accumulators,
                                                  buffers=prefetch.next())
        elif client is None:
            block_avg, configs = vmc_worker(wf, configs, tstep, nsteps_per_block, accumulators)
        else:
            block_avg, configs = vmc_parallel(wf, configs, tstep, nsteps_per_block, accumulators, client, npartitions)
        block_avg["block"] = block
        block_avg["nconfig"] = nsteps_per_block * configs.configs.shape[0]
        df.append(block_avg)
    if verbose:
        print("vmc done")
    df_return = {}
    if len(df) > 0:
        for k in df[0].keys():
            df_return[k] = np.asarray([d[k] for d in df])
    return df_return, configs
