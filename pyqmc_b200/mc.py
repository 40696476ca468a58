"""VMC driver with the reference's signature (``pyqmc/method/mc.py``).

``initial_guess`` (mc.py:25-73) and ``vmc`` (176-274) keep the reference's arguments, output
dictionary and checkpoint layout.  When the wave function is a fused device ``MultiplyWF`` (or a
single device factor) and the accumulator, if any, is a ``pyqmc_b200.EnergyAccumulator``, a block
runs DEVICE-RESIDENT: the random variates of the whole block are drawn up front from the global
legacy ``np.random`` stream in exactly the order ``vmc_worker`` (mc.py:102-153) and ``eval_ecp``
would consume them, shipped once, and ``qmcb_vmc_block`` executes all sweeps and energy evaluations
without host round trips.  The per-electron protocol loop itself is NOT restated here: the reference's
``pyqmc.method.mc.vmc`` drives these objects unchanged (tests/test_gpu_reference_drivers.py) and is
what ``vmc`` delegates to for anything outside the device-resident path.
"""
import logging
import time

import numpy as np

from . import _lib
from .accumulators import KEYS, EnergyAccumulator, _device_context
from .coord import OpenConfigs, PeriodicConfigs


def initial_guess(mol, nconfig, r=1.0):
    """Starting walkers: every electron sits on a "home" atom plus isotropic Gaussian noise of width
    ``r``.  Each spin channel gives atom ``I`` a quota of ``floor(n_s Z_I / sum Z)`` electrons; what is
    left over goes to distinct atoms picked per walker.  Consumes the global legacy stream exactly as
    ``pyqmc.method.mc.initial_guess`` (mc.py:25-73) does -- per spin one ``random((N, natom))`` only if
    electrons are left over, then a single ``randn(N, nelec, 3)`` -- so equal seeds give equal walkers."""
    charges = np.asarray(mol.atom_charges(), dtype=float)
    share = charges / np.sum(charges)
    coords = mol.atom_coords()
    atoms = np.arange(len(share))
    home = []
    for n_s in mol.nelec:
        quota = np.array(np.floor(n_s * share), dtype=int)
        fixed = np.repeat(atoms, quota)
        home.append(np.broadcast_to(fixed, (nconfig, len(fixed))))
        spare = int(n_s) - len(fixed)
        if spare > 0:
            lots = np.random.random((nconfig, len(share)))
            home.append(np.argpartition(lots, spare, axis=1)[:, :spare])
    home = np.concatenate(home, axis=1)
    positions = coords[home]
    positions += r * np.random.randn(*positions.shape)
    if hasattr(mol, "a"):
        return PeriodicConfigs(positions, mol.lattice_vectors())
    return OpenConfigs(positions)


def _device_path(wf, accumulators):
    try:
        _device_context(wf)
    except TypeError:
        return False
    factors = getattr(wf, "wf_factors", [wf])
    # complex wave functions: the query kernels of csrc/cplx.cuh chained on the device (k_cx_chain), host-drawn variates
    # periodic wave functions: single-determinant Slater x JastrowSpin takes the fused two-launch chain, multi-determinant
    # and three-body ones k_pbc_move_general + the update kernels (qmcb_vmc_block_device decides)
    return all(isinstance(a, EnergyAccumulator) for a in accumulators.values()) and len(accumulators) <= 1


class BlockBuffers:
    """Page-locked host buffers for the variates and results of one device-resident block."""

    def __init__(self, nconf, nelec, nsteps, necp, with_energy, rows=6):
        P = _lib.PinnedArray
        self.key = (nconf, nelec, nsteps, necp, with_energy, rows)
        self._own = [P((nsteps, nelec, nconf, 3)), P((nsteps, nelec, nconf))]
        self.gauss, self.unif = self._own[0].array, self._own[1].array
        self.ecp_u = self.ecp_rot = self.energy = None
        if with_energy:
            self._own += [P((nsteps, nelec, necp, nconf)), P((nsteps, nelec, necp, 3, 3)), P((nsteps, rows, nconf))]
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
    # per step and walker: ke, ee, ei, ecp, grad2, total; complex wave functions add Im ecp, Im total (eval_ecp.py:26)
    key = (nconf, nelec, nsteps, accumulator.necp if accumulator is not None else 0, accumulator is not None,
           8 if ctx.cplx else 6)
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
    if res is None or ctx.periodic or ctx.cplx or res[1] != ctx.epoch or res[0].newconf.shape != configs.configs.shape:
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


def _block_averages(buffers, nsteps, nconf, nelec, acc_name, accumulator):
    """Block dictionary entries from the per-walker results a device block left in ``buffers``."""
    block_avg = {}
    if accumulator is not None:
        # per-step walker means (one pairwise-summed reduction per (step, key) row, as np.mean of the row),
        # accumulated over the steps in order as the reference loop does (mc.py:139-147)
        means = np.mean(buffers.energy[:nsteps], axis=2)
        cplx = {}
        if means.shape[1] == 8:  # complex wave function: the ECP term and the total carry wf.dtype (rows 6, 7 = Im);
            # the mean is taken over the complex per-walker row, as accumulator.avg does (its pairwise order differs
            # from that of two real rows)
            en = buffers.energy[:nsteps]
            cplx = {"ecp": np.mean(en[:, 3] + 1j * en[:, 6], axis=1), "total": np.mean(en[:, 5] + 1j * en[:, 7], axis=1)}
        for i, m in enumerate(KEYS):
            col = cplx[m] if m in cplx else means[:, i]
            tot = col[0] / nsteps
            for step in range(1, nsteps):
                tot += col[step] / nsteps
            block_avg[acc_name + m] = tot
    acc = 0.0
    for e in range(nelec):
        acc += (buffers.nacc[nsteps - 1, e] / nconf) / nelec
    block_avg["acceptance"] = acc
    return block_avg


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
    block_avg = _block_averages(buffers, nsteps, nconf, nelec, acc_name, accumulator)
    block_avg["move time"] = end - start
    block_avg["accumulator time"] = 0.0
    if return_walker_data:
        return block_avg, configs, {"accept": accept.astype(bool), "energy": np.array(energy)}
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


_DEVICE_RNG_OK = None


def device_rng_usable():
    """The device generator reproduces numpy's Gaussians only if its restated ``log`` (csrc/glibc_log.h) rounds
    like THIS host's libm -- checked once per process on 2^16 arguments -- and the global generator is the
    legacy MT19937.  ``QMCB_HOST_RNG=1`` forces the host generator (csrc/legacy_rng.cpp)."""
    import os

    global _DEVICE_RNG_OK
    if os.environ.get("QMCB_HOST_RNG"):
        return False
    if _DEVICE_RNG_OK is None:
        _DEVICE_RNG_OK = _lib.load().qmcb_glibc_log_mismatches(1 << 16, 12345) == 0
    return _DEVICE_RNG_OK and np.random.get_state()[0] == "MT19937"


class _DeviceVariates:
    """Variate source of the device-resident driver when the generator itself runs on the GPU
    (csrc/device_rng.cuh): the global legacy ``np.random`` state is handed to the device once, every block's
    draw program is enqueued one block ahead on the copy stream (it overlaps the previous block's kernels), and
    the advanced state is written back to ``np.random`` by ``close()`` -- the stream ends where the reference's
    loop would have left it.  Same interface as ``_VariatePrefetcher``."""

    NSLOT = 3

    def __init__(self, wf, configs, tstep, nsteps, accumulators, nblocks):
        import ctypes

        nconf, nelec, _ = configs.configs.shape
        self.accumulator = next(iter(accumulators.values())) if accumulators else None
        if _device_context(wf) is None:
            wf.recompute(configs)
        self.ctx = _device_context(wf)
        self.lib = self.ctx.lib
        self.shape = (nsteps, nelec, nconf, self.accumulator.necp if self.accumulator is not None else 0)
        self.sigma = float(np.sqrt(tstep))
        self.remaining, self.issued, self.queue, self.open = nblocks, 0, [], True
        state = np.random.get_state()
        key = np.ascontiguousarray(state[1], dtype=np.uint32)
        _lib.check(self.lib.qmcb_devrng_set_state(self.ctx.h, key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                  int(state[2]), int(state[3]), float(state[4])))
        for _ in range(2):
            self._submit()

    def _submit(self):
        if self.remaining <= 0:
            return
        nsteps, nelec, nconf, necp = self.shape
        slot = self.issued % self.NSLOT
        self.issued += 1
        self.remaining -= 1
        _lib.check(self.lib.qmcb_devrng_vmc_block(self.ctx.h, slot, nsteps, nelec, nconf, necp, self.sigma))
        buf = _block_buffers(self.ctx, nconf, nelec, nsteps, self.accumulator, slot=slot)
        buf.uploaded_slot = slot
        self.queue.append(buf)

    def next(self):
        buf = self.queue.pop(0)
        self._submit()
        return buf

    def close(self):
        """Fetches the generator state back into ``np.random`` (also surfaces a generator error)."""
        import ctypes

        if not self.open:
            return
        self.open = False
        key = np.empty(624, dtype=np.uint32)
        pos, has_gauss, cached = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_double(0.0)
        _lib.check(self.lib.qmcb_devrng_get_state(self.ctx.h, key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                  ctypes.byref(pos), ctypes.byref(has_gauss), ctypes.byref(cached)))
        np.random.set_state(("MT19937", key, pos.value, has_gauss.value, cached.value))


def _variate_source(wf, configs, tstep, nsteps, accumulators, nblocks):
    ctx = None
    try:
        ctx = _device_context(wf)
    except TypeError:
        pass
    if device_rng_usable():
        return _DeviceVariates(wf, configs, tstep, nsteps, accumulators, nblocks)
    del ctx
    return _VariatePrefetcher(wf, configs, tstep, nsteps, accumulators, nblocks)


def _pipelined_blocks(wf, configs, tstep, nsteps, accumulators, source, blocks, on_block):
    """The block loop with block b+1 enqueued before block b's results are read (qmcb_vmc_block_slot_begin / _end):
    the device never waits for the host's read-back and bookkeeping.  Needs the device generator (no host thread has
    to produce block b+1's variates) and open boundaries.  ``on_block(block_index, row, configs)`` is called in order."""
    nconf, nelec, _ = configs.configs.shape
    ctx = _device_context(wf)
    acc_name, accumulator = (next(iter(accumulators.items())) if accumulators else (None, None))
    with_energy = 1 if accumulator is not None else 0
    if not _recompute_resident(wf, configs):
        wf.recompute(configs)
    if accumulator is not None:
        accumulator._attach(wf)
    which = wf._which
    i64p = _lib.c_i64_p

    def begin(buf, recompute):
        slot = buf.uploaded_slot
        buf.uploaded_slot = None
        _lib.check(ctx.lib.qmcb_vmc_block_slot_begin(ctx.h, slot, nsteps, float(tstep), with_energy, which if recompute else 0,
                                                     _lib.dptr(buf.newconf), _lib.dptr(buf.energy), buf.nacc.ctypes.data_as(i64p)))
        return slot, buf, time.perf_counter()

    inflight = begin(source.next(), False)
    for k, block in enumerate(blocks):
        following = begin(source.next(), True) if k + 1 < len(blocks) else None
        slot, buf, started = inflight
        _lib.check(ctx.lib.qmcb_vmc_block_slot_end(ctx.h, slot))
        row = _block_averages(buf, nsteps, nconf, nelec, acc_name, accumulator)
        row["move time"] = time.perf_counter() - started
        row["accumulator time"] = 0.0
        configs.configs[...] = buf.newconf
        ctx._resident = (buf, ctx.epoch)
        on_block(block, row, configs)
        inflight = following
    return configs


def _can_pipeline(wf, source):
    import os

    ctx = _device_context(wf)
    return (isinstance(source, _DeviceVariates) and ctx is not None and not ctx.periodic
            and not os.environ.get("QMCB_NO_PIPELINE"))


def _reference_driver():
    """The reference's own driver module, when PyQMC is installed next to this plugin."""
    try:
        import pyqmc.method.mc as refmc
    except ImportError:
        return None
    return refmc


def _restart_point(hdf_file, continue_from):
    """Which file to resume from under the reference's rules (mc.py:225-236): an existing ``hdf_file``
    is continued; ``continue_from`` must exist and excludes an existing ``hdf_file``."""
    import os

    if continue_from is None:
        return hdf_file
    if not os.path.isfile(continue_from):
        raise RuntimeError(f"cannot continue from {continue_from}; the file does not exist!")
    if hdf_file is not None and os.path.isfile(hdf_file):
        raise RuntimeError(f"continue_from is not None but hdf_file={hdf_file} already exists! "
                           f"Delete or rename {hdf_file} and try again.")
    return continue_from


def vmc(wf, configs, tstep=0.5, nblocks=10, nsteps_per_block=10, nsteps=None, blockoffset=0,
        accumulators=None, verbose=False, hdf_file=None, continue_from=None, client=None, npartitions=None):
    """Device-resident VMC with the arguments, block dictionary and checkpoint layout of
    ``pyqmc.method.mc.vmc`` (mc.py:176-274).

    Device wave function + (at most) one ``pyqmc_b200.EnergyAccumulator``: every block is ONE library
    call (``qmcb_vmc_block*``) fed by the three-stage variate pipeline above.  Anything else -- another
    accumulator, a futures ``client``, a non-device wave function -- is not this package's path: it is
    handed to the reference's driver, which runs these objects through the protocol calls."""
    from . import blockio

    accumulators = {} if accumulators is None else accumulators
    if client is not None or not _device_path(wf, accumulators):
        refmc = _reference_driver()
        if refmc is None:
            raise TypeError("pyqmc_b200.vmc runs device-resident blocks (pyqmc_b200 wave function, optional "
                            "pyqmc_b200.EnergyAccumulator, client=None); drive other combinations with "
                            "pyqmc.method.mc.vmc, which accepts these objects unchanged")
        return refmc.vmc(wf, configs, tstep=tstep, nblocks=nblocks, nsteps_per_block=nsteps_per_block,
                         nsteps=nsteps, blockoffset=blockoffset, accumulators=accumulators, verbose=verbose,
                         hdf_file=hdf_file, continue_from=continue_from, client=client, npartitions=npartitions)
    if nsteps is not None:
        nblocks, nsteps_per_block = nsteps, 1
    if verbose and not accumulators:
        print("WARNING: running VMC with no accumulators")
    resume = _restart_point(hdf_file, continue_from)
    if resume is not None and blockio.exists(resume):
        with blockio.open_store(resume, "r") as store:
            if "configs" in store:
                blockoffset = int(store.last("block")) + 1
                blockio.load_walkers(store, configs)
                if verbose:
                    print(f"Restarting calculation {resume} from block {blockoffset}")
    if blockoffset >= nblocks:
        logging.warning(f"blockoffset {blockoffset} >= nblocks {nblocks}; no steps will be run.")
    rows = []
    todo = max(0, nblocks - blockoffset)
    cplx = wf.dtype == complex  # complex blocks draw their variates on the host and take the plain library call
    prefetch = _variate_source(wf, configs, tstep, nsteps_per_block, accumulators, todo) if todo and not cplx else None

    def finish_block(block, row, walkers):
        row["block"] = block
        row["nconfig"] = nsteps_per_block * walkers.configs.shape[0]
        if hdf_file is not None:
            with blockio.open_store(hdf_file, "a") as store:
                store.append_block(row, attrs={"tstep": tstep}, walkers=walkers)
        rows.append(row)

    try:
        if todo and not cplx and _can_pipeline(wf, prefetch):
            configs = _pipelined_blocks(wf, configs, tstep, nsteps_per_block, accumulators, prefetch,
                                        list(range(blockoffset, nblocks)), finish_block)
        else:
            for block in range(blockoffset, nblocks):
                if verbose:
                    print("-", end="", flush=True)
                row, configs = vmc_block_device(wf, configs, tstep, nsteps_per_block, accumulators,
                                                buffers=None if cplx else prefetch.next())
                finish_block(block, row, configs)
    finally:
        if prefetch is not None:
            prefetch.close()
    if verbose:
        print("vmc done")
    return ({k: np.asarray([r_[k] for r_ in rows]) for k in rows[0]} if rows else {}), configs
