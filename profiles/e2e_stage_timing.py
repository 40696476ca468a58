"""Where the wall time of pyqmc_b200.vmc (C2, 4096 walkers, 10 steps per block) goes: durations of the
three pipeline stages (phase A of the host generator, phase B + upload, device block) and of the main
thread's wait for the next variate buffer.  Needs a GPU.  Not a bench: the numbers explain bench.py's e2e."""
import os
import sys
import time
from collections import defaultdict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import pyqmc_b200 as pq  # noqa: E402
from pyqmc_b200 import _lib, mc  # noqa: E402

lib = _lib.load()
T = defaultdict(list)


class Timed:
    def __init__(self, f, name):
        self.f, self.name = f, name

    def __call__(self, *a, **k):
        t = time.perf_counter()
        r = self.f(*a, **k)
        T[self.name].append(time.perf_counter() - t)
        return r


class LibProxy:
    def __init__(self, lib):
        self._lib = lib
        self._wrapped = {n: Timed(getattr(lib, n), n) for n in
                         ("qmcb_rng_phase_a", "qmcb_rng_phase_b", "qmcb_vmc_upload", "qmcb_vmc_block", "qmcb_vmc_block_slot", "qmcb_recompute")
                         if hasattr(lib, n)}

    def __getattr__(self, n):
        w = self.__dict__["_wrapped"].get(n)
        return w if w is not None else getattr(self.__dict__["_lib"], n)


mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
acc = pq.EnergyAccumulator(mol)
np.random.seed(1000)
N = int(os.environ.get("WALKERS", "4096"))
configs = pq.initial_guess(mol, N)
pq.vmc(wf, configs, tstep=0.5, nblocks=3, nsteps_per_block=10, accumulators={"energy": acc})
ctx = wf._ctx
ctx.lib = LibProxy(ctx.lib)
orig_next = mc._VariatePrefetcher.next


def timed_next(self):
    t = time.perf_counter()
    r = orig_next(self)
    T["main: wait for variates (prefetch.next)"].append(time.perf_counter() - t)
    return r


mc._VariatePrefetcher.next = timed_next
mc._recompute_resident = Timed(mc._recompute_resident, "main: _recompute_resident (compare + push parameters + launch)")
acc._attach = Timed(acc._attach, "main: accumulator._attach")
orig_block = mc.vmc_block_device
mc.vmc_block_device = Timed(orig_block, "main: vmc_block_device")
nb = 40
t0 = time.perf_counter()
pq.vmc(wf, configs, tstep=0.5, nblocks=nb, nsteps_per_block=10, accumulators={"energy": acc})
wall = time.perf_counter() - t0
print("wall per block %.3f ms  (%.3g walker-steps/s)" % (wall / nb * 1e3, N * nb * 10 / wall))
for k, v in T.items():
    v = np.array(v[3:]) * 1e3
    print("%-45s n=%3d  median %.3f ms  mean %.3f ms  max %.3f ms" % (k, len(v), np.median(v), v.mean(), v.max()))
