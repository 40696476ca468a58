"""Host-side timing of the legacy-stream generator for one C2 block (10 steps x 8 electrons x 4096 walkers,
3 ECP atoms): phase A (sequential stream walk + producer thread) and phase B (log/sqrt, threaded) separately,
the one-call form at several thread counts, and numpy itself.  CPU only."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyqmc_b200 import _lib, mc  # noqa: E402

print("cpus", os.cpu_count())
np.random.seed(1)
nsteps, ne, N, necp = 10, 8, 4096, 3
out = (np.empty((nsteps, ne, N, 3)), np.empty((nsteps, ne, N)), np.empty((nsteps, ne, necp, N)), np.empty((nsteps, ne, necp, 3, 3)))
lib = _lib.load()
lib.qmcb_rng_plan_create.restype = ctypes.c_void_p
plan = ctypes.c_void_p(lib.qmcb_rng_plan_create())
for thr in (1, 2, 4, 8):
    ta, tb = [], []
    for rep in range(8):
        state = np.random.get_state()
        key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
        pos, hg, cg = ctypes.c_int32(int(state[2])), ctypes.c_int32(int(state[3])), ctypes.c_double(float(state[4]))
        t0 = time.perf_counter()
        rc = lib.qmcb_rng_phase_a(plan, key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(pos), ctypes.byref(hg),
                                  ctypes.byref(cg), nsteps, ne, N, necp, 0.7, *[_lib.dptr(a) for a in out], thr)
        t1 = time.perf_counter()
        lib.qmcb_rng_phase_b(plan, thr)
        t2 = time.perf_counter()
        assert rc == 0
        np.random.set_state(("MT19937", key, pos.value, hg.value, cg.value))
        ta.append(t1 - t0)
        tb.append(t2 - t1)
    print("%d threads: phase A min %.3f med %.3f | phase B min %.3f med %.3f ms per step" % (
        thr, min(ta) / nsteps * 1e3, sorted(ta)[4] / nsteps * 1e3, min(tb) / nsteps * 1e3, sorted(tb)[4] / nsteps * 1e3))
for thr in (1, 8):
    os.environ["QMCB_RNG_THREADS"] = str(thr)
    ts = []
    for _ in range(6):
        t = time.perf_counter()
        mc._draw_block_variates_native(N, ne, 0.5, nsteps, necp, out)
        ts.append(time.perf_counter() - t)
    print("one call, %d threads: min %.3f ms per step" % (thr, min(ts) / nsteps * 1e3))
t = time.perf_counter()
for step in range(nsteps):
    for e in range(ne):
        np.random.normal(scale=0.7, size=(N, 3))
        np.random.rand(N)
    np.random.rand(ne, necp, N)
print("numpy: %.3f ms per step" % ((time.perf_counter() - t) / nsteps * 1e3))
