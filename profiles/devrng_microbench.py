"""Throughput of the device MT19937 generator alone (csrc/device_rng.cuh: k_mt_generate): one uniform draw of
40 M doubles = 80 M words = 128 k state blocks, wall clock around set_state .. get_state, per generator variant
(QMCB_MT_MODE: 0 = one barrier per block, every word from the old block; 1 = twists staged in shared memory, two barriers).
Measured on B200: per-thread stores vs one cp.async.bulk per block vs no stores at all differ by < 10 % (the single CTA is
issue-bound, not store-bound), so plain stores are used."""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import numpy as np
    import torch

    from pyqmc_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.qmcb_create(0, ctypes.byref(h)) == 0
    n = 40_000_000
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    kind = np.array([0], dtype=np.int32)
    count = np.array([n], dtype=np.int64)
    dst = np.array([out.data_ptr()], dtype=np.uint64)
    scale = np.array([1.0])
    U32P = ctypes.POINTER(ctypes.c_uint32)
    for rep in range(3):
        np.random.seed(rep)
        st = np.random.get_state()
        key = np.ascontiguousarray(st[1], dtype=np.uint32)
        assert lib.qmcb_devrng_set_state(h, key.ctypes.data_as(U32P), int(st[2]), int(st[3]), float(st[4])) == 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert lib.qmcb_devrng_program(h, 1, kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                       count.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                       dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                       scale.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
        k2 = np.empty(624, dtype=np.uint32)
        pos, has, cached = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        rc = lib.qmcb_devrng_get_state(h, k2.ctypes.data_as(U32P), ctypes.byref(pos), ctypes.byref(has), ctypes.byref(cached))
        dt = time.perf_counter() - t0
        tim = np.zeros(3, dtype=np.int64)
        lib.qmcb_devrng_generator_timing(h, tim.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        print(f"   last generator launch: {tim[2]} blocks, {tim[0] / max(tim[2], 1):.0f} SM cycles/block, "
              f"{tim[1] / max(tim[2], 1):.0f} ns/block -> SM clock {tim[0] / max(tim[1], 1) * 1e3:.0f} MHz")
        ok = rc == 0 and np.array_equal(out.cpu().numpy()[:1000], np.random.random(size=n)[:1000])
        print(f"mode {os.environ.get('QMCB_MT_MODE', 'default')} rep {rep}: {dt * 1e3:.2f} ms for {2 * n / 624:.0f} blocks "
              f"= {dt / (2 * n / 624) * 1e9:.0f} ns/block, values {'match numpy' if ok else 'DIFFER'}")
    lib.qmcb_destroy(h)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for mode in ("0", "1"):
            subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, QMCB_MT_MODE=mode))
