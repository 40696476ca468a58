"""csrc/glibc_log.h restates glibc's log() (the function numpy's legacy Gaussian generator calls) operation by
operation; the device generator relies on it.  Host check against this machine's libm, bit for bit."""
import ctypes

import numpy as np


def test_restated_log_equals_libm_bit_for_bit(lib):
    # polar-method arguments (sums of two squares below 1) and the separate branch for x in [1 - 2^-4, 1)
    assert lib.qmcb_glibc_log_mismatches(20_000_000, 2024) == 0
    assert lib.qmcb_glibc_log_mismatches(1 << 16, 12345) == 0


def test_libm_log_is_what_numpy_uses():
    """numpy's np.log on scalars may use its own SIMD loops; the legacy generator calls libm directly.  Pin the
    assumption the device generator rests on: a Gaussian drawn by numpy equals the polar formula with libm's log."""
    libm = ctypes.CDLL("libm.so.6")
    libm.log.restype = ctypes.c_double
    libm.log.argtypes = [ctypes.c_double]
    libm.sqrt.restype = ctypes.c_double
    libm.sqrt.argtypes = [ctypes.c_double]
    for seed in range(50):
        rs = np.random.RandomState(seed)
        rs2 = np.random.RandomState(seed)
        g = rs.standard_normal(2)
        while True:
            x1 = 2.0 * rs2.random_sample() - 1.0
            x2 = 2.0 * rs2.random_sample() - 1.0
            r2 = x1 * x1 + x2 * x2
            if r2 < 1.0 and r2 != 0.0:
                break
        f = libm.sqrt(-2.0 * libm.log(r2) / r2)
        assert g[0] == f * x2 and g[1] == f * x1
