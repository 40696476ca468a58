"""Times the SR product C = A^T B (A, B [N][P], FP64) in its two device variants: python profiles/gemm_microbench.py
Prints ms per launch and TFLOP/s (2 N P^2) next to the FP64 roof measured by qmcb_fp64_peak."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyqmc_b200 import _lib  # noqa: E402

lib = _lib.load()
peak = ctypes.c_double(0.0)
_lib.check(lib.qmcb_fp64_peak(0, ctypes.byref(peak)))
rows = []
only = int(os.environ["GEMM_VARIANT"]) if "GEMM_VARIANT" in os.environ else None
for N, P in ((4096, 256), (4096, 512), (4096, 1024), (16384, 512)):
    rng = np.random.RandomState(1)
    A, B = rng.randn(N, P), rng.randn(N, P)
    C = np.empty((P, P))
    for variant, label in ((0, "fma"), (1, "dmma")):
        if only is not None and variant != only:
            continue
        ms = ctypes.c_double(0.0)
        _lib.check(lib.qmcb_gemm_tn(0, N, P, _lib.dptr(A), _lib.dptr(B), _lib.dptr(C), variant, 20, ctypes.byref(ms)))
        tf = 2.0 * N * P * P / (ms.value * 1e-3) / 1e12
        rows.append({"N": N, "P": P, "variant": label, "ms": ms.value, "tflops": tf, "frac_of_fp64_fma_peak": tf / peak.value})
        print(rows[-1])
print(json.dumps({"fp64_fma_peak_tflops": peak.value, "rows": rows}))
