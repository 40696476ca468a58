"""The one dense product of the path -- dpidpj = dp^T (w f dp) of StochasticReconfiguration.avg
(stochastic_reconfiguration.py:110-113) -- in both device variants (FP64-FMA tiles and DMMA mma.sync m8n8k4 with a
split walker range) against numpy, ragged shapes included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("N,P", [(1000, 70), (4096, 129), (37, 5), (513, 64)])
def test_gemm_tn_matches_numpy(lib, variant, N, P):
    import ctypes

    from pyqmc_b200 import _lib

    rng = np.random.RandomState(N + P)
    A, B = rng.randn(N, P), rng.randn(N, P)
    C = np.empty((P, P))
    ms = ctypes.c_double(0.0)
    _lib.check(lib.qmcb_gemm_tn(0, N, P, _lib.dptr(A), _lib.dptr(B), _lib.dptr(C), variant, 1, ctypes.byref(ms)))
    ref = A.T @ B
    assert np.abs(C - ref).max() < 1e-12 * np.abs(ref).max() * np.sqrt(N)
