"""One launch of the Sherman-Morrison kernel at the C4 shape (n = 32, 131072 matrices = 1 GiB of inverses) for
`ncu --set full -k regex:k_sm_`: python profiles/sm_kernel_capture.py   (QMCB_SM_NO_TMA=1 selects k_sm_warp<32>)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyqmc_b200 import _lib  # noqa: E402

lib = _lib.load()
n, nmat = 32, 131072
inv = torch.randn(nmat, n, n, dtype=torch.float64, device="cuda") * 0.1 + torch.eye(n, dtype=torch.float64, device="cuda")
vec = torch.randn(nmat, n, dtype=torch.float64, device="cuda") + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")[n // 2]
ratio = torch.empty(nmat, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    assert lib.qmcb_sm_update_device(n, n // 2, nmat, ctypes.c_void_p(inv.data_ptr()), ctypes.c_void_p(vec.data_ptr()), None,
                                     ctypes.c_void_p(ratio.data_ptr()), None) == 0
torch.cuda.synchronize()
