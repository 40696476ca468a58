"""Protocol calls + energy accumulator of a COMPLEX periodic wave function (diamond 2x1x1 at a general twist, 512 walkers)
for `ncu --set full -k regex:k_cx_|k_pbc_mo`: python profiles/cx_kernels_capture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import pyqmc_b200 as pq  # noqa: E402

mol, mf, wf, _ = helpers.make_pair("diamond211_twist", seed=1)
np.random.seed(1)
configs = pq.initial_guess(mol, 512)
acc = pq.EnergyAccumulator(mol, ewald_gmax=10)
for _ in range(2):
    wf.recompute(configs)
    e = 3
    ep = configs.make_irreducible(e, configs.configs[:, e] + 0.2 * np.random.randn(512, 3))
    g, v, saved = wf.gradient_value(e, ep)
    wf.gradient_laplacian(e, ep)
    wf.updateinternals(e, ep, configs, mask=np.abs(v) ** 2 > np.random.rand(512), saved_values=saved)
    en = acc(configs, wf)
    wf.pgradient()
print("mean total energy", en["total"].mean())
