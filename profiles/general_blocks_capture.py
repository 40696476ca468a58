"""Device-resident blocks of the GENERAL paths added last -- periodic multi-determinant VMC, periodic DMC, open-boundary
multi-determinant + three-body DMC -- at 1024 walkers, with wall-clock rates and (under ncu
`--metrics gpu__time_duration.sum`) the launch list:  python profiles/general_blocks_capture.py [nconf]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import pyqmc_b200 as pq  # noqa: E402
from pyqmc_b200 import dmc, mc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NSTEPS = 4


def vmc_case(name):
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    ekw = {"ewald_gmax": 10} if hasattr(mol, "a") else {}
    acc = {"energy": pq.EnergyAccumulator(mol, **ekw)}
    np.random.seed(1)
    configs = pq.initial_guess(mol, N)
    avg, configs = mc.vmc_block_device(wf, configs, 0.5, NSTEPS, acc)  # warm-up (tables, allocations)
    ctx = wf._ctx
    n0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    avg, configs = mc.vmc_block_device(wf, configs, 0.5, NSTEPS, acc)
    dt = time.perf_counter() - t0
    ne = configs.configs.shape[1]
    print(f"VMC {name:14s} ne={ne:3d} N={N}: {dt / NSTEPS * 1e3:8.3f} ms/step incl. host variates and copies, "
          f"{(ctx.kernel_launches() - n0) / NSTEPS / ne:5.1f} launches per electron move, acceptance {avg['acceptance']:.3f}, "
          f"E {avg['energytotal']:.6f}")


def dmc_case(name):
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    ekw = {"ewald_gmax": 10} if hasattr(mol, "a") else {}
    acc = {"energy": pq.EnergyAccumulator(mol, **ekw)}
    np.random.seed(1)
    configs = pq.initial_guess(mol, N)
    _, configs = mc.vmc_block_device(wf, configs, 0.5, 2, {})
    w = np.ones(N)
    out, configs, w = dmc.dmc_propagate(wf, configs, w, 0.02, 10.0, -10.0, -10.0, nsteps=NSTEPS, accumulators=acc)
    ctx = wf._ctx
    n0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    out, configs, w = dmc.dmc_propagate(wf, configs, w, 0.02, 10.0, -10.0, -10.0, nsteps=NSTEPS, accumulators=acc)
    dt = time.perf_counter() - t0
    ne = configs.configs.shape[1]
    print(f"DMC {name:14s} ne={ne:3d} N={N}: {dt / NSTEPS * 1e3:8.3f} ms/step incl. host variates and copies, "
          f"{(ctx.kernel_launches() - n0) / NSTEPS / ne:5.1f} launches per electron (T-move + move), "
          f"acceptance {out['acceptance']:.3f}, T-move acceptance {out['tmove_acceptance']:.4f}")


for name in ("diamond211", "diamond211_md", "diamond211_3b"):
    vmc_case(name)
for name in ("diamond211", "diamond211_md", "h2o_md_3b"):
    dmc_case(name)
