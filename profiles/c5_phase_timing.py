"""Where a multi-GPU DMC block spends its wall time: torchrun --nproc-per-node N profiles/c5_phase_timing.py
Phases per block on rank 0 (mean over blocks, ms): propagate (recompute + device block + read-back), the statistics
allreduce, and the global branching (weights all-gather, comb, all-to-all of the walkers that change owner)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["PYQMC_B200_DEVICE"] = str(local)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import helpers
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc, parallel

    N, tstep, spb, K = 2048, 0.02, 5, 40
    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    acc = {"energy": pq.EnergyAccumulator(mol)}
    np.random.seed(1000 + rank)
    configs = pq.initial_guess(mol, N)
    df0, configs = pq.vmc(wf, configs, tstep=0.5, nblocks=2, nsteps_per_block=10, accumulators=acc)
    weights = np.ones(N)
    e0 = float(df0["energytotal"][-1])
    src = dmc.dmc_variate_source(wf, configs, tstep, spb, acc["energy"], K + 4)
    t = {"propagate": 0.0, "allreduce": 0.0, "branch": 0.0, "all_gather(+wait for the slowest rank)": 0.0, "all_to_all": 0.0}
    if world > 1:  # time spent inside the two collectives of branch_global (call + the host sync that follows)
        def timed(name, fn):
            def wrapper(*a, **k):
                t0 = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                t[name] += time.perf_counter() - t0
                return r
            return wrapper
        dist.all_gather_into_tensor = timed("all_gather(+wait for the slowest rank)", dist.all_gather_into_tensor)
        dist.all_to_all_single = timed("all_to_all", dist.all_to_all_single)
    for b in range(K + 4):
        if b == 4:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            for k in t:
                t[k] = 0.0
            t_all = time.perf_counter()
        a = time.perf_counter()
        out, configs, weights = dmc.dmc_propagate(wf, configs, weights, tstep, 10.0, e0, e0, nsteps=spb, accumulators=acc,
                                                  variates=src.next())
        b_ = time.perf_counter()
        c = time.perf_counter()
        configs, weights, _ = parallel.branch_global(configs, weights, base_draw=src.branch_draw(), block_avg=out)
        d = time.perf_counter()
        t["propagate"] += b_ - a
        t["allreduce"] += c - b_
        t["branch"] += d - c
    total = time.perf_counter() - t_all
    src.shutdown()
    if rank == 0:
        print({k: round(1e3 * v / K, 3) for k, v in t.items()}, "ms per block; total", round(1e3 * total / K, 3),
              "ms per block;", world, "ranks;", "e2e %.3e walker-steps/s" % (N * world * K * spb / total), flush=True)
    if world > 1:
        dist.destroy_process_group()


main()
