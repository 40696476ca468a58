#!/usr/bin/env python
"""bench.py -- walker-steps/s of VMC on synthetic H2O ccECP-cc-pVTZ Slater-Jastrow (BASELINE.json
configs[1]: 4096 walkers per GPU) + HBM roofline of the Sherman-Morrison update kernel.

One "step" = one VMC step of every walker of this rank: a sweep of single-electron
drift-diffusion moves over all 8 electrons plus the local-energy accumulator (ke, ee, ei, ECP)
-- one iteration of the loop at pyqmc/method/mc.py:112-152.

  value   device-timed (CUDA events on the launching stream), random variates and walkers
          already resident in HBM, one library call per block of 10 steps, L2 flushed between timed blocks;
  e2e     the public call pyqmc_b200.vmc(...) with host numpy walkers: host RNG draws in the
          reference's order, H2D of the variates, the device block, D2H of energies + walkers;
  cpu_baseline / --impl reference   the UNMODIFIED reference (oracle/_ref: pyqmc's numpy + numba path,
          staged by oracle/stage_reference.py) driven by its own mc.vmc / dmc.rundmc on all host cores through
          its own parallel mechanism (a ProcessPoolExecutor client, npartitions = cores; precedent
          tests/integration/test_vmc_parallel.py:40-49).  Falls back to the numpy oracle port (kind "port")
          only when the staged copy is absent.

Launch: python bench.py --gpus N --steps K --warmup W   (N>1: under torchrun, one rank per GPU).
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

NWALKERS = 4096
TSTEP = 0.5
SPB = 10  # VMC steps per block (reference default nsteps_per_block)
WORKLOAD = "H2O ccECP-cc-pVTZ-shaped Slater-Jastrow VMC (synthetic basis/MOs), 8 e-, 57 AOs, 4096 walkers/GPU"
# --workload c4 (not the headline line): BASELINE.json configs[3], diamond 2x2x2 supercell, 64 e-, 8 k-points
WORKLOADS = {
    # cpu_*: the bounded sample of the CPU arm (walkers per host core, steps per block, blocks)
    "c2": dict(system="h2o", walkers=4096, text=WORKLOAD, cpu_walkers=512, cpu_steps=200, cpu_blocks=2),
    "c4": dict(system="diamond222", walkers=1024, cpu_walkers=8, cpu_steps=2, cpu_spb=1, cpu_blocks=2,
               metric="walker-steps/sec (VMC, diamond 2x2x2 PBC SJ); Sherman-Morrison HBM GB/s vs roofline",
               text="diamond-C 2x2x2 supercell PBC Slater-Jastrow VMC (synthetic basis/MOs, 8 k-points), 64 e-, "
                    "16 atoms, Ewald + ECP, 1024 walkers/GPU"),
}
WORKLOADS["c3"] = dict(
    system="h2o_cas_3b", walkers=4096, cpu_walkers=32, cpu_steps=1, cpu_spb=2, cpu_blocks=2,
    metric="walker-steps/sec (VMC, H2O CAS(8e,8o) 4900 determinants + 3-body Jastrow); Sherman-Morrison HBM GB/s vs roofline",
    text="H2O ccECP-cc-pVTZ-shaped multi-determinant (full CAS(8e,8o): 70 x 70 = 4900 determinants) x 2-body x 3-body "
         "Jastrow VMC (synthetic basis/MOs/CI coefficients), 4096 walkers/GPU")
WORKLOADS["c5"] = dict(
    system="h2o", walkers=2048, cpu_walkers=128, cpu_steps=20, tstep=0.02, spb=5, cpu_blocks=2,
    metric="walker-steps/sec (DMC with T-moves, H2O cc-pVTZ SJ, tstep 0.02); Sherman-Morrison HBM GB/s vs roofline",
    text="H2O ccECP-cc-pVTZ-shaped Slater-Jastrow DMC (synthetic basis/MOs), tstep 0.02, 5 steps per block, T-moves, "
         "branching every block, 2048 walkers/GPU")
# DRAM bytes per launch of k_sm_tma32 on 131072 matrices from the committed ncu --set full capture
# (profiles/r2_ncu_k_sm_tma32.txt: dram__bytes_read.sum 1.1075 GB + dram__bytes_write.sum 1.0209 GB)
SM32_TRAFFIC_BYTES = 2.128e9
# FP64 flop of one C2 step at 4096 walkers from the committed ncu capture (profiles/r2_ncu_c2_step_raw.csv):
# (2 DFMA + DADD + DMUL) thread instructions per cycle x elapsed cycles, summed over the four step kernels
SWEEP_STEP_FLOP = 6.49e8


E2E_MIN_S = 2.0
NELEC = {"h2o": 8, "h2o_cas_3b": 8, "diamond222": 64}


def workload_config(workload, walkers_per_gpu, world):
    """The `config` object of the JSON line -- identical for the device arm and the reference arm."""
    wl = WORKLOADS[workload]
    cfg = {"workload": wl["text"].replace(f"{wl['walkers']} walkers/GPU", f"{walkers_per_gpu} walkers/GPU"),
           "walkers_per_gpu": walkers_per_gpu, "nelec": NELEC[wl["system"]], "tstep": wl.get("tstep", TSTEP),
           "ecp_threshold": 10}
    if workload == "c5":
        cfg["l2"] = "walker state (2 KB/walker) is L2-resident by design; no flush between blocks"
        cfg["parallelism"] = f"walker-sharded x{world}, one NCCL allreduce of the weighted sums + global branching per block"
    else:
        cfg["l2"] = (f"256 MiB buffer written between timed blocks of {SPB} steps (L2 flush); one library call per block, "
                     "preceded by the per-block wf.recompute of the resident walkers (inside the timed region)")
        cfg["parallelism"] = f"walker-sharded x{world}, one NCCL allreduce of the energy sums per block of {SPB} steps"
    return cfg


def e2e_bytes_per_block(N, ne, necp, spb, nblocks, device_rng):
    """Host<->device bytes of one block of pyqmc_b200.vmc.  In: the walkers (uploaded by the first block only, later
    blocks recompute from the resident walkers) and either the generator state once per call (device generator: the
    variates are produced on the GPU) or the block's variates (host generator: gauss, unif, ECP uniforms + rotations).
    Out: the per-walker energies, the walkers and the acceptance counts of every block, the generator state once."""
    h2d = N * ne * 3 * 8 / nblocks
    d2h = spb * 6 * N * 8 + N * ne * 3 * 8 + spb * ne * 8
    if device_rng:
        h2d += (624 * 4 + 24) / nblocks
        d2h += (624 * 4 + 24) / nblocks
    else:
        h2d += spb * ne * N * (3 + 1) * 8 + spb * ne * necp * (N + 9) * 8
    return h2d, d2h


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        v = d["hbm_gbs"]
        if isinstance(v, dict):  # tolerate {"value": ...} style entries
            v = v.get("value", v.get("burst"))
        return float(v), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, TypeError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores (the reference's own parallel mechanism is a
# futures pool over walker partitions, mc.py:156-173)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, nwalk, nsteps, warm, system = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle import vmc_driver
    from oracle.local_energy import EnergyOracle

    mol, mf, dets = helpers.make_system(system)
    from oracle.jastrow2 import JastrowOracle
    from oracle.product import ProductOracle
    from oracle.slater_det import SlaterOracle

    oj = JastrowOracle.default(mol)
    a0, ac, bc = helpers.jastrow_coefficients(oj.parameters["acoeff"].shape, oj.parameters["bcoeff"].shape, False, 1)
    oj.parameters["acoeff"][:, a0:, :] = ac[:, a0:, :]
    oj.parameters["bcoeff"][1:, :] = bc[1:, :]
    if hasattr(mol, "a"):
        from oracle.pbc import SlaterPbcOracle

        wf = ProductOracle(SlaterPbcOracle(mol, mf), oj)
    elif system.endswith("_3b"):
        from oracle.jastrow3 import Jastrow3Oracle

        oj3 = Jastrow3Oracle.default(mol)
        oj3.parameters["ccoeff"][...] = helpers.three_body_coefficients(oj3.parameters["ccoeff"].shape)
        wf = ProductOracle(SlaterOracle(mol, mf, determinants=dets), oj, oj3)
    else:
        wf = ProductOracle(SlaterOracle(mol, mf, determinants=dets), oj)
    np.random.seed(seed)
    configs = vmc_driver.initial_guess(mol, nwalk)
    acc = {"energy": EnergyOracle(mol)}
    if warm:
        vmc_driver.vmc_worker(wf, configs, TSTEP, 1, acc)
    t0 = time.perf_counter()
    vmc_driver.vmc_worker(wf, configs, TSTEP, nsteps, acc)
    return time.perf_counter() - t0


def _cpu_worker_dmc(args):
    seed, nwalk, nsteps, system = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle import dmc_driver, vmc_driver
    from oracle.jastrow2 import JastrowOracle
    from oracle.local_energy import EnergyOracle
    from oracle.product import ProductOracle
    from oracle.slater_det import SlaterOracle

    mol, mf, dets = helpers.make_system(system)
    oj = JastrowOracle.default(mol)
    a0, ac, bc = helpers.jastrow_coefficients(oj.parameters["acoeff"].shape, oj.parameters["bcoeff"].shape, False, 1)
    oj.parameters["acoeff"][:, a0:, :] = ac[:, a0:, :]
    oj.parameters["bcoeff"][1:, :] = bc[1:, :]
    wf = ProductOracle(SlaterOracle(mol, mf), oj)
    np.random.seed(seed)
    configs = vmc_driver.initial_guess(mol, nwalk)
    t0 = time.perf_counter()
    dmc_driver.dmc_propagate(wf, configs, np.ones(nwalk), WORKLOADS["c5"]["tstep"], 10.0, 20.0, 20.0, nsteps=nsteps,
                             accumulators={"energy": EnergyOracle(mol)})
    return time.perf_counter() - t0


def cpu_arm_dmc(steps, walkers_per_core, cores=None, system="h2o"):
    cores = min(cores or os.cpu_count() or 1, 64)
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker_dmc, [(100 + i, 16, 1, system) for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_cpu_worker_dmc, [(i, walkers_per_core, steps, system) for i in range(cores)])
        wall = time.perf_counter() - t0
    return (walkers_per_core * cores * steps / wall, cores, wall,
            f"{walkers_per_core} walkers/core x {cores} cores x {steps} DMC steps with T-moves (oracle port, numpy)")


def cpu_arm(steps, warmup, walkers_per_core=256, cores=None, system="h2o"):
    cores = cores or os.cpu_count() or 1
    cores = min(cores, 64)
    with mp.get_context("spawn").Pool(cores) as pool:
        if warmup:
            pool.map(_cpu_worker, [(100 + i, min(32, walkers_per_core), 1, False, system) for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, walkers_per_core, steps, False, system) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = walkers_per_core * cores * steps
    return total / wall, cores, wall, f"{walkers_per_core} walkers/core x {cores} cores x {steps} steps (oracle port, numpy)"


def reference_available():
    from oracle import refload

    return refload.available()


def reference_objects(system):
    """The reference's own wave function + energy accumulator for a workload, with the parameters the
    device arm uses (tests/helpers.py seeds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import warnings

    warnings.filterwarnings("ignore")
    import helpers
    import make_golden
    from oracle import refload

    refload.load()
    from pyqmc.observables.accumulators import EnergyAccumulator

    mol, wf = make_golden.build_reference(system)
    ekw = {"ewald_gmax": 200} if hasattr(mol, "a") else {}
    del helpers
    return mol, wf, EnergyAccumulator(mol, **ekw)


def reference_cpu_arm(workload, nblocks, cores=None):
    """walker-steps/s of the reference itself: pyqmc.method.mc.vmc (or dmc.rundmc) with
    evaluate_orbitals_with="numba", client = ProcessPoolExecutor(cores), npartitions = cores."""
    import concurrent.futures

    wl = WORKLOADS[workload]
    cores = min(cores or os.cpu_count() or 1, 64)
    mol, wf, enacc = reference_objects(wl["system"])
    import pyqmc.method.dmc as refdmc
    import pyqmc.method.mc as refmc

    n = wl["cpu_walkers"] * cores
    spb = wl.get("cpu_spb", wl.get("spb", SPB))
    tstep = wl.get("tstep", TSTEP)
    np.random.seed(17)
    configs = refmc.initial_guess(mol, n)
    acc = {"energy": enacc}
    # numba JIT in the parent (one walker partition, protocol calls + accumulator), so that the forked
    # workers inherit the compiled kernels; then one untimed pass through the pool
    small = configs.split(cores)[0]
    refmc.vmc(wf, small, tstep=tstep, nblocks=1, nsteps_per_block=1, accumulators=acc)
    if workload == "c5":
        refdmc.dmc_propagate(wf, small, np.ones(len(small.configs)), tstep, 10.0, 0.0, 0.0, nsteps=1, accumulators=acc)
    with concurrent.futures.ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork")) as client:
        refmc.vmc(wf, configs, tstep=0.5, nblocks=1, nsteps_per_block=1, accumulators=acc, client=client, npartitions=cores)
        t0 = time.perf_counter()
        if workload == "c5":
            refdmc.rundmc(wf, configs, tstep=tstep, nblocks=nblocks, nsteps_per_block=spb, accumulators=acc, vmc_warmup=0,
                          client=client, npartitions=cores)
            what = f"pyqmc.method.dmc.rundmc(tstep={tstep}, nblocks={nblocks}, nsteps_per_block={spb}, T-moves, branching)"
        else:
            refmc.vmc(wf, configs, tstep=tstep, nblocks=nblocks, nsteps_per_block=spb, accumulators=acc, client=client,
                      npartitions=cores)
            what = f"pyqmc.method.mc.vmc(tstep={tstep}, nblocks={nblocks}, nsteps_per_block={spb})"
        wall = time.perf_counter() - t0
    sample = (f"{what} of the unmodified reference (numba orbitals, EnergyAccumulator), {wl['cpu_walkers']} walkers/core x "
              f"{cores} cores via ProcessPoolExecutor(client), npartitions={cores}")
    return n * nblocks * spb / wall, cores, wall, sample


def cpu_reference_or_port(workload, steps, warmup):
    """(value, cores, wall, sample, kind): the staged reference when present, else the oracle port."""
    wl = WORKLOADS[workload]
    if reference_available():
        spb = wl.get("cpu_spb", wl.get("spb", SPB))
        nblocks = max(1, min(steps // spb, wl.get("cpu_blocks", 2)))
        return reference_cpu_arm(workload, nblocks) + ("reference",)
    if workload == "c5":
        return cpu_arm_dmc(max(1, min(steps, wl["cpu_steps"])), wl["cpu_walkers"], system=wl["system"]) + ("port",)
    return cpu_arm(max(1, min(steps, wl["cpu_steps"])), warmup > 0, walkers_per_core=wl["cpu_walkers"],
                   system=wl["system"]) + ("port",)


def cpu_baseline_subprocess(workload, steps):
    """cpu_baseline of the device arm: the reference arm in a fresh interpreter (no CUDA context to fork)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", str(steps),
           "--warmup", "1"]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)["cpu_baseline"]
    raise RuntimeError("reference arm printed no JSON line: " + out.stderr[-2000:])


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML in-process (SM clock, max SM clock, throttle reasons) every few milliseconds from a
    thread, from construction until stop(): the window covers the warm-up, the timed steps and the
    cross-check run, so it holds samples even when the timed region is a few milliseconds long."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=0.004):
        import threading

        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.marks = []
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.period = period_s
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def mark(self):
        """Timestamp (start / end of the timed region) to report how many samples fell inside it."""
        self.marks.append(time.perf_counter())

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._stop.set()
        self.thread.join(timeout=2)
        mhz = [m for _, m in self.samples]
        inside = [m for t, m in self.samples if len(self.marks) >= 2 and self.marks[0] <= t <= self.marks[1]]
        return {"sm_mhz": float(np.median(inside if inside else mhz)) if mhz else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(mhz), "samples_in_timed_region": len(inside),
                "sm_mhz_median_whole_window": float(np.median(mhz)) if mhz else None,
                "window": "NVML polled every 4 ms from the first warm-up step to the end of the e2e runs"}


def sm_roofline(torch, lib, n, nmat, reps=5):
    """HBM GB/s of the Sherman-Morrison update kernel alone (same kernel updateinternals launches)."""
    inv = torch.randn(nmat, n, n, dtype=torch.float64, device="cuda") * 0.1 + torch.eye(n, dtype=torch.float64, device="cuda")
    vec = torch.randn(nmat, n, dtype=torch.float64, device="cuda") + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")[n // 2]
    ratio = torch.empty(nmat, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # the kernel is launched on this torch stream (a real, non-default stream handle) and the
    # events are recorded on the same stream
    ts = torch.cuda.Stream(priority=-1)  # above the library's energy stream (lowest priority), as its own stream is
    assert ts.cuda_stream != 0
    torch.cuda.synchronize()
    times = []
    with torch.cuda.stream(ts):
        for it in range(reps + 2):
            flush.fill_(it & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            rc = lib.qmcb_sm_update_device(n, n // 2, nmat, ctypes.c_void_p(inv.data_ptr()), ctypes.c_void_p(vec.data_ptr()),
                                           None, ctypes.c_void_p(ratio.data_ptr()), ctypes.c_void_p(ts.cuda_stream))
            assert rc == 0, lib.qmcb_last_error()
            e1.record(ts)
            ts.synchronize()
            if it >= 2:
                times.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.mean(times))
    bytes_alg = nmat * 8 * (2 * n * n + n + 1)  # read inv + row, write inv + ratio
    return bytes_alg / t / 1e9, t, bytes_alg


def gpu_arm(args):
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
    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import pyqmc_b200 as pq
    from pyqmc_b200 import _lib, mc

    lib = _lib.load()
    wl = WORKLOADS[args.workload]
    K, W, N = args.steps, args.warmup, (args.walkers or wl["walkers"])
    mol, mf, wf, _ = helpers.make_pair(wl["system"], seed=1)
    acc = pq.EnergyAccumulator(mol)
    np.random.seed(1000 + rank)
    configs = pq.initial_guess(mol, N)
    ne = configs.configs.shape[1]
    # equilibrate with the public driver (also builds every device buffer)
    pq.vmc(wf, configs, tstep=TSTEP, nblocks=1, nsteps_per_block=args.equil, accumulators={"energy": acc})
    ctx = wf._ctx

    # ---------------- device-resident measurement ----------------
    tot = W + K
    gauss, unif, ecp_u, ecp_rot = mc.draw_block_variates(N, ne, TSTEP, tot, acc)
    d_gauss = torch.from_numpy(gauss).cuda()
    d_unif = torch.from_numpy(unif).cuda()
    d_u = torch.from_numpy(ecp_u).cuda()
    d_rot = torch.from_numpy(ecp_rot).cuda()
    d_energy = torch.empty((SPB, 6, N), dtype=torch.float64, device="cuda")
    d_esum = torch.zeros((tot, 6), dtype=torch.float64, device="cuda")
    d_nacc = torch.zeros((tot, ne), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    wf.recompute(configs)
    # All kernels of the timed region are launched on this torch stream (passed to the C ABI as a
    # raw cudaStream_t) and the CUDA events are recorded on the same stream.
    ts = torch.cuda.Stream(priority=-1)
    stream = ts.cuda_stream
    assert stream != 0, "need a real stream handle: 0 would select the library's internal stream"
    torch.cuda.synchronize()
    vp = ctypes.c_void_p

    def chunks(lo, hi):
        """[lo, hi) cut into blocks of SPB steps (the reference's nsteps_per_block; the last one may be shorter)."""
        return [(s, min(SPB, hi - s)) for s in range(lo, hi, SPB)]

    def run_block(s, n):
        """n consecutive VMC steps in ONE library call, as the public driver issues them (a block); a full block
        starts, as every block of the reference's driver does (mc.py:110), with wf.recompute of the resident walkers."""
        if n == SPB and lib.qmcb_recompute_resident_on(ctx.h, wf._which, vp(stream)) != 0:
            raise RuntimeError(lib.qmcb_last_error().decode())
        rc = lib.qmcb_vmc_block_device(ctx.h, n, TSTEP, 1, vp(d_gauss[s].data_ptr()), vp(d_unif[s].data_ptr()),
                                       vp(d_u[s].data_ptr()), vp(d_rot[s].data_ptr()), None, vp(d_energy.data_ptr()),
                                       vp(d_esum[s].data_ptr()), vp(d_nacc[s].data_ptr()), vp(stream))
        if rc != 0:
            raise RuntimeError(lib.qmcb_last_error().decode())
        if world > 1:
            # one allreduce per block: the block's energy sums, over NVLink
            dist.all_reduce(d_esum[s : s + n])

    torch.cuda.set_stream(ts)
    sampler = ClockSampler(local) if rank == 0 else None  # samples the warm-up and the timed steps
    for s, n in chunks(0, W):
        run_block(s, n)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.kernel_launches()
    evs = []
    if sampler:
        sampler.mark()
    wall0 = time.perf_counter()
    for s, n in chunks(W, tot):
        flush.fill_(s & 0xFF)  # evict the walker state from L2 between timed blocks
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        run_block(s, n)
        e1.record(ts)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if sampler:
        sampler.mark()
    torch.cuda.set_stream(torch.cuda.default_stream())
    launches = ctx.kernel_launches() - launches0
    t_dev = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    # cross-check of the event timing: run the same K steps back to back, wall-clock, no flush
    torch.cuda.synchronize()
    tw0 = time.perf_counter()
    with torch.cuda.stream(ts):
        for s, n in chunks(W, tot):
            run_block(s, n)
    torch.cuda.synchronize()
    t_wall_noflush = time.perf_counter() - tw0
    # the same steps as K single-step calls (no step of one call can overlap the next): what round 1 reported
    single = []
    with torch.cuda.stream(ts):
        for s in range(W, tot):
            flush.fill_(s & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            run_block(s, 1)
            e1.record(ts)
            single.append((e0, e1))
    torch.cuda.synchronize()
    t_single = sum(a.elapsed_time(b) for a, b in single) * 1e-3
    tmax = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_dev_max = float(tmax.item())
    value = N * world * K / t_dev_max
    e_mean = float(d_esum[W:tot, 5].sum().item()) / (N * world * K)
    accept = float(d_nacc[W:tot].sum().item()) / (N * ne * K)

    # ---------------- end-to-end through the public API (host buffers) ----------------
    # The sample is independent of --steps: the long call runs >= 50 blocks and >= E2E_MIN_S seconds.  Two calls
    # of different length separate the pipeline fill (first blocks of a call: generator plans, first variates)
    # from the steady state: steady = extra blocks / extra time.
    spb = SPB
    t0 = time.perf_counter()
    df, configs = pq.vmc(wf, configs, tstep=TSTEP, nblocks=6, nsteps_per_block=spb, accumulators={"energy": acc})
    t_block = (time.perf_counter() - t0) / 6  # warm-up of the call shape doubles as the block-time estimate
    nb_long = int(max(50, np.ceil(E2E_MIN_S / max(t_block, 1e-4))))
    nb_short = max(4, nb_long // 4)
    if world > 1:
        nbt = torch.tensor([nb_long], dtype=torch.int64, device="cuda")
        dist.all_reduce(nbt, op=dist.ReduceOp.MAX)
        nb_long = int(nbt.item())
        nb_short = max(4, nb_long // 4)

    def timed_vmc(nb):
        nonlocal configs
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        _, configs = pq.vmc(wf, configs, tstep=TSTEP, nblocks=nb, nsteps_per_block=spb, accumulators={"energy": acc})
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_short = timed_vmc(nb_short)
    t_long = timed_vmc(nb_long)
    nb_e2e, t_e2e = nb_long, t_long
    e2e = N * world * nb_long * spb / t_long
    e2e_steady = N * world * (nb_long - nb_short) * spb / max(t_long - t_short, 1e-9)
    necp = acc.necp
    device_rng = mc.device_rng_usable()
    h2d, d2h = e2e_bytes_per_block(N, ne, necp, spb, nb_e2e, device_rng)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    fp64_peak = ctypes.c_double(0.0)
    lib.qmcb_fp64_peak(local, ctypes.byref(fp64_peak))
    g32, t32, b32 = sm_roofline(torch, lib, 32, 1 << 17)
    g4, t4, b4 = sm_roofline(torch, lib, 4, 1 << 22)
    g4s, t4s, _ = sm_roofline(torch, lib, 4, N)
    out = {
        "metric": wl.get("metric", "walker-steps/sec (VMC, H2O cc-pVTZ SJ); Sherman-Morrison HBM GB/s vs roofline"),
        "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_dev_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, N, world),
        "e2e": {"value": e2e, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d / spb, "d2h_bytes_per_step": d2h / spb,
                "call": f"pyqmc_b200.vmc(nblocks={nb_e2e}, nsteps_per_block={spb}), host walkers in / host walkers + block "
                        f"averages out, bit-exact legacy np.random stream "
                        f"({'continued on the device from np.random.get_state()' if device_rng else 'host generator'}); "
                        f"{t_e2e:.2f} s",
                "rng": "device" if device_rng else "host",
                "e2e_fill_included": e2e, "e2e_steady": e2e_steady,
                "steady_definition": f"({nb_long} - {nb_short}) blocks / (t[{nb_long} blocks] - t[{nb_short} blocks])"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_sm_tma32: Sherman-Morrison row update, n=32 (C4 shape), 131072 matrices, staged by cp.async.bulk",
                     "bound": "hbm", "achieved": g32, "peak": peak, "unit": "GB/s", "frac": g32 / peak,
                     "traffic": SM32_TRAFFIC_BYTES, "traffic_source": "profiles/r2_ncu_k_sm_tma32.txt (ncu --set full, same launch shape)",
                     "peak_source": peak_src, "launch_ms": 1e3 * t32, "algorithmic_bytes": b32,
                     "scope": "the kernel BASELINE.json's metric names, timed alone at the C4 matrix shape (1 GiB of inverses, "
                              "larger than L2): kernel capability.  Inside the C2 step the update is fused into the sweep and "
                              "works on L2-resident 4x4 matrices; the step's own roofline is roofline_step"},
        # the whole VMC step against the same HBM roof (~27.5 KB of algorithmic traffic per
        # walker-step -- 15 KB sweep + ~10 KB energy accumulator); the step is FP64-latency bound, not HBM bound
        # the whole VMC step: against the HBM roof (~27.5 KB of algorithmic traffic per walker-step) it sits at
        # a few per cent because it is not memory bound; against the FP64 roof (measured here: qmcb_fp64_peak) with the
        # FP64 operation count of the committed ncu capture (profiles/r2_ncu_c2_step_raw.csv: DFMA x2 + DADD + DMUL thread
        # instructions of k_vmc_sweep<16,0> + k_ecp_points<4> + k_energy_finalize<8> + k_ecp_prepare per step, 4096 walkers)
        "roofline_step": {"kernel": "k_vmc_sweep<16> + energy kernels (one VMC step)", "bound": "fp64",
                          "achieved": SWEEP_STEP_FLOP * (N / 4096.0) * K / t_dev / 1e12, "peak": fp64_peak.value,
                          "unit": "TFLOP/s", "frac": SWEEP_STEP_FLOP * (N / 4096.0) * K / t_dev / 1e12 / max(fp64_peak.value, 1e-9),
                          "peak_source": "measured in this run (qmcb_fp64_peak: 8 independent DFMA chains per thread)",
                          "flop_per_step_source": "profiles/r2_ncu_c2_step_raw.csv",
                          "hbm_view": {"achieved_GBps": 27.5e3 * N * K / t_dev / 1e9, "frac": 27.5e3 * N * K / t_dev / 1e9 / peak,
                                       "algorithmic_bytes_per_walker_step": 27.5e3},
                          "binding_limit": "dependent FP64 / shared-memory latency of one walker's chain of 8 electron moves: "
                                           "0.20 ms at 1024 walkers, 0.23 ms at 4096, throughput-bound only above ~16 k walkers "
                                           "(profiles/r2_walker_scaling.txt)"},
        "sm_kernel_other_shapes": {
            "n4_4M_matrices": {"achieved_GBps": g4, "frac": g4 / peak, "launch_ms": 1e3 * t4},
            "n4_4096_matrices_C2_shape": {"achieved_GBps": g4s, "frac": g4s / peak, "launch_ms": 1e3 * t4s}},
        "clocks": clocks,
        "check": {"mean_local_energy": e_mean, "acceptance": accept, "wall_s_timed_region_incl_flush": wall,
                  "ms_per_step_as_single_step_calls": 1e3 * t_single / K,
                  "wall_s_same_steps_back_to_back_no_flush": t_wall_noflush, "device_s_timed_steps": t_dev_max},
    }
    if args.workload != "c2":
        out.pop("roofline_step")  # the per-walker-step byte count above is the C2 figure
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline_subprocess(args.workload, K)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gpu_arm_dmc(args):
    """--workload c5: DMC blocks (dmc_propagate + branch) through pyqmc_b200.dmc; one step = T-moves of
    every electron + drift-diffusion sweep + local energy + weight update for every walker."""
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
    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc

    wl = WORKLOADS["c5"]
    N, tstep, spb = (args.walkers or wl["walkers"]), wl["tstep"], wl["spb"]
    K, W = max(1, args.steps // spb), max(1, args.warmup // spb)  # blocks
    mol, mf, wf, _ = helpers.make_pair(wl["system"], seed=1)
    acc = {"energy": pq.EnergyAccumulator(mol)}
    np.random.seed(1000 + rank)
    configs = pq.initial_guess(mol, N)
    df0, configs = pq.vmc(wf, configs, tstep=0.5, nblocks=2, nsteps_per_block=args.equil, accumulators=acc)
    weights = np.ones(N)
    ctx = wf._ctx
    e0 = float(df0["energytotal"][-1])  # trial / estimated energy from the VMC warm-up (rundmc, dmc.py:497-506)

    # the block's variates + the branching draw: generated on the device one block ahead (host thread if unavailable)
    prefetch = dmc.dmc_variate_source(wf, configs, tstep, spb, acc["energy"], W + K)

    e_trial = e0

    def block():
        nonlocal configs, weights, e_trial
        out, configs, weights = dmc.dmc_propagate(wf, configs, weights, tstep, 10.0, e_trial, e0, nsteps=spb,
                                                  accumulators=acc, variates=prefetch.next())
        from pyqmc_b200 import parallel

        # global comb; the block's weighted sums (dmc.py:288-303) ride on its weights all-gather
        configs, weights, info = parallel.branch_global(configs, weights, base_draw=prefetch.branch_draw(), block_avg=out)
        # population control as in rundmc (dmc.py:572): e_trial = e_est - feedback * log(mean weight), feedback = 1
        e_trial = e0 - np.log(info["block_avg"]["weight"])
        return out

    for _ in range(W):
        block()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    for _ in range(K):
        out = block()
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    prefetch.shutdown()
    tt = torch.tensor([t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    e2e = N * world * K * spb / t
    # device share: the same blocks without branching / allreduce, variates ready before each block (generated on the
    # device one block ahead, or drawn by the host thread): recompute + qmcb_dmc_block incl. its H2D/D2H
    src = dmc.dmc_variate_source(wf, configs, tstep, spb, acc["energy"], K + 1, with_branch=False)
    out, configs, weights = dmc.dmc_propagate(wf, configs, weights, tstep, 10.0, e0, e0, nsteps=spb, accumulators=acc,
                                              variates=src.next())
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(K):
        out, configs, weights = dmc.dmc_propagate(wf, configs, weights, tstep, 10.0, e0, e0, nsteps=spb, accumulators=acc,
                                                  variates=src.next())
    torch.cuda.synchronize()
    tdev = time.perf_counter() - t1
    src.shutdown()
    td = torch.tensor([tdev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    value = N * world * K * spb / float(td.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ne, necp = configs.configs.shape[1], acc["energy"].necp
    # H2D per step: the variates when the host generator is used; with the device generator only the walkers and
    # weights go up (per block: dmc_propagate recomputes from host walkers, as the reference does)
    per_step = 0 if isinstance(prefetch, dmc.DeviceDmcVariates) else ne * N * (3 + 1) * 8 + ne * necp * (N + 9) * 8 * 2 + ne * N * 2 * 8
    res = {
        "metric": wl["metric"], "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": K * spb, "warmup": W * spb,
        "ms_per_step": 1e3 * float(td.item()) / (K * spb), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config("c5", N, world),
        "e2e": {"value": e2e, "unit": "walker-steps/s", "h2d_bytes_per_step": per_step + (N * ne * 3 * 8 + N * 8) / spb,
                "d2h_bytes_per_step": (N * ne * 3 * 8 + N * 8) / spb,
                "call": "pyqmc_b200.dmc.dmc_propagate + global branching per block; bit-exact legacy np.random stream ("
                        + ("continued on the device" if isinstance(prefetch, dmc.DeviceDmcVariates) else "host generator") + ")"},
        "gpu_launches": int(launches), "clocks": clocks,
        "check": {"block_energy": float(out["energytotal"]), "acceptance": float(out["acceptance"]),
                  "tmove_acceptance": float(out["tmove_acceptance"]), "weight": float(out["weight"])},
    }
    if world == 1 and not args.no_cpu:
        res["cpu_baseline"] = cpu_baseline_subprocess("c5", K * spb)
    print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    v, cores, wall, sample, kind = cpu_reference_or_port(args.workload, args.steps, args.warmup)
    out = {
        "impl": "reference",
        "metric": wl.get("metric", "walker-steps/sec (VMC, H2O cc-pVTZ SJ); Sherman-Morrison HBM GB/s vs roofline"),
        "value": v, "unit": "walker-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, args.walkers or wl["walkers"], args.gpus),
        "cpu_baseline": {"value": v, "unit": "walker-steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "wall_s": wall},
        "e2e": {"value": v, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--walkers", type=int, default=0)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--equil", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload == "c5":
        gpu_arm_dmc(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
