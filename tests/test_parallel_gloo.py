"""world_size-2 gloo test (CPU) of the walker sharding and the one-allreduce-per-block statistics:
the combined averages must equal the reference's weighted mean of vmc_parallel (mc.py:166-172)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_block(wf, configs, tstep, nsteps, accumulators):
    """Stands in for the device block: deterministic 'averages' that depend on the walkers."""
    c = configs.configs
    avg = {"energytotal": float(np.mean(c[:, 0, 0])), "energyke": float(np.mean(c[:, 0, 1] ** 2)),
           "acceptance": float(np.mean(c[:, 0, 2] > 0)), "move time": 1.0, "accumulator time": 0.0}
    configs.configs = c + 1.0
    return avg, configs


def _worker(rank, world, port, nwalk, out_dir):
    import torch.distributed as dist

    from pyqmc_b200 import parallel
    from pyqmc_b200.coord import OpenConfigs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    configs = OpenConfigs(rng.randn(nwalk, 2, 3))
    df, local = parallel.vmc_distributed(None, configs, nblocks=3, nsteps_per_block=2, block_fn=_fake_block)
    allc = parallel.gather_configs(local)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), configs=allc, local=local.configs, **df)
    dist.destroy_process_group()


@pytest.mark.parametrize("nwalk", [10, 11])
def test_two_rank_block_statistics(tmp_path, nwalk):
    import torch.multiprocessing as mp

    from pyqmc_b200.coord import OpenConfigs

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nwalk, str(tmp_path)), nprocs=world, join=True)
    res = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]
    # serial reference: the vmc_parallel formula on the same partitions
    rng = np.random.RandomState(0)
    configs = OpenConfigs(rng.randn(nwalk, 2, 3))
    parts = configs.split(world)
    assert [len(p.configs) for p in parts] == [len(a) for a in np.array_split(np.arange(nwalk), world)]
    expect = {k: [] for k in ("energytotal", "energyke", "acceptance")}
    for block in range(3):
        outs = []
        for i in range(world):
            avg, parts[i] = _fake_block(None, parts[i], 0.5, 2, {})
            outs.append(avg)
        w = np.array([len(p.configs) for p in parts], dtype=float)
        w /= np.mean(w) * world
        for k in expect:
            expect[k].append(np.sum([o[k] * wi for o, wi in zip(outs, w)]))
    for r in range(world):
        for k in expect:
            assert np.allclose(res[r][k], expect[k], rtol=1e-14, atol=1e-14), k
        assert np.array_equal(res[r]["nconfig"], [2 * nwalk] * 3)
        assert np.array_equal(res[r]["block"], [0, 1, 2])
        assert np.allclose(res[r]["configs"], np.concatenate([p.configs for p in parts]))
    assert len(res[0]["local"]) + len(res[1]["local"]) == nwalk


def test_single_process_path_needs_no_process_group():
    from pyqmc_b200 import parallel

    avg, total = parallel.allreduce_block({"energytotal": 2.0, "acceptance": 0.5, "block": 3}, 7)
    assert total == 7 and avg == {"acceptance": 0.5, "energytotal": 2.0}


def _dmc_worker(rank, world, port, nwalk, out_dir):
    import torch.distributed as dist

    from pyqmc_b200 import parallel
    from pyqmc_b200.coord import OpenConfigs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(1)
    allc, allw = rng.randn(nwalk, 2, 3), 0.5 + rng.rand(nwalk)
    idx = np.array_split(np.arange(nwalk), world)[rank]
    local, w = OpenConfigs(allc[idx].copy()), allw[idx].copy()
    block = {"energytotal": float(np.average(allc[idx, 0, 0], weights=w)), "weight": float(np.mean(w)), "acceptance": 0.9}
    glob, total = parallel.allreduce_dmc_block(block, len(idx))
    np.random.seed(7)  # only rank 0's draw is used
    local, w, info = parallel.branch_global(local, w, block_avg=block)  # statistics ride on the weights all-gather
    fused = info["block_avg"]
    assert np.isclose(fused["energytotal"], glob["energytotal"], rtol=1e-14) and np.isclose(fused["weight"], glob["weight"], rtol=1e-14)
    first = dict(configs=local.configs.copy(), weights=w.copy())
    # second round on the returned container: the rank layout it carries replaces the counts all-reduce
    assert getattr(local, "_rank_layout", None) is not None
    w2 = w * (0.5 + np.random.RandomState(100 + rank).rand(len(w)))
    local, w2b, _ = parallel.branch_global(local, w2, base_draw=0.37)
    np.savez(os.path.join(out_dir, f"dmc{rank}.npz"), configs=first["configs"], weights=first["weights"],
             energy=glob["energytotal"], weight=glob["weight"], total=total, killed=info["Number of walkers killed"],
             w2=w2, configs2=local.configs, weights2=w2b)
    dist.destroy_process_group()


@pytest.mark.parametrize("nwalk", [12, 13])
def test_two_rank_dmc_statistics_and_global_branching(tmp_path, nwalk):
    """allreduce_dmc_block == the weighting of dmc_propagate_parallel (dmc.py:238-303); branch_global ==
    join -> branch (dmc.py:342-376) -> split with the same comb offset."""
    import torch.multiprocessing as mp

    from pyqmc_b200 import dmc
    from pyqmc_b200.coord import OpenConfigs

    world = 2
    mp.spawn(_dmc_worker, args=(world, _free_port(), nwalk, str(tmp_path)), nprocs=world, join=True)
    res = [dict(np.load(tmp_path / f"dmc{r}.npz")) for r in range(world)]
    rng = np.random.RandomState(1)
    allc, allw = rng.randn(nwalk, 2, 3), 0.5 + rng.rand(nwalk)
    for r in range(world):
        assert np.isclose(res[r]["energy"], np.average(allc[:, 0, 0], weights=allw), rtol=1e-13)
        assert np.isclose(res[r]["weight"], np.mean(allw), rtol=1e-13)
        assert res[r]["total"] == nwalk
    serial = OpenConfigs(allc.copy())
    np.random.seed(7)
    serial, w, info = dmc.branch(serial, allw.copy())
    got = np.concatenate([res[r]["configs"] for r in range(world)], axis=0)
    assert np.array_equal(got, serial.configs)
    assert np.allclose(np.concatenate([res[r]["weights"] for r in range(world)]), w)
    assert res[0]["killed"] == info["Number of walkers killed"]
    # second round (cached rank layout, explicit comb offset)
    w2 = np.concatenate([res[r]["w2"] for r in range(world)])
    serial2, w2s, _ = dmc.branch(serial, w2.copy(), base_draw=0.37)
    assert np.array_equal(np.concatenate([res[r]["configs2"] for r in range(world)], axis=0), serial2.configs)
    assert np.allclose(np.concatenate([res[r]["weights2"] for r in range(world)]), w2s)


def _dmc_periodic_worker(rank, world, port, nwalk, out_dir):
    import torch.distributed as dist

    from pyqmc_b200 import parallel
    from pyqmc_b200.coord import PeriodicConfigs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allc, allwrap, allw, lat = _periodic_population(nwalk)
    idx = np.array_split(np.arange(nwalk), world)[rank]
    local = PeriodicConfigs(allc[idx].copy(), lat, wrap=allwrap[idx].copy())
    np.random.seed(7)
    local, w, info = parallel.branch_global(local, allw[idx].copy())
    np.savez(os.path.join(out_dir, f"pdmc{rank}.npz"), configs=local.configs, wrap=local.wrap, weights=w)
    dist.destroy_process_group()


def _periodic_population(nwalk):
    rng = np.random.RandomState(3)
    lat = np.array([[4.0, 0.0, 0.0], [0.5, 5.0, 0.0], [0.0, 0.3, 6.0]])
    frac = rng.rand(nwalk, 3, 3)
    return frac @ lat, rng.randint(-2, 3, size=(nwalk, 3, 3)).astype(float), 0.5 + rng.rand(nwalk), lat


@pytest.mark.parametrize("nwalk", [11, 16])
def test_two_rank_global_branching_moves_wrap_vectors_with_the_walkers(tmp_path, nwalk):
    """Periodic walkers (device-resident DMC on solids shards them like any other): the all-to-all of branch_global
    carries the wrap vectors with the coordinates; result == serial branch (dmc.py:342-376) on the joined population."""
    import torch.multiprocessing as mp

    from pyqmc_b200 import dmc
    from pyqmc_b200.coord import PeriodicConfigs

    world = 2
    mp.spawn(_dmc_periodic_worker, args=(world, _free_port(), nwalk, str(tmp_path)), nprocs=world, join=True)
    res = [dict(np.load(tmp_path / f"pdmc{r}.npz")) for r in range(world)]
    allc, allwrap, allw, lat = _periodic_population(nwalk)
    serial = PeriodicConfigs(allc.copy(), lat, wrap=allwrap.copy())
    np.random.seed(7)
    serial, w, _ = dmc.branch(serial, allw.copy())
    assert np.array_equal(np.concatenate([res[r]["configs"] for r in range(world)], axis=0), serial.configs)
    assert np.array_equal(np.concatenate([res[r]["wrap"] for r in range(world)], axis=0), serial.wrap)
    assert np.abs(serial.wrap).max() > 0
    assert np.allclose(np.concatenate([res[r]["weights"] for r in range(world)]), w)


def test_native_comb_equals_the_numpy_definition():
    """qmcb_comb_indices (one linear pass, native) == cumsum / linspace / mod / searchsorted as the reference's branch
    writes them (dmc.py:358-366): every offset class (no wrap, wrap, offset 0), equal weights (ties in the ladder),
    zero weights, one walker."""
    from pyqmc_b200 import dmc

    rng = np.random.RandomState(3)
    cases = [(rng.rand(n) * (1 + 3 * (i % 3)), rng.rand()) for i, n in enumerate([1, 2, 7, 64, 1000, 4096, 16384])]
    cases += [(np.ones(100), 0.0), (np.ones(100), 0.5), (np.ones(64), 0.999999), (np.r_[0.0, 1.0, 0.0, 2.0, 0.0], 0.3),
              (rng.rand(513), 0.0)]
    for w, off in cases:
        got, tot = dmc.comb_indices(w, off)
        ref, rtot = dmc.comb_indices_numpy(w, off)
        assert tot == rtot
        assert np.array_equal(got, ref), (len(w), off)


def test_block_allreduce_packs_complex_averages():
    """Complex wave functions (device-resident blocks since round 2) return complex ECP / total block averages: the
    one-collective-per-block vector carries real and imaginary parts in separate slots."""
    from pyqmc_b200 import parallel

    block = {"energytotal": 2.0 + 0.5j, "energyecp": -1.0 - 0.25j, "energyke": 1.5, "acceptance": 0.5, "block": 3}
    keys, vec = parallel.pack_block(block, 7)
    assert vec.dtype == np.float64 and len(vec) == 1 + 4 + 2
    out, total = parallel.unpack_block(keys, 2 * vec)  # two ranks with equal shards
    assert total == 14
    assert out["energytotal"] == 2.0 + 0.5j and out["energyecp"] == -1.0 - 0.25j
    assert out["energyke"] == 1.5 and not np.iscomplexobj(out["energyke"])
