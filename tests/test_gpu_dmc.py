"""GPU parity of the device-resident DMC propagation (qmcb_dmc_block) against the reference's own
dmc_propagate (golden vectors) and against the oracle restatement of the loop."""
import numpy as np
import pytest

import golden_replay
import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["h2o", "c2", "open", "h2o_md"])
def test_device_resident_dmc_matches_reference_golden(lib, name):
    """dmc.py:123-221 with T-moves: walkers, weights and weighted block averages of the reference run."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = pq.OpenConfigs(data["dmc_configs0"].copy())
    weights = np.ones(len(configs.configs))
    acc = {"energy": pq.EnergyAccumulator(mol)}
    assert dmc._device_dmc_path(wf, acc, ("energy", "total"))
    np.random.seed(41)
    launches0 = None
    out, configs, weights = dmc.dmc_propagate(wf, configs, weights, 0.02, 10.0, 1.5, 1.7, nsteps=3, accumulators=acc)
    assert wf._ctx.kernel_launches() > 0
    golden_replay.check_dmc(data, out, configs, weights)
    del launches0


@pytest.mark.parametrize("name", ["h2o", "he", "hatom", "h2o_md", "h2o_3b", "h2o_md_3b"])
def test_device_resident_dmc_matches_oracle_loop(lib, name):
    """Same seed, more walkers and steps: device block vs the oracle loop over the oracle wave function;
    the RNG stream must end at the same position (the block draws exactly what the loop consumes).
    h2o_md / h2o_3b / h2o_md_3b take the general path (k_vmc_move_coop<16, true> + the update kernels, unfused
    T-moves)."""
    import pyqmc_b200 as pq
    from oracle import dmc_driver
    from oracle.local_energy import EnergyOracle
    from pyqmc_b200 import dmc

    mol, mf, wf, orc = helpers.make_pair(name, seed=1)
    np.random.seed(5)
    configs = pq.initial_guess(mol, 96)
    oc = helpers.to_oracle_walkers(configs)
    w1, w2 = np.ones(96), np.ones(96)
    np.random.seed(6)
    out1, configs, w1 = dmc.dmc_propagate(wf, configs, w1, 0.03, 5.0, -1.0, -1.1, nsteps=4,
                                          accumulators={"energy": pq.EnergyAccumulator(mol)})
    tail1 = np.random.rand()
    np.random.seed(6)
    out2, oc, w2 = dmc_driver.dmc_propagate(orc, oc, w2, 0.03, 5.0, -1.0, -1.1, nsteps=4,
                                            accumulators={"energy": EnergyOracle(mol)})
    tail2 = np.random.rand()
    assert tail1 == tail2
    assert np.abs(configs.configs - oc.configs).max() < 1e-9
    assert helpers.relerr(w1, w2) < 1e-9
    for k in out2:
        assert abs(out1[k] - out2[k]) <= 1e-9 * max(1.0, abs(out2[k])), k


def test_rundmc_runs_blocks_with_branching(lib):
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc

    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    np.random.seed(1)
    configs = pq.initial_guess(mol, 256)
    df, configs, weights = dmc.rundmc(wf, configs, tstep=0.02, nblocks=3, accumulators={"energy": pq.EnergyAccumulator(mol)},
                                      vmc_warmup=2)
    assert df["energytotal"].shape == (3,) and np.all(np.isfinite(df["energytotal"]))
    assert np.all(df["nsteps_per_block"] == 5)
    assert np.allclose(weights, weights[0]) and configs.configs.shape == (256, 8, 3)
    assert 0.5 < df["acceptance"].min() <= 1.0


def test_rundmc_prefetch_reproduces_the_sequential_stream(lib):
    """rundmc with the variate prefetcher (block draws and branching draws taken ahead of time on a
    host thread) vs the plain sequence dmc_propagate -> branch -> ... with the same seed: identical
    walkers, weights and stream position."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc

    def run(prefetch):
        mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
        acc = {"energy": pq.EnergyAccumulator(mol)}
        np.random.seed(9)
        configs = pq.initial_guess(mol, 64)
        weights = np.ones(64)
        pf = dmc.DmcPrefetcher(wf, configs, 0.02, 3, acc["energy"], 3) if prefetch else None
        for b in range(3):
            out, configs, weights = dmc.dmc_propagate(wf, configs, weights, 0.02, 10.0, 20.0, 20.1, nsteps=3, accumulators=acc,
                                                      variates=pf.next() if pf else None)
            configs, weights, info = dmc.branch(configs, weights, pf.branch_draw() if pf else None)
        return configs.configs.copy(), weights.copy(), out["energytotal"], np.random.rand()

    a, b = run(False), run(True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
