"""hdf_file / continue_from of pyqmc_b200.vmc and rundmc: block rows and walkers land in the block file,
an interrupted run continued from its file reproduces the uninterrupted one (mc.py:225-236, dmc.py:474-511)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_vmc_block_file_and_restart(lib, tmp_path):
    import pyqmc_b200 as pq
    from pyqmc_b200 import blockio

    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    acc = {"energy": pq.EnergyAccumulator(mol)}
    np.random.seed(3)
    start = pq.initial_guess(mol, 64)
    # uninterrupted: 4 blocks
    np.random.seed(5)
    df_full, c_full = pq.vmc(wf, start.copy(), nblocks=4, nsteps_per_block=2, accumulators=acc, hdf_file=str(tmp_path / "full"))
    # interrupted after 2 blocks, then continued from the same file (the RNG stream continues too)
    path = str(tmp_path / "part")
    np.random.seed(5)
    df_a, c_a = pq.vmc(wf, start.copy(), nblocks=2, nsteps_per_block=2, accumulators=acc, hdf_file=path)
    df_b, c_b = pq.vmc(wf, start.copy(), nblocks=4, nsteps_per_block=2, accumulators=acc, hdf_file=path)
    assert list(df_b["block"]) == [2, 3]
    assert np.array_equal(c_b.configs, c_full.configs)
    with blockio.open_store(path, "r") as store:
        assert list(store["block"]) == [0, 1, 2, 3]
        assert np.array_equal(store["energytotal"], df_full["energytotal"])
        assert np.array_equal(store["configs"], c_full.configs)
        assert float(store.attrs["tstep"]) == 0.5
        assert set(df_full) <= set(store.keys())
    with pytest.raises(RuntimeError):
        pq.vmc(wf, start.copy(), nblocks=5, accumulators=acc, hdf_file=path, continue_from=str(tmp_path / "full"))
    with pytest.raises(RuntimeError):
        pq.vmc(wf, start.copy(), nblocks=5, accumulators=acc, continue_from=str(tmp_path / "absent"))


def test_rundmc_block_file_and_restart(lib, tmp_path):
    import pyqmc_b200 as pq
    from pyqmc_b200 import blockio

    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    acc = {"energy": pq.EnergyAccumulator(mol)}
    np.random.seed(3)
    start = pq.initial_guess(mol, 48)
    kw = dict(tstep=0.02, nsteps_per_block=2, accumulators=acc, vmc_warmup=2)
    np.random.seed(7)
    df_full, c_full, w_full = pq.rundmc(wf, start.copy(), nblocks=4, hdf_file=str(tmp_path / "full"), **kw)
    path = str(tmp_path / "part")
    np.random.seed(7)
    pq.rundmc(wf, start.copy(), nblocks=2, hdf_file=path, **kw)
    df_b, c_b, w_b = pq.rundmc(wf, start.copy(), nblocks=4, hdf_file=path, **kw)
    assert list(df_b["block"]) == [2, 3]
    # The reference's restart rule (dmc.py:496-498) resumes with the e_trial / e_est / esigma that were USED by the
    # last stored block, not the ones updated after it, so a continued run is not a bit-copy of an uninterrupted
    # one; what must hold: the stored blocks are the uninterrupted run's, the resumed block starts from the stored
    # walkers, weights and control values, and the file ends up with every block and the final population.
    assert df_b["e_trial"][0] == df_full["e_trial"][1] and df_b["e_est"][0] == df_full["e_est"][1]
    assert df_b["esigma"][0] == df_full["esigma"][1]
    assert np.all(np.isfinite(w_b)) and c_b.configs.shape == c_full.configs.shape
    with blockio.open_store(path, "r") as store:
        for k in ("energytotal", "weight", "e_trial", "e_est", "esigma", "block", "weight_std", "max branches"):
            assert np.array_equal(store[k][:2], df_full[k][:2]), k
            assert np.array_equal(store[k][2:], df_b[k]), k
        assert list(store["block"]) == [0, 1, 2, 3]
        assert np.array_equal(store["weights"], w_b) and np.array_equal(store["configs"], c_b.configs)
    with blockio.open_store(str(tmp_path / "full"), "r") as store:
        assert np.array_equal(store["weights"], w_full) and np.array_equal(store["configs"], c_full.configs)
