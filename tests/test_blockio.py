"""Block files (pyqmc_b200/blockio.py): the reference's block schema (hdftools.py:19-53, mc.py:92-99,
dmc.py:379-391) in the dependency-free npz backend, restart semantics of vmc / rundmc."""
import os

import numpy as np
import pytest

from pyqmc_b200 import blockio, coord


def _row(i):
    return {"energytotal": -17.0 + 0.1 * i, "acceptance": 0.5, "block": i, "nconfig": 40, "vec": np.arange(3.0) * i}


def test_npz_store_appends_rows_and_replaces_walkers(tmp_path):
    path = str(tmp_path / "run.hdf5")  # any name: the backend is chosen from the magic bytes
    walkers = coord.PeriodicConfigs(np.random.RandomState(0).randn(4, 2, 3), 3.0 * np.eye(3))
    for i in range(3):
        walkers.configs += 0.01
        with blockio.NpzStore(path, "a") as store:
            store.append_block(_row(i), attrs={"tstep": 0.5}, walkers=walkers)
    with blockio.open_store(path, "r") as store:
        assert isinstance(store, blockio.NpzStore)
        assert store["energytotal"].shape == (3,) and store["vec"].shape == (3, 3)
        assert np.array_equal(store["block"], [0, 1, 2]) and int(store.last("block")) == 2
        assert np.array_equal(store["configs"], walkers.configs) and np.array_equal(store["wrap"], walkers.wrap)
        assert float(store.attrs["tstep"]) == 0.5
        fresh = coord.PeriodicConfigs(np.zeros((4, 2, 3)), 3.0 * np.eye(3))
        blockio.load_walkers(store, fresh)
        assert np.array_equal(fresh.configs, walkers.configs) and np.array_equal(fresh.wrap, walkers.wrap)
    assert not [f for f in os.listdir(tmp_path) if f.startswith(".blockio-")], "temporary file left behind"


def test_read_only_store_refuses_writes_and_missing_file(tmp_path):
    with pytest.raises(FileNotFoundError):
        blockio.NpzStore(str(tmp_path / "absent"), "r")
    path = str(tmp_path / "a")
    with blockio.NpzStore(path, "a") as store:
        store.append_block(_row(0))
    with pytest.raises(IOError):
        blockio.NpzStore(path, "r").append_block(_row(1))


def test_walker_count_follows_the_file(tmp_path):
    path = str(tmp_path / "w")
    with blockio.NpzStore(path, "a") as store:
        store.append_block(_row(0), walkers=coord.OpenConfigs(np.ones((6, 2, 3))), extra_walker_arrays={"weights": np.ones(6)})
    target = coord.OpenConfigs(np.zeros((4, 2, 3)))
    with blockio.open_store(path, "r") as store:
        blockio.load_walkers(store, target)
        assert store["weights"].shape == (6,)
    assert target.configs.shape == (6, 2, 3)


def test_walkers_container_protocol():
    """The operations the reference's drivers call on a configs object (coord.py), open and periodic."""
    rng = np.random.RandomState(1)
    lat = np.array([[3.0, 0.2, 0.0], [0.0, 2.5, 0.1], [0.3, 0.0, 4.0]])
    for lattice in (None, lat):
        c = coord.Walkers(rng.randn(6, 3, 3) * 3, lattice)
        if lattice is not None:
            frac = c.configs @ np.linalg.inv(lat)
            assert frac.min() >= 0 and frac.max() < 1
            assert np.allclose((c.configs + c.wrap @ lat), (c.configs + c.wrap @ lat))
        trial = c.make_irreducible(1, c.configs[:, 1] + 2.5)
        accept = np.array([True, False, True, True, False, False])
        before = c.configs.copy()
        c.move(1, trial, accept)
        assert np.array_equal(c.configs[~accept], before[~accept]) and np.array_equal(c.configs[accept, 1], trial.configs[accept])
        parts = c.split(4)
        assert [len(p.configs) for p in parts] == [2, 2, 1, 1]
        d = c.copy()
        d.join(parts)
        assert np.array_equal(d.configs, c.configs)
        c.resample(np.array([0, 0, 5, 5, 2, 1]))
        assert np.array_equal(c.configs[1], c.configs[0]) and len(c.mask(accept).configs) == 3
        aux = c.make_irreducible(0, rng.randn(6, 5, 3), mask=accept)
        assert aux.configs.shape == (6, 5, 3) and (lattice is None or aux.wrap.shape == (6, 5, 3))
        assert c.electron(2).configs.shape == (6, 3) and c.select_electrons([0, 2]).configs.shape == (6, 2, 3)
