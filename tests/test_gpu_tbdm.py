"""Two-body density matrix: pyqmc_b200.TBDMAccumulator (device orbitals, qmcb_testvalue / qmcb_updateinternals /
qmcb_testvalue_many) against a golden vector of the reference's TBDMAccumulator (tbdm.py:26-282), up-down sector with
the full index list and up-up sector with a chosen one, and the reference's own accumulator consuming the device wave
function unchanged."""
import os
import sys

import numpy as np
import pytest

import golden_replay
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _setup():
    import pyqmc_b200 as pq

    data = golden_replay.load("h2o")
    gold = golden_replay.load("tbdm_h2o")
    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    configs = pq.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    cu = np.ascontiguousarray(np.asarray(mf.mo_coeff[0])[:, :3])
    cd = np.ascontiguousarray(np.asarray(mf.mo_coeff[1])[:, 1:5])
    return pq, mol, wf, configs, cu, cd, gold


@pytest.mark.parametrize("tag", ["ud", "uu"])
def test_device_tbdm_matches_reference_golden(lib, tag):
    pq, mol, wf, configs, cu, cd, gold = _setup()
    orb, spin, ijkl = ([cu, cd], (0, 1), None) if tag == "ud" else ([cu, cu], (0, 0), gold["ijkl_uu"])
    acc = pq.TBDMAccumulator(mol, orb, spin, nsweeps=2, tstep=0.5, warmup=7, ijkl=ijkl)
    assert acc.keys() == {"value", "norm_a", "norm_b"}
    assert acc.shapes()["value"] == (gold[f"{tag}_value"].shape[1],)
    before = wf.value()
    np.random.seed(71)
    first = acc(configs, wf)
    second = acc.avg(configs, wf)
    for k in ("value", "norm_a", "norm_b"):
        assert helpers.relerr(first[k], gold[f"{tag}_{k}"]) < TOL, k
        assert helpers.relerr(second[k], gold[f"{tag}_avg_{k}"]) < TOL, k
    after = wf.value()  # every moved electron was put back
    assert np.array_equal(before[0], after[0]) and np.abs(before[1] - after[1]).max() < 1e-11


@pytest.mark.skipif(not refload.available(), reason="staged reference (oracle/_ref) absent")
def test_reference_tbdm_accumulator_consumes_device_wf(lib):
    """The reference's TBDMAccumulator, unchanged, over the device wave function: equals its own golden run."""
    pq, mol, wf, configs, cu, cd, gold = _setup()
    refload.load()
    import pyqmc.configurations.coord as coord
    import pyqmc.wf.orbitals
    from pyqmc.observables.tbdm import TBDMAccumulator

    mol.cart = False
    rconfigs = coord.OpenConfigs(configs.configs.copy())
    acc = TBDMAccumulator(mol, [cu, cd], (0, 1), nsweeps=2, tstep=0.5, warmup=7)
    acc.orbitals = pyqmc.wf.orbitals.MoleculeOrbitalEvaluator(mol, [cu, cd], evaluate_orbitals_with="numba")
    np.random.seed(71)
    first = acc(rconfigs, wf)
    for k in ("value", "norm_a", "norm_b"):
        assert helpers.relerr(first[k], gold[f"ud_{k}"]) < TOL, k
