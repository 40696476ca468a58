"""One-body density matrix: pyqmc_b200.OBDMAccumulator (device orbitals + qmcb_testvalue_many) against a golden
vector of the reference's OBDMAccumulator (obdm.py:25-214), and the reference's own accumulator consuming the device
wave function's testvalue_many unchanged."""
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
    gold = golden_replay.load("obdm_h2o")
    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    configs = pq.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    coeff = np.ascontiguousarray(np.asarray(mf.mo_coeff[0])[:, :6])
    return pq, mol, wf, configs, coeff, gold


@pytest.mark.parametrize("tag,kw", [("all", {}), ("up", {"spin": 0})])
def test_device_obdm_matches_reference_golden(lib, tag, kw):
    pq, mol, wf, configs, coeff, gold = _setup()
    acc = pq.OBDMAccumulator(mol, coeff, nsweeps=3, tstep=0.5, warmup=25, **kw)
    assert acc.shapes() == {"value": (6, 6), "norm": (6,)} and acc.keys() == {"value", "norm"}
    np.random.seed(61)
    first = acc(configs, wf)
    second = acc.avg(configs, wf)
    assert helpers.relerr(first["value"], gold[f"{tag}_value"]) < TOL
    assert helpers.relerr(first["norm"], gold[f"{tag}_norm"]) < TOL
    assert helpers.relerr(second["value"], gold[f"{tag}_avg_value"]) < TOL
    assert helpers.relerr(second["norm"], gold[f"{tag}_avg_norm"]) < TOL


def test_orbitals_at_points_equal_the_oracle(lib):
    from oracle import gto

    pq, mol, wf, configs, coeff, gold = _setup()
    acc = pq.OBDMAccumulator(mol, coeff, warmup=0)
    acc._ctx = wf._ctx
    pts = np.random.RandomState(3).randn(777, 3) * 1.5
    mine = acc._orbitals(pts)
    ref = gto.BasisTable(mol).eval(0, pts) @ coeff
    assert helpers.relerr(mine, ref) < 1e-12


@pytest.mark.skipif(not refload.available(), reason="staged reference (oracle/_ref) absent")
def test_reference_obdm_accumulator_consumes_device_testvalue_many(lib):
    """The reference's OBDMAccumulator, unchanged, with the device wave function: equals its own golden run."""
    pq, mol, wf, configs, coeff, gold = _setup()
    refload.load()
    import pyqmc.configurations.coord as coord
    import pyqmc.wf.orbitals
    from pyqmc.observables.obdm import OBDMAccumulator

    rconfigs = coord.OpenConfigs(configs.configs.copy())
    acc = OBDMAccumulator(mol, coeff, nsweeps=3, tstep=0.5, warmup=25)
    acc.orbitals = pyqmc.wf.orbitals.MoleculeOrbitalEvaluator(mol, [coeff, coeff], evaluate_orbitals_with="numba")
    np.random.seed(61)
    first = acc(rconfigs, wf)
    assert helpers.relerr(first["value"], gold["all_value"]) < TOL
    assert helpers.relerr(first["norm"], gold["all_norm"]) < TOL
