"""GPU parity against the committed golden vectors generated from the reference itself."""
import numpy as np
import pytest

import golden_replay
import helpers

pytestmark = pytest.mark.gpu


def device_vmc(wf, configs, accumulators):
    from pyqmc_b200 import mc

    nb, spb = 2, 3
    rows, accepts = [], []
    for block in range(nb):
        avg, configs, data = mc.vmc_block_device(wf, configs, 0.5, spb, accumulators, return_walker_data=True)
        rows.append(avg)
        accepts.append(data["accept"])
    df = {k: np.asarray([r[k] for r in rows]) for k in rows[0]}
    return df, configs, np.array(accepts)


def check_internal(wf, data):
    sl, ja = wf.wf_factors
    for s in (0, 1):
        assert helpers.relerr(sl._inverse[s], data[f"inverse{s}"]) < 1e-9
        assert np.array_equal(sl._dets[s][0], data[f"dets{s}"][0])
        assert np.abs(sl._dets[s][1] - data[f"dets{s}"][1]).max() < 1e-10
    assert helpers.relerr(ja._a_partial, data["a_partial"]) < 1e-10
    assert helpers.relerr(ja._b_partial, data["b_partial"]) < 1e-10
    pg = ja.pgradient()
    assert helpers.relerr(pg["acoeff"], data["pgrad_wf2acoeff"]) < 1e-10
    assert helpers.relerr(pg["bcoeff"], data["pgrad_wf2bcoeff"]) < 1e-10


@pytest.mark.parametrize("name", ["he", "h2o", "open", "c2", "h2o_md"])
def test_cuda_reproduces_reference_golden(lib, name):
    import pyqmc_b200 as pq

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    assert np.array_equal(wf.parameters["wf2acoeff"], data["acoeff"])
    configs = pq.OpenConfigs(data["configs0"].copy())
    golden_replay.replay(data, wf, configs, lambda: pq.EnergyAccumulator(mol), device_vmc, check_internal)
