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
    sl, ja = wf.wf_factors[:2]
    if len(wf.wf_factors) > 2:
        assert helpers.relerr(wf.wf_factors[2].P_i, data["P_i"]) < 1e-10
        assert helpers.relerr(wf.wf_factors[2].a_values, data["a3_values"]) < 1e-10
    for s in (0, 1):
        assert helpers.relerr(sl._inverse[s], data[f"inverse{s}"]) < 1e-9
        assert np.array_equal(sl._dets[s][0], data[f"dets{s}"][0])
        assert np.abs(sl._dets[s][1] - data[f"dets{s}"][1]).max() < 1e-10
    assert helpers.relerr(ja._a_partial, data["a_partial"]) < 1e-10
    assert helpers.relerr(ja._b_partial, data["b_partial"]) < 1e-10
    pg = ja.pgradient()
    assert helpers.relerr(pg["acoeff"], data["pgrad_wf2acoeff"]) < 1e-10
    assert helpers.relerr(pg["bcoeff"], data["pgrad_wf2bcoeff"]) < 1e-10
    pgw = wf.pgradient()
    keys = ["wf1det_coeff", "wf1mo_coeff_alpha", "wf1mo_coeff_beta", "wf2acoeff", "wf2bcoeff"]
    if len(wf.wf_factors) > 2:
        keys.append("wf3ccoeff")
    for k in keys:
        assert pgw[k].shape == data["pgrad_" + k].shape, k
        assert helpers.relerr(pgw[k], data["pgrad_" + k]) < 1e-9, k


@pytest.mark.parametrize("name", ["he", "h2o", "open", "c2", "h2o_md", "h2o_3b", "h2o_md_3b", "h2o_cas", "h2o_cas_3b", "high_l"])
def test_cuda_reproduces_reference_golden(lib, name):
    import pyqmc_b200 as pq

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    assert np.array_equal(wf.parameters["wf2acoeff"], data["acoeff"])
    configs = pq.OpenConfigs(data["configs0"].copy())
    golden_replay.replay(data, wf, configs, lambda: pq.EnergyAccumulator(mol), device_vmc, check_internal)


@pytest.mark.parametrize("name", ["he", "h2o", "c2", "h2o_md"])
def test_tmoves_match_reference_golden(lib, name):
    """EnergyAccumulator.nonlocal_tmoves vs compute_tmoves of the reference (eval_ecp.py:43-80)."""
    import pyqmc_b200 as pq

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = pq.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    before = wf.value()[1].copy()
    acc = pq.EnergyAccumulator(mol)
    assert acc.has_nonlocal_moves()
    np.random.seed(22)
    tm = acc.nonlocal_tmoves(configs, wf, int(data["elist"][-1]), 0.02)
    assert tm["ratio"].shape == data["tmove_ratio"].shape
    assert helpers.relerr(tm["ratio"], data["tmove_ratio"]) < 1e-9
    assert helpers.relerr(tm["weight"], data["tmove_weight"]) < 1e-10
    assert np.abs(tm["configs"].configs - data["tmove_configs"]).max() < 1e-12
    assert np.array_equal(wf.value()[1], before), "T-move evaluation must not change the wave function state"


@pytest.mark.parametrize("name", ["h2o", "c2", "open"])
def test_dmc_propagate_on_device_objects_matches_reference_golden(lib, name):
    """The DMC loop (dmc.py:123-221: T-moves with masked updates without saved values, fixed-node
    drift-diffusion, weights) over the pyqmc_b200 protocol objects vs the reference's own run."""
    import pyqmc_b200 as pq
    from oracle import dmc_driver

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = pq.OpenConfigs(data["dmc_configs0"].copy())
    weights = np.ones(len(configs.configs))
    np.random.seed(41)
    out, configs, weights = dmc_driver.dmc_propagate(wf, configs, weights, 0.02, 10.0, 1.5, 1.7, nsteps=3,
                                                     accumulators={"energy": pq.EnergyAccumulator(mol)})
    golden_replay.check_dmc(data, out, configs, weights)
    # stochastic-comb branching keeps working on the host container
    np.random.seed(5)
    configs, weights, inds = dmc_driver.branch(configs, weights)
    assert configs.configs.shape == data["dmc_configs"].shape and np.allclose(weights, weights[0])
    wf.recompute(configs)


@pytest.mark.parametrize("name", ["h2o", "open", "h2o_md", "h2o_3b"])
def test_stochastic_reconfiguration_avg_on_device(lib, name):
    """StochasticReconfiguration.avg (stochastic_reconfiguration.py:85-118) on the device vs the reference
    formula evaluated with numpy on the per-walker arrays of the same objects (same ECP variates), for
    Jastrow, determinant and orbital parameters, with weights and a regularised walker."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import sr

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = pq.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    rng = np.random.RandomState(4)
    to_opt = {}
    for k in wf.parameters.keys():
        shape = np.shape(wf.parameters[k])
        m = rng.rand(*shape) > 0.35
        if k.endswith("bcoeff"):
            m[0, :] = False
        if np.any(m):
            to_opt[k] = m
    acc = sr.StochasticReconfiguration(pq.EnergyAccumulator(mol), sr.LinearTransform(wf.parameters, to_opt))
    weights = 0.5 + rng.rand(len(configs.configs))
    np.random.seed(77)
    dev = acc.avg(configs, wf, weights=weights)
    # numpy evaluation of the reference formula on the same per-walker quantities
    np.random.seed(77)
    den = acc.enacc(configs, wf)
    dp = acc.transform.serialize_gradients(wf.pgradient())
    assert dp.shape == (len(weights), acc.transform.nparams) and acc.transform.nparams > 20
    w = weights / weights.sum()
    _, f = sr.nodal_regularization(den["grad2"])
    dpr = dp * f[:, None]
    assert helpers.relerr(dev["dppsi"], np.average(dpr, weights=w, axis=0)) < 1e-10
    assert helpers.relerr(dev["dpH"], np.einsum("i,ij->j", den["total"], w[:, None] * dpr)) < 1e-10
    assert helpers.relerr(dev["dpidpj"], np.einsum("ij,ik->jk", dp, w[:, None] * dpr)) < 1e-10
    for k in ("ke", "ee", "ei", "ecp", "grad2", "total"):
        assert abs(dev[k] - np.average(den[k], weights=w)) <= 1e-10 * max(1.0, abs(np.average(den[k], weights=w))), k
    steps, report = acc.delta_p([0.1, 0.2], dev)
    assert len(steps) == 2 and np.all(np.isfinite(steps[0])) and np.isfinite(report["SRdot"])


@pytest.mark.parametrize("name", ["h2o", "h2o_md"])
def test_stochastic_reconfiguration_matches_reference_golden(lib, name):
    """qmcb_sr_avg vs the reference's own StochasticReconfiguration.avg (tests/golden/sr_<name>.npz)."""
    import make_golden
    import pyqmc_b200 as pq
    from pyqmc_b200 import sr

    data = golden_replay.load(name)
    gold = golden_replay.load("sr_" + name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = pq.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    to_opt, weights = make_golden.sr_inputs(wf.parameters, len(configs.configs))
    acc = sr.StochasticReconfiguration(pq.EnergyAccumulator(mol), sr.LinearTransform(wf.parameters, to_opt))
    np.random.seed(77)
    d = acc.avg(configs, wf, weights=weights)
    for k in ("dpH", "dppsi", "dpidpj", "total", "ke", "ecp", "grad2"):
        assert helpers.relerr(np.asarray(d[k]), gold["sr_" + k]) < 1e-9, k
