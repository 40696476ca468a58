"""GPU parity for periodic systems: the CUDA path against the reference's golden vectors
(tests/golden/{ortho,rotcubic,diamond211}.npz: the three minimal-image modes, two k-points with
the wrap phase, Ewald energy, stochastic ECP, VMC accept masks) and against the numpy oracle."""
import numpy as np
import pytest

import golden_replay
import helpers

pytestmark = pytest.mark.gpu

PBC_SYSTEMS = ["ortho", "rotcubic", "diamond211"]
EWALD_GMAX = 10  # as in tests/golden/make_golden.py


def periodic_configs(data, mol, key="configs0", wkey="wrap0"):
    import pyqmc_b200 as pq

    c = pq.PeriodicConfigs(data[key].copy(), mol.lattice_vectors())
    c.configs = data[key].copy()
    c.wrap = data[wkey].copy()
    return c


def device_vmc(wf, configs, accumulators):
    from pyqmc_b200 import mc

    rows, accepts = [], []
    for block in range(2):
        avg, configs, data = mc.vmc_block_device(wf, configs, 0.5, 3, accumulators, return_walker_data=True)
        rows.append(avg)
        accepts.append(data["accept"])
    df = {k: np.asarray([r[k] for r in rows]) for k in rows[0]}
    return df, configs, np.array(accepts)


def check_internal(wf, data):
    sl, ja = wf.wf_factors[:2]
    for s in (0, 1):
        assert helpers.relerr(sl._inverse[s], data[f"inverse{s}"]) < 1e-9
        assert np.array_equal(sl._dets[s][0], data[f"dets{s}"][0])
        assert np.abs(sl._dets[s][1] - data[f"dets{s}"][1]).max() < 1e-10
    assert helpers.relerr(ja._a_partial, data["a_partial"]) < 1e-10
    assert helpers.relerr(ja._b_partial, data["b_partial"]) < 1e-10
    pg = wf.pgradient()  # periodic orbitals: every MO column uses the AO set of its own k-point
    keys = ["wf1det_coeff", "wf1mo_coeff_alpha", "wf1mo_coeff_beta", "wf2acoeff", "wf2bcoeff"]
    if len(wf.wf_factors) > 2:  # periodic three-body factor (minimal-image displacements)
        keys.append("wf3ccoeff")
        assert helpers.relerr(wf.wf_factors[2].P_i, data["P_i"]) < 1e-10
    for k in keys:
        assert pg[k].shape == data["pgrad_" + k].shape, k
        assert helpers.relerr(pg[k], data["pgrad_" + k]) < 1e-9, k


@pytest.mark.parametrize("name", PBC_SYSTEMS + ["ortho_3b", "diamond211_3b", "ortho_md", "diamond211_md"])
def test_cuda_reproduces_reference_golden_periodic(lib, name):
    import pyqmc_b200 as pq

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    assert np.array_equal(wf.parameters["wf2acoeff"], data["acoeff"])
    configs = periodic_configs(data, mol)
    golden_replay.replay(data, wf, configs, lambda: pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX), device_vmc,
                         check_internal)


@pytest.mark.parametrize("name", PBC_SYSTEMS + ["ortho_3b", "diamond211_3b", "ortho_md", "diamond211_md"])
def test_periodic_per_call_vmc_equals_device_resident_block(lib, name):
    """The reference driver loop over the protocol calls and the device-resident periodic block
    consume the same variates and must accept the same moves (single-determinant Slater-Jastrow: the fused
    two-launch chain; multi-determinant / three-body: k_pbc_move_general + the update kernels)."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import mc

    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    np.random.seed(11)
    c1 = pq.initial_guess(mol, 40)
    c2 = c1.copy()
    acc = {"energy": pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)}
    np.random.seed(12)
    avg1, c1 = mc.vmc_block_device(wf, c1, 0.4, 2, acc)
    mol2, mf2, wf2, _ = helpers.make_pair(name, seed=1)

    np.random.seed(12)
    from oracle import vmc_driver  # the oracle's restatement of the reference loop (mc.py:102-153), over device objects

    avg2, c2 = vmc_driver.vmc_worker(wf2, c2, 0.4, 2, {"energy": pq.EnergyAccumulator(mol2, ewald_gmax=EWALD_GMAX)})
    assert avg1["acceptance"] == avg2["acceptance"]
    assert np.abs(c1.configs - c2.configs).max() < 1e-9
    assert np.array_equal(c1.wrap, c2.wrap)
    for k in ("energytotal", "energyke", "energyee", "energyei", "energyecp"):
        assert abs(avg1[k] - avg2[k]) <= 1e-9 * max(1.0, abs(avg2[k])), k


def test_diamond_supercell_against_oracle(lib):
    """Config C4 shape (2x2x2 diamond, 64 electrons, 8 k-points, n = 32 Sherman-Morrison) at a few
    walkers: protocol calls and the energy against the numpy oracle."""
    import pyqmc_b200 as pq
    from oracle.local_energy import EnergyOracle

    mol, mf, wf, orc = helpers.make_pair("diamond222", seed=1)
    np.random.seed(3)
    configs = pq.initial_guess(mol, 6)
    oc = helpers.to_oracle_walkers(configs)
    s, l = wf.recompute(configs)
    so, lo = orc.recompute(oc)
    assert np.array_equal(s, so)
    assert np.abs(l - lo).max() < 1e-10 * max(1.0, np.abs(lo).max())
    rng = np.random.RandomState(4)
    for e in (0, 31, 40, 63):
        new = configs.configs[:, e] + 0.3 * rng.randn(6, 3)
        ep, eo = configs.make_irreducible(e, new.copy()), oc.make_irreducible(e, new.copy())
        g, v, saved = wf.gradient_value(e, ep)
        go, vo, savedo = orc.gradient_value(e, eo)
        assert helpers.relerr(g, go) < 1e-10 and helpers.relerr(v, vo) < 1e-10
        gl, lap = wf.gradient_laplacian(e, ep)
        glo, lapo = orc.gradient_laplacian(e, eo)
        assert helpers.relerr(lap, lapo) < 1e-10
        mask = rng.rand(6) > 0.4
        wf.updateinternals(e, ep, configs, mask=mask, saved_values=saved)
        orc.updateinternals(e, eo, oc, mask=mask, saved_values=savedo)
        configs.move(e, ep, mask)
        oc.move(e, eo, mask)
        assert np.abs(wf.value()[1] - orc.value()[1]).max() < 1e-10 * max(1.0, np.abs(lo).max())
    np.random.seed(9)
    en = pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)(configs, wf)
    np.random.seed(9)
    eno = EnergyOracle(mol, ewald_gmax=EWALD_GMAX)(oc, orc)
    for k in ("ke", "ee", "ei", "ecp", "total"):
        assert helpers.relerr(en[k], eno[k]) < 1e-10, k


def test_periodic_stochastic_reconfiguration_avg(lib):
    """qmcb_sr_avg on a periodic wave function (Jastrow and orbital parameters) vs numpy on the same
    per-walker arrays."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import sr

    data = golden_replay.load("diamond211")
    mol, mf, wf, _ = helpers.make_pair("diamond211", seed=1)
    configs = periodic_configs(data, mol)
    wf.recompute(configs)
    to_opt = {k: np.ones(np.shape(wf.parameters[k]), dtype=bool) for k in wf.parameters.keys()}
    to_opt["wf2bcoeff"][0, :] = False
    acc = sr.StochasticReconfiguration(pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX), sr.LinearTransform(wf.parameters, to_opt))
    np.random.seed(77)
    dev = acc.avg(configs, wf)
    np.random.seed(77)
    den = acc.enacc(configs, wf)
    dp = acc.transform.serialize_gradients(wf.pgradient())
    w = np.full(len(dp), 1.0 / len(dp))
    _, f = sr.nodal_regularization(den["grad2"])
    dpr = dp * f[:, None]
    assert helpers.relerr(dev["dppsi"], np.average(dpr, weights=w, axis=0)) < 1e-10
    assert helpers.relerr(dev["dpH"], np.einsum("i,ij->j", den["total"], w[:, None] * dpr)) < 1e-10
    assert helpers.relerr(dev["dpidpj"], np.einsum("ij,ik->jk", dp, w[:, None] * dpr)) < 1e-10


@pytest.mark.parametrize("name", PBC_SYSTEMS)
def test_periodic_tmoves_match_reference_golden(lib, name):
    """EnergyAccumulator.nonlocal_tmoves on a periodic system vs compute_tmoves of the reference
    (eval_ecp.py:43-80 with make_irreducible, coord.py:168-184): ratios, weights, wrapped positions."""
    import pyqmc_b200 as pq

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = periodic_configs(data, mol, "configs1", "wrap1")
    wf.recompute(configs)
    acc = pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)
    np.random.seed(22)
    tm = acc.nonlocal_tmoves(configs, wf, int(data["elist"][-1]), 0.02)
    assert helpers.relerr(tm["ratio"], data["tmove_ratio"]) < 1e-9
    assert helpers.relerr(tm["weight"], data["tmove_weight"]) < 1e-10
    assert np.abs(tm["configs"].configs - data["tmove_configs"]).max() < 1e-9


@pytest.mark.parametrize("name", ["ortho", "diamond211"])
def test_periodic_dmc_through_the_protocol(lib, name):
    """dmc_propagate (dmc.py:123-221: T-moves, fixed-node drift-diffusion, weights) driving the periodic
    device objects through the protocol calls vs the same loop over the oracle objects."""
    import pyqmc_b200 as pq
    from oracle import dmc_driver
    from oracle.local_energy import EnergyOracle

    mol, mf, wf, orc = helpers.make_pair(name, seed=1)
    np.random.seed(3)
    configs = pq.initial_guess(mol, 10)
    oc = helpers.to_oracle_walkers(configs)
    w1, w2 = np.ones(10), np.ones(10)
    np.random.seed(4)
    out1, configs, w1 = dmc_driver.dmc_propagate(wf, configs, w1, 0.02, 10.0, -5.0, -5.1, nsteps=2,
                                                 accumulators={"energy": pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)})
    np.random.seed(4)
    out2, oc, w2 = dmc_driver.dmc_propagate(orc, oc, w2, 0.02, 10.0, -5.0, -5.1, nsteps=2,
                                            accumulators={"energy": EnergyOracle(mol, ewald_gmax=EWALD_GMAX)})
    assert np.abs(configs.configs - oc.configs).max() < 1e-9
    assert np.array_equal(configs.wrap, oc.wrap)
    assert helpers.relerr(w1, w2) < 1e-8
    for k in ("energytotal", "acceptance", "tmove_acceptance", "weight"):
        assert abs(out1[k] - out2[k]) <= 1e-8 * max(1.0, abs(out2[k])), k


@pytest.mark.parametrize("name", ["ortho", "diamond211", "ortho_md", "rotcubic", "diamond211_md"])
def test_device_resident_periodic_dmc_matches_reference_golden(lib, name):
    """qmcb_dmc_block on periodic wave functions (k_pbc_move_general<16, true>: fixed-node drift-diffusion with the
    wrapped proposal, T-moves wrapped as propose_tmoves wraps them, Ewald energy) against the reference's own
    dmc_propagate: walkers, wrap vectors, weights and weighted block averages."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = periodic_configs(data, mol, "dmc_configs0", "dmc_wrap0")
    weights = np.ones(len(configs.configs))
    acc = {"energy": pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)}
    assert dmc._device_dmc_path(wf, acc, ("energy", "total"))
    np.random.seed(41)
    out, configs, weights = dmc.dmc_propagate(wf, configs, weights, 0.02, 10.0, 1.5, 1.7, nsteps=3, accumulators=acc)
    golden_replay.check_dmc(data, out, configs, weights)


@pytest.mark.parametrize("name", PBC_SYSTEMS + ["ortho_3b", "diamond211_3b", "ortho_md", "diamond211_md"])
def test_device_resident_periodic_dmc_matches_oracle_loop(lib, name):
    """Device block vs the oracle loop over the oracle objects, more walkers: same walkers, wrap vectors, weights,
    averages, and the same position of the random stream afterwards.  tstep = 0.1 so that T-moves are accepted too
    (their weight grows with the time step)."""
    import pyqmc_b200 as pq
    from oracle import dmc_driver
    from oracle.local_energy import EnergyOracle
    from pyqmc_b200 import dmc

    mol, mf, wf, orc = helpers.make_pair(name, seed=1)
    np.random.seed(3)
    configs = pq.initial_guess(mol, 24)
    oc = helpers.to_oracle_walkers(configs)
    w1, w2 = np.ones(24), np.ones(24)
    np.random.seed(4)
    out1, configs, w1 = dmc.dmc_propagate(wf, configs, w1, 0.1, 10.0, -5.0, -5.1, nsteps=2,
                                          accumulators={"energy": pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)})
    tail1 = np.random.rand()
    np.random.seed(4)
    out2, oc, w2 = dmc_driver.dmc_propagate(orc, oc, w2, 0.1, 10.0, -5.0, -5.1, nsteps=2,
                                            accumulators={"energy": EnergyOracle(mol, ewald_gmax=EWALD_GMAX)})
    tail2 = np.random.rand()
    assert tail1 == tail2
    assert np.abs(configs.configs - oc.configs).max() < 1e-9
    assert np.array_equal(configs.wrap, oc.wrap)
    assert helpers.relerr(w1, w2) < 1e-9
    for k in out2:
        assert abs(out1[k] - out2[k]) <= 1e-9 * max(1.0, abs(out2[k])), k


@pytest.mark.parametrize("name,nconf", [("diamond211", 37), ("ortho", 10)])
def test_periodic_block_walker_ranges_on_separate_streams(lib, monkeypatch, name, nconf):
    """The fused periodic chain cuts the ensemble into walker ranges whose move chains run concurrently on separate
    streams (QMCB_PBC_SPLIT, default 2 from 512 walkers): 1, 2, 3 and 4 ranges (ragged sizes) give bit-identical
    accept masks, walkers, wrap vectors and per-walker energies."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import mc

    results = []
    for nsplit in ("1", "2", "3", "4"):
        monkeypatch.setenv("QMCB_PBC_SPLIT", nsplit)
        mol, mf, wf, _ = helpers.make_pair(name, seed=1)
        np.random.seed(11)
        configs = pq.initial_guess(mol, nconf)
        np.random.seed(12)
        avg, configs, data = mc.vmc_block_device(wf, configs, 0.4, 3, {"energy": pq.EnergyAccumulator(mol, ewald_gmax=EWALD_GMAX)},
                                                 return_walker_data=True)
        results.append((avg, configs.configs.copy(), configs.wrap.copy(), data["accept"].copy(), data["energy"].copy()))
    ref = results[0]
    assert ref[3].any() and not ref[3].all()
    for r in results[1:]:
        assert np.array_equal(r[3], ref[3])
        assert np.array_equal(r[1], ref[1]) and np.array_equal(r[2], ref[2])
        assert np.array_equal(r[4], ref[4])
        assert r[0]["energytotal"] == ref[0]["energytotal"] and r[0]["acceptance"] == ref[0]["acceptance"]


def test_periodic_pair_caches_equal_recomputation_at_full_walker_count(lib, monkeypatch):
    """C4 shape, 1024 walkers: the fused chain with pair caches (drift at the current position from cached pair
    gradients) against the same chain recomputing the minimal-image Jastrow (QMCB_PBC_NO_PAIRCACHE=1).  The two
    differ by rounding in the drift only, so the accept masks of the first step are identical and the walkers agree
    to rounding after it (over long runs rounding differences grow along each walker's trajectory, as they do
    between any two summation orders)."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import mc

    out = []
    for env in (None, "1"):
        if env is None:
            monkeypatch.delenv("QMCB_PBC_NO_PAIRCACHE", raising=False)
        else:
            monkeypatch.setenv("QMCB_PBC_NO_PAIRCACHE", env)
        mol, mf, wf, _ = helpers.make_pair("diamond222", seed=1)
        np.random.seed(3)
        configs = pq.initial_guess(mol, 1024)
        np.random.seed(4)
        avg, configs, data = mc.vmc_block_device(wf, configs, 0.5, 1, {}, return_walker_data=True)
        out.append((data["accept"].copy(), configs.configs.copy(), configs.wrap.copy()))
    (a1, c1, w1), (a2, c2, w2) = out
    assert 0.2 < a1.mean() < 0.8
    assert np.array_equal(a1, a2)
    assert np.array_equal(w1, w2) and np.abs(c1 - c2).max() < 1e-11
