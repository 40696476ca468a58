"""CPU tests: the numpy oracle against the committed golden vectors produced by the reference
(tests/golden/*.npz), and -- when /root/reference is present (this container only) -- directly
against the live reference."""
import numpy as np
import pytest

import golden_replay
import helpers
import refload

SYSTEMS = ["he", "h2o", "open", "c2", "h2o_md", "h2o_3b", "h2o_md_3b", "h2o_cas", "h2o_cas_3b", "high_l",
           "h2o_cx", "h2o_md_cx"]  # *_cx: complex orbital / determinant coefficients (wf.dtype == complex)


def oracle_vmc(wf, configs, accumulators):
    from oracle import vmc_driver

    record = []
    df, configs = vmc_driver.vmc(wf, configs, tstep=0.5, nblocks=2, nsteps_per_block=3, accumulators=accumulators,
                                 record=record)
    N, ne = configs.configs.shape[:2]
    accepts = np.array([r["accept"] for r in record]).reshape(2, 3, ne, N)
    return df, configs, accepts


def check_internal(wf, data):
    sl, ja = wf.wf_factors[:2]
    if len(wf.wf_factors) > 2:
        assert helpers.relerr(wf.wf_factors[2].P_i, data["P_i"]) < 1e-10
        assert helpers.relerr(wf.wf_factors[2].a_values, data["a3_values"]) < 1e-10
    for s in (0, 1):
        assert helpers.relerr(sl._inverse[s], data[f"inverse{s}"]) < 1e-9
        assert golden_replay.same_sign(sl._dets[s][0], data[f"dets{s}"][0])
        assert np.abs(sl._dets[s][1] - data[f"dets{s}"][1]).max() < 1e-10
    assert helpers.relerr(ja._a_partial, data["a_partial"]) < 1e-10
    assert helpers.relerr(ja._b_partial, data["b_partial"]) < 1e-10
    if not any(k.startswith("pgrad_") for k in data):
        return
    pg = wf.pgradient()
    for k in ("wf1det_coeff", "wf1mo_coeff_alpha", "wf1mo_coeff_beta", "wf2acoeff", "wf2bcoeff", "wf3ccoeff"):
        if "pgrad_" + k in data:
            assert helpers.relerr(pg[k], data["pgrad_" + k]) < 1e-9, k


@pytest.mark.parametrize("name", SYSTEMS)
def test_oracle_reproduces_reference_golden(name):
    from oracle.local_energy import EnergyOracle
    from oracle.walkers import Walkers

    data = golden_replay.load(name)
    mol, mf, _, orc = helpers.make_pair(name, seed=1) if False else _oracle_only(name)
    assert np.array_equal(orc.wf_factors[1].parameters["acoeff"], data["acoeff"])
    assert np.array_equal(orc.wf_factors[1].parameters["bcoeff"], data["bcoeff"])
    configs = Walkers(data["configs0"].copy())
    golden_replay.replay(data, orc, configs, lambda: EnergyOracle(mol), oracle_vmc, check_internal)


PBC_SYSTEMS = ["ortho", "rotcubic", "diamond211"]
EWALD_GMAX = 10  # as in tests/golden/make_golden.py


def periodic_walkers(cls, data, mol, key="configs0", wkey="wrap0"):
    """Walker container holding exactly the recorded (already wrapped) positions and wrap vectors."""
    w = cls(data[key].copy(), mol.lattice_vectors())
    w.configs = data[key].copy()
    w.wrap = data[wkey].copy()
    return w


@pytest.mark.parametrize("name", PBC_SYSTEMS + ["ortho_3b", "diamond211_3b", "ortho_twist", "diamond211_twist", "ortho_md",
                                  "diamond211_md"])
def test_oracle_reproduces_reference_golden_periodic(name):
    """Periodic systems (minimal-image modes diagonal / orthogonal / general, two k-points with the
    wrap phase, Ewald): the oracle replays the reference's recorded calls."""
    from oracle.local_energy import EnergyOracle
    from oracle.pbc import PeriodicWalkers

    data = golden_replay.load(name)
    mol, mf, _, orc = _oracle_only(name)
    assert np.array_equal(orc.wf_factors[1].parameters["acoeff"], data["acoeff"])
    configs = periodic_walkers(PeriodicWalkers, data, mol)
    golden_replay.replay(data, orc, configs, lambda: EnergyOracle(mol, ewald_gmax=EWALD_GMAX), oracle_vmc,
                         check_internal)


def _oracle_only(name):
    """helpers.make_pair builds the device objects too; here only the oracle side is needed."""
    from oracle.jastrow2 import JastrowOracle
    from oracle.product import ProductOracle
    from oracle.slater_det import SlaterOracle

    mol, mf, dets = helpers.make_system(name)
    oj = JastrowOracle.default(mol)
    has_cusp = len(oj.a_basis) > 4
    a0, ac, bc = helpers.jastrow_coefficients(oj.parameters["acoeff"].shape, oj.parameters["bcoeff"].shape, has_cusp, 1)
    oj.parameters["acoeff"][:, a0:, :] = ac[:, a0:, :]
    oj.parameters["bcoeff"][1:, :] = bc[1:, :]
    if hasattr(mol, "a"):
        from oracle.pbc import SlaterPbcOracle

        factors = [SlaterPbcOracle(mol, mf, determinants=dets), oj]
    else:
        factors = [SlaterOracle(mol, mf, determinants=dets), oj]
    if name.endswith("_3b"):
        from oracle.jastrow3 import Jastrow3Oracle

        oj3 = Jastrow3Oracle.default(mol)
        oj3.parameters["ccoeff"][...] = helpers.three_body_coefficients(oj3.parameters["ccoeff"].shape)
        factors.append(oj3)
    return mol, mf, None, ProductOracle(*factors)


@pytest.mark.parametrize("name", ["h2o", "c2"])
def test_oracle_tmoves_match_golden(name):
    from oracle.local_energy import EnergyOracle
    from oracle.walkers import Walkers

    data = golden_replay.load(name)
    mol, mf, _, orc = _oracle_only(name)
    configs = Walkers(data["configs1"].copy())
    orc.recompute(configs)
    np.random.seed(22)
    tm = EnergyOracle(mol).nonlocal_tmoves(configs, orc, int(data["elist"][-1]), 0.02)
    assert helpers.relerr(tm["ratio"], data["tmove_ratio"]) < 1e-9
    assert helpers.relerr(tm["weight"], data["tmove_weight"]) < 1e-10
    assert np.abs(tm["configs"] - data["tmove_configs"]).max() < 1e-12


def test_sherman_morrison_oracle_vs_numpy():
    """tests/unit/test_sherman_morrison.py:20-82 restated for the oracle's rank-1 update."""
    from oracle.slater_det import rank1_row_update

    rng = np.random.RandomState(0)
    n, nconf, ndet, e = 10, 4, 8, 2
    u, _, v = np.linalg.svd(rng.randn(n, n))
    sv = (rng.rand(nconf, ndet, n) + 1) * rng.choice([-1, 1], (nconf, ndet, n))
    mat = np.einsum("ij,...hj,jk->...hik", u, sv, v)
    inv = np.linalg.inv(mat)
    vec = rng.randn(nconf, ndet, n) + 2 * mat[..., e, :]
    new = mat.copy()
    new[..., e, :] = vec
    ratio, upd = rank1_row_update(e, inv, vec)
    assert np.abs(ratio - np.linalg.det(new) / np.linalg.det(mat)).max() < 1e-12
    assert np.abs(upd - np.linalg.inv(new)).max() < 1e-11


@pytest.mark.skipif(not refload.available(), reason="/root/reference only exists in the build container")
def test_solid_harmonics_match_reference_tables():
    refload.load()
    import pyqmc.wf.numba.spherical_harmonics as hsh
    from oracle import solid_harmonics as sh

    rng = np.random.RandomState(0)
    for l in (2, 3, 4):  # the closed form used for l = 5 reproduces the hand-tabulated polynomials
        for i, m in enumerate(range(-l, l + 1)):
            a, b = sh.TABLES[l][i], sh.closed_form(l, m)
            assert all(abs(a.get(k, 0.0) - b.get(k, 0.0)) < 1e-13 for k in set(a) | set(b)), (l, m)
    for _ in range(20):
        x, y, z = rng.randn(3)
        s, dx, dy, dz = np.zeros(36), np.zeros(36), np.zeros(36), np.zeros(36)
        hsh.SPH5_GRAD(x, y, z, x * x, y * y, z * z, s, dx, dy, dz)
        S, dS = sh.evaluate(5, np.array(x), np.array(y), np.array(z), deriv=True)
        scale = max(1.0, np.abs(s).max())
        assert np.abs(S - s).max() < 1e-12 * scale
        for a, d in enumerate((dx, dy, dz)):
            assert np.abs(dS[a] - d).max() < 1e-12 * max(1.0, np.abs(d).max())


def test_device_sph_tables_equal_oracle_tables():
    """The CUDA code generator and the oracle must tabulate the same polynomials."""
    import importlib.util
    import os

    from oracle import solid_harmonics as sh

    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyqmc_b200", "csrc", "gen_sph.py")
    spec = importlib.util.spec_from_file_location("gen_sph", p)
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    t = gen.tables()
    for l in range(6):
        assert len(t[l]) == len(sh.TABLES[l]) == 2 * l + 1
        for a, b in zip(t[l], sh.TABLES[l]):
            assert a == b


@pytest.mark.parametrize("name", ["h2o", "c2", "open", "h2o_md"])
def test_oracle_dmc_propagate_matches_reference_golden(name):
    """oracle/dmc_driver.py (restating dmc.py:22-235) driving the oracle wave function vs the
    reference's own dmc_propagate with T-moves: weights, walkers and weighted averages."""
    from oracle import dmc_driver
    from oracle.local_energy import EnergyOracle
    from oracle.walkers import Walkers

    data = golden_replay.load(name)
    mol, mf, _, orc = _oracle_only(name)
    configs = Walkers(data["dmc_configs0"].copy())
    weights = np.ones(len(configs.configs))
    np.random.seed(41)
    out, configs, weights = dmc_driver.dmc_propagate(orc, configs, weights, 0.02, 10.0, 1.5, 1.7, nsteps=3,
                                                     accumulators={"energy": EnergyOracle(mol)})
    golden_replay.check_dmc(data, out, configs, weights)


@pytest.mark.parametrize("name", ["ortho", "diamond211", "ortho_md", "rotcubic", "diamond211_md"])
def test_oracle_dmc_propagate_matches_reference_golden_periodic(name):
    """The same for periodic systems (Ewald energy, T-moves wrapped twice as propose_tmoves does, dmc.py:110):
    walkers, wrap vectors, weights and weighted averages of the reference's dmc_propagate."""
    from oracle import dmc_driver
    from oracle.local_energy import EnergyOracle
    from oracle.pbc import PeriodicWalkers

    data = golden_replay.load(name)
    mol, mf, _, orc = _oracle_only(name)
    configs = periodic_walkers(PeriodicWalkers, data, mol, "dmc_configs0", "dmc_wrap0")
    weights = np.ones(len(configs.configs))
    np.random.seed(41)
    out, configs, weights = dmc_driver.dmc_propagate(orc, configs, weights, 0.02, 10.0, 1.5, 1.7, nsteps=3,
                                                     accumulators={"energy": EnergyOracle(mol, ewald_gmax=EWALD_GMAX)})
    golden_replay.check_dmc(data, out, configs, weights)


@pytest.mark.parametrize("name", ["h2o", "h2o_md"])
def test_oracle_stochastic_reconfiguration_matches_reference_golden(name):
    """The SR averages (stochastic_reconfiguration.py:85-118) formed with numpy from the oracle's
    pgradient and energy vs the reference's own accumulator output."""
    import make_golden
    from oracle.local_energy import EnergyOracle
    from oracle.walkers import Walkers

    data = golden_replay.load(name)
    gold = golden_replay.load("sr_" + name)
    mol, mf, _, orc = _oracle_only(name)
    configs = Walkers(data["configs1"].copy())
    orc.recompute(configs)
    to_opt, weights = make_golden.sr_inputs(orc.parameters, len(configs.configs))
    pg = orc.pgradient()
    dp = np.concatenate([pg[k].reshape(len(weights), -1)[:, m.ravel()] for k, m in to_opt.items()], axis=1)
    np.random.seed(77)
    en = EnergyOracle(mol)(configs, orc)
    w = weights / weights.sum()
    r = 1.0 / en["grad2"]
    cut = 1e-3
    f = np.where(r < cut**2, 9 / cut**2 * r - 15 / cut**4 * r**2 + 7 / cut**6 * r**3, 1.0)
    dpr = dp * f[:, None]
    assert helpers.relerr(np.average(dpr, weights=w, axis=0), gold["sr_dppsi"]) < 1e-9
    assert helpers.relerr(np.einsum("i,ij->j", en["total"], w[:, None] * dpr), gold["sr_dpH"]) < 1e-9
    assert helpers.relerr(np.einsum("ij,ik->jk", dp, w[:, None] * dpr), gold["sr_dpidpj"]) < 1e-9


@pytest.mark.parametrize("name", PBC_SYSTEMS)
def test_oracle_periodic_tmoves_match_golden(name):
    from oracle.local_energy import EnergyOracle
    from oracle.pbc import PeriodicWalkers

    data = golden_replay.load(name)
    mol, mf, _, orc = _oracle_only(name)
    configs = periodic_walkers(PeriodicWalkers, data, mol, "configs1", "wrap1")
    orc.recompute(configs)
    np.random.seed(22)
    tm = EnergyOracle(mol, ewald_gmax=EWALD_GMAX).nonlocal_tmoves(configs, orc, int(data["elist"][-1]), 0.02)
    assert helpers.relerr(tm["ratio"], data["tmove_ratio"]) < 1e-9
    assert helpers.relerr(tm["weight"], data["tmove_weight"]) < 1e-10
    assert np.abs(tm["configs"] - data["tmove_configs"]).max() < 1e-12
