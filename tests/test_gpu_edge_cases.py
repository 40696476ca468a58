"""Edge cases of the device path the reference tests or tolerates: single walker, empty and
all-False masks, changing walker counts, in-place parameter changes, an empty spin channel, no
ECP, Slater-only energies; plus finite-difference checks in the style of pyqmc/wf/testwf.py:149-289
that do not involve the oracle at all."""
import numpy as np
import pytest

import helpers
from helpers import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nconf", [1, 2, 33])
def test_small_and_ragged_walker_counts(lib, nconf):
    import pyqmc_b200 as pq
    from oracle.local_energy import EnergyOracle

    mol, mf, wf, orc = helpers.make_pair("h2o")
    np.random.seed(1)
    configs = pq.initial_guess(mol, nconf)
    oc = helpers.to_oracle_walkers(configs)
    assert relerr(wf.recompute(configs)[1], orc.recompute(oc)[1]) < 1e-10
    np.random.seed(2)
    en = pq.EnergyAccumulator(mol)(configs, wf)
    np.random.seed(2)
    eo = EnergyOracle(mol)(oc, orc)
    assert relerr(en["total"], eo["total"]) < 1e-10
    np.random.seed(3)
    df, configs = pq.vmc(wf, configs, nblocks=1, nsteps_per_block=2, accumulators={"energy": pq.EnergyAccumulator(mol)})
    assert configs.configs.shape == (nconf, 8, 3) and np.isfinite(df["energytotal"]).all()


def test_all_false_and_all_true_masks(lib):
    import pyqmc_b200 as pq

    mol, mf, wf, orc = helpers.make_pair("h2o")
    np.random.seed(1)
    N = 9
    configs = pq.initial_guess(mol, N)
    oc = helpers.to_oracle_walkers(configs)
    wf.recompute(configs)
    orc.recompute(oc)
    e = 5
    newpos = configs.configs[:, e] + 0.2
    ep = configs.make_irreducible(e, newpos)
    none = np.zeros(N, dtype=bool)
    r, _ = wf.testvalue(e, ep, none)
    assert r.shape == (0,)
    aux = configs.make_irreducible(e, np.repeat(newpos[:, None, :], 6, axis=1))
    r, _ = wf.testvalue(e, aux, none)
    assert r.shape == (0, 6)
    before = wf.value()[1].copy()
    wf.updateinternals(e, ep, configs, mask=none)  # dmc.py:175 with no accepted T-move
    assert np.array_equal(wf.value()[1], before)
    wf.updateinternals(e, ep, configs, mask=[True] * N)
    orc.updateinternals(e, oc.make_irreducible(e, newpos.copy()), oc, mask=np.ones(N, dtype=bool))
    assert relerr(wf.value()[1], orc.value()[1]) < 1e-10


def test_walker_count_change_and_parameter_update(lib):
    import pyqmc_b200 as pq

    mol, mf, wf, orc = helpers.make_pair("h2o")
    for n in (7, 40, 3):
        np.random.seed(n)
        configs = pq.initial_guess(mol, n)
        oc = helpers.to_oracle_walkers(configs)
        assert relerr(wf.recompute(configs)[1], orc.recompute(oc)[1]) < 1e-10
    # parameters are read at recompute (wftools.read_wf / linemin assign into wf.parameters)
    wf.parameters["wf2bcoeff"][1:, :] *= 1.7
    orc.wf_factors[1].parameters["bcoeff"][1:, :] *= 1.7
    wf.parameters["wf1mo_coeff_alpha"][:, 0] += 0.05
    orc.wf_factors[0].parameters["mo_coeff_alpha"][:, 0] += 0.05
    wf.parameters["wf2acoeff"] = wf.parameters["wf2acoeff"] * 0.5
    orc.wf_factors[1].parameters["acoeff"] *= 0.5
    assert relerr(wf.recompute(configs)[1], orc.recompute(oc)[1]) < 1e-10
    g1, v1, _ = wf.gradient_value(2, configs.electron(2))
    g2, v2, _ = orc.gradient_value(2, oc.electron(2))
    assert relerr(g1, g2) < 1e-10


def test_empty_spin_channel_and_no_ecp(lib):
    import pyqmc_b200 as pq
    from oracle import vmc_driver
    from oracle.local_energy import EnergyOracle

    mol, mf, wf, orc = helpers.make_pair("hatom")
    N = 11
    np.random.seed(4)
    configs = pq.initial_guess(mol, N)
    oc = helpers.to_oracle_walkers(configs)
    s1, l1 = wf.recompute(configs)
    s2, l2 = orc.recompute(oc)
    assert np.array_equal(s1, s2) and relerr(l1, l2) < 1e-10
    acc = pq.EnergyAccumulator(mol)
    assert not acc.has_nonlocal_moves()
    tm = acc.nonlocal_tmoves(configs, wf, 0, 0.02)
    assert tm["ratio"].shape == (N, 0) and tm["weight"].shape == (N, 0)
    np.random.seed(5)
    en = acc(configs, wf)
    np.random.seed(5)
    eo = EnergyOracle(mol)(oc, orc)
    for k in eo:
        assert relerr(en[k], eo[k]) < 1e-10 or np.abs(en[k] - eo[k]).max() < 1e-12, k
    np.random.seed(6)
    df, configs = pq.vmc(wf, configs, nblocks=1, nsteps_per_block=3, accumulators={"energy": acc})
    np.random.seed(6)
    odf, oc = vmc_driver.vmc(orc, oc, nblocks=1, nsteps_per_block=3, accumulators={"energy": EnergyOracle(mol)})
    assert np.array_equal(df["acceptance"], odf["acceptance"])
    assert np.abs(configs.configs - oc.configs).max() < 1e-10


def test_slater_only_energy_and_vmc(lib):
    import pyqmc_b200 as pq
    from oracle import vmc_driver
    from oracle.local_energy import EnergyOracle

    mol, mf, wf, orc = helpers.make_pair("c2", jastrow=False)
    np.random.seed(7)
    configs = pq.initial_guess(mol, 13)
    oc = helpers.to_oracle_walkers(configs)
    np.random.seed(8)
    df, configs = pq.vmc(wf, configs, nblocks=1, nsteps_per_block=2, accumulators={"energy": pq.EnergyAccumulator(mol)})
    np.random.seed(8)
    odf, oc = vmc_driver.vmc(orc, oc, nblocks=1, nsteps_per_block=2, accumulators={"energy": EnergyOracle(mol)})
    assert np.array_equal(df["acceptance"], odf["acceptance"])
    assert abs(df["energytotal"][0] - odf["energytotal"][0]) < 1e-9 * abs(odf["energytotal"][0])


@pytest.mark.parametrize("name", ["h2o", "h2o_md", "h2o_3b", "open"])
def test_finite_difference_gradient_and_laplacian(lib, name):
    """testwf.py:149-217, 221-289: grad ln psi and lap psi / psi against central differences of
    testvalue -- no oracle involved."""
    import pyqmc_b200 as pq

    mol, mf, wf, _ = helpers.make_pair(name)
    N = 6
    np.random.seed(9)
    configs = pq.initial_guess(mol, N)
    wf.recompute(configs)
    delta = 1e-5
    ne = configs.configs.shape[1]
    for e in (0, ne - 1):
        pos = configs.configs[:, e].copy()
        grad, lap = wf.gradient_laplacian(e, configs.make_irreducible(e, pos))
        g2, val, _ = wf.gradient_value(e, configs.make_irreducible(e, pos))
        assert relerr(grad, g2) < 1e-10 and np.abs(val - 1).max() < 1e-12
        assert relerr(wf.gradient(e, configs.make_irreducible(e, pos)), grad) < 1e-10
        num_grad = np.zeros((3, N))
        num_lap = np.zeros(N)
        for d in range(3):
            plus, minus = pos.copy(), pos.copy()
            plus[:, d] += delta
            minus[:, d] -= delta
            rp = wf.testvalue(e, configs.make_irreducible(e, plus))[0]
            rm = wf.testvalue(e, configs.make_irreducible(e, minus))[0]
            num_grad[d] = (rp - rm) / (2 * delta)        # d(psi'/psi)/dx at psi'/psi = 1
            num_lap += (rp + rm - 2.0) / delta**2
        assert np.abs(num_grad - grad).max() < 1e-5 * max(1.0, np.abs(grad).max())
        assert np.abs(num_lap - lap).max() < 2e-4 * max(1.0, np.abs(lap).max())


def test_resident_recompute_between_blocks_is_bit_identical(lib, monkeypatch):
    """vmc() recomputes the wave function at the start of every block (mc.py:110).  When the walkers are the
    ones the previous device block returned, the driver recomputes from the device-resident coordinates
    (qmcb_recompute_resident) instead of uploading them again: same numbers, bit for bit, as the ordinary
    recompute, also after a parameter change and after the host array was modified between blocks."""
    import pyqmc_b200 as pq

    def run(modify):
        mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
        acc = pq.EnergyAccumulator(mol)
        np.random.seed(11)
        configs = pq.initial_guess(mol, 96)
        df1, configs = pq.vmc(wf, configs, tstep=0.5, nblocks=3, nsteps_per_block=4, accumulators={"energy": acc})
        wf.parameters["wf2bcoeff"][1, :] *= 1.01  # parameters change between runs; state must follow
        if modify:
            configs.configs[3, 2, :] += 0.125  # host array no longer equals the resident walkers
        df2, configs = pq.vmc(wf, configs, tstep=0.5, nblocks=2, nsteps_per_block=4, accumulators={"energy": acc})
        return df1, df2, configs.configs.copy(), wf.recompute(configs)

    for modify in (False, True):
        monkeypatch.delenv("QMCB_NO_RESIDENT_RECOMPUTE", raising=False)
        a = run(modify)
        monkeypatch.setenv("QMCB_NO_RESIDENT_RECOMPUTE", "1")
        b = run(modify)
        for k in a[0]:
            if "time" not in k:
                assert np.array_equal(a[0][k], b[0][k]) and np.array_equal(a[1][k], b[1][k]), k
        assert np.array_equal(a[2], b[2])
        assert np.array_equal(a[3][0], b[3][0]) and np.array_equal(a[3][1], b[3][1])


def test_resident_recompute_is_dropped_after_protocol_calls(lib, monkeypatch):
    """A protocol call that changes the device state between two blocks (here a masked updateinternals that
    moves some walkers' first electron, with the host array left as the block returned it) must make the next
    block fall back to the ordinary recompute from the HOST walkers, as the reference's vmc_worker would."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import mc

    def run(disable):
        if disable:
            monkeypatch.setenv("QMCB_NO_RESIDENT_RECOMPUTE", "1")
        else:
            monkeypatch.delenv("QMCB_NO_RESIDENT_RECOMPUTE", raising=False)
        mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
        acc = pq.EnergyAccumulator(mol)
        np.random.seed(5)
        configs = pq.initial_guess(mol, 40)
        _, configs = mc.vmc_block_device(wf, configs, 0.5, 3, {"energy": acc})
        epos = configs.electron(0)
        moved = pq.OpenElectron(epos.configs + 0.05, dist=epos.dist)
        mask = np.arange(40) % 3 == 0
        wf.updateinternals(0, moved, configs, mask=mask)  # device walkers now differ from the host array
        avg, configs = mc.vmc_block_device(wf, configs, 0.5, 3, {"energy": acc})
        return avg, configs.configs.copy()

    a, b = run(False), run(True)
    for k in a[0]:
        if "time" not in k:
            assert np.array_equal(a[0][k], b[0][k]), k
    assert np.array_equal(a[1], b[1])


def test_tmove_tables_keep_the_default_quadrature_when_naip_is_passed(lib):
    """EnergyAccumulator(naip=12): the energy uses 12 points per ECP atom, the T-move tables the default 6 / 12
    (the reference's nonlocal_tmoves calls compute_tmoves without naip, accumulators.py:80-81)."""
    import pyqmc_b200 as pq
    from oracle.local_energy import EnergyOracle

    mol, mf, wf, orc = helpers.make_pair("h2o", seed=1)
    np.random.seed(5)
    configs = pq.initial_guess(mol, 9)
    oc = helpers.to_oracle_walkers(configs)
    wf.recompute(configs)
    orc.recompute(oc)
    acc, oacc = pq.EnergyAccumulator(mol, naip=12), EnergyOracle(mol, naip=12)
    np.random.seed(6)
    en = acc(configs, wf)
    np.random.seed(6)
    oen = oacc(oc, orc)
    assert helpers.relerr(en["ecp"], oen["ecp"]) < 1e-10
    np.random.seed(7)
    tm = acc.nonlocal_tmoves(configs, wf, 2, 0.02)
    np.random.seed(7)
    otm = oacc.nonlocal_tmoves(oc, orc, 2, 0.02)
    assert tm["ratio"].shape == otm["ratio"].shape and tm["ratio"].shape[1] < 3 * 12
    assert helpers.relerr(tm["ratio"], otm["ratio"]) < 1e-10
    assert helpers.relerr(tm["weight"], otm["weight"]) < 1e-10
    np.random.seed(8)
    en2 = acc(configs, wf)  # back to the energy tables
    np.random.seed(8)
    assert helpers.relerr(en2["ecp"], oacc(oc, orc)["ecp"]) < 1e-10
