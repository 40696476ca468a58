"""The reference's OWN drivers and test harness, unchanged, over the device objects.

``pyqmc.method.mc.vmc`` (mc.py:176-274), ``pyqmc.method.dmc.rundmc`` (dmc.py:413-591, ``hdf_file=None``) and
the ``pyqmc/wf/testwf.py`` harness are imported from the staged, unmodified reference package
(``oracle/_ref``, see ``oracle/stage_reference.py``) and handed ``pyqmc_b200`` wave functions and
accumulators together with the reference's own ``OpenConfigs`` / ``PeriodicConfigs`` containers: this is the
"pyqmc.method.vmc/dmc drive it unchanged" claim of the drop-in boundary.  Outputs are
compared with golden vectors the same drivers produced with the reference's wave functions
(``tests/golden/make_golden.py``): accept masks bit for bit, energies to 1e-10 relative.
"""
import os
import sys

import numpy as np
import pytest

import golden_replay
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

TOL = 1e-10
EWALD = {"ewald_gmax": 10}
needs_reference = pytest.mark.skipif(not refload.available(), reason="staged reference (oracle/_ref) absent")


def _ref_configs(mol, configs, wrap=None):
    refload.load()
    import pyqmc.configurations.coord as coord

    if hasattr(mol, "a"):
        c = coord.PeriodicConfigs(configs.copy(), mol.lattice_vectors())
        if wrap is not None:
            c.wrap[...] = wrap
        return c
    return coord.OpenConfigs(configs.copy())


def _spy_accepts(wf):
    accepts = []
    orig = wf.updateinternals

    def spy(e, epos, cfg, mask=None, saved_values=None):
        accepts.append(np.array(mask))
        return orig(e, epos, cfg, mask=mask, saved_values=saved_values)

    wf.updateinternals = spy
    return accepts


def run_reference_vmc(name, wf, make_energy):
    """The VMC segment of the golden file through the reference's mc.vmc; returns nothing, asserts."""
    refload.load()
    import pyqmc.method.mc as refmc

    data = golden_replay.load(name)
    mol = helpers.make_system(name)[0]
    periodic = hasattr(mol, "a")
    configs = _ref_configs(mol, data["configs1"], data.get("wrap1"))
    ne = configs.configs.shape[1]
    accepts = _spy_accepts(wf)
    np.random.seed(31)
    df, configs = refmc.vmc(wf, configs, tstep=0.5, nblocks=2, nsteps_per_block=3,
                            accumulators={"energy": make_energy(mol, **(EWALD if periodic else {}))})
    got = np.array(accepts).reshape(2, 3, ne, -1)
    assert np.array_equal(got, data["vmc_accept"]), "accept masks differ from the reference's own run"
    assert np.array_equal(df["acceptance"], data["vmc_acceptance"])
    assert np.abs(configs.configs - data["vmc_configs"]).max() < 1e-10
    if periodic:
        assert np.array_equal(configs.wrap, data["vmc_wrap"])
    for k in ("energytotal", "energyke", "energyecp", "energyee", "energyei", "energygrad2"):
        assert helpers.relerr(df[k], data["vmc_" + k]) < TOL, k


def run_reference_rundmc(name, wf, make_energy):
    refload.load()
    import pyqmc.method.dmc as refdmc

    data = golden_replay.load("rundmc_" + name)
    mol = helpers.make_system(name)[0]
    configs = _ref_configs(mol, data["configs0"])
    np.random.seed(52)
    df, configs, weights = refdmc.rundmc(wf, configs, tstep=0.02, nblocks=3, nsteps_per_block=2, vmc_warmup=2,
                                         accumulators={"energy": make_energy(mol)})
    assert np.abs(configs.configs - data["configs"]).max() < 1e-10
    assert helpers.relerr(weights, data["weights"]) < TOL
    for k in df:
        ref = data["df_" + k]
        if np.issubdtype(ref.dtype, np.integer):
            assert np.array_equal(df[k], ref), k
        else:
            assert helpers.relerr(df[k], ref) < TOL, k


def run_testwf_harness(name, wf, mol, pgradient=True):
    refload.load()
    import pyqmc.method.mc as refmc
    from pyqmc.wf import testwf

    np.random.seed(5)
    # test_wf_gradient_value indexes a per-walker array by electron (testwf.py:264,276): needs nconf >= nelec
    configs = refmc.initial_guess(mol, max(11, int(sum(mol.nelec)) + 1))
    if not name.endswith("_3b"):  # the reference's vmc leaves a three-body cache stale: harness only
        _, configs = refmc.vmc(wf, configs, nblocks=1, nsteps=2, tstep=1)
    for k, item in testwf.test_updateinternals(wf, configs).items():
        assert item < 1e-5, (k, item)
    wf.recompute(configs)
    np.random.seed(6)
    testwf.test_mask(wf, 0, configs.electron(0))
    if wf.dtype != complex:
        # these two harness functions store ratios in float arrays (testwf.py:50): the reference's own complex wave
        # functions fail them too, so they are run for real wave functions only
        testwf.test_testvalue_many(wf, configs)
        aux = configs.make_irreducible(0, configs.configs[:, 0][:, None, :] + 0.2 * np.random.randn(len(configs.configs), 6, 3))  # noqa: E501
        aux_configs = _ref_configs(mol, aux.configs, getattr(aux, "wrap", None))
        testwf.test_testvalue_aux(wf, configs, aux_configs)
    err = [testwf.test_wf_gradient(wf, configs, delta) for delta in (1e-4, 1e-5, 1e-6)]
    assert min(err) < 1e-5, err
    if pgradient:
        err = [testwf.test_wf_pgradient(wf, configs, delta) for delta in (1e-5, 1e-6)]
        assert min(err) < 1e-5, err
    for func in (testwf.test_wf_gradient_value, testwf.test_wf_gradient_laplacian):
        for k, v in func(wf, configs).items():
            assert v < 1e-10, (func.__name__, k, v)


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("name", ["he", "h2o", "open", "c2", "h2o_md", "h2o_cas", "ortho", "rotcubic", "diamond211"])
def test_reference_vmc_drives_device_objects(lib, name):
    """mc.vmc of the reference over device wf + device accumulator == the reference-only golden run."""
    import pyqmc_b200 as pq

    run_reference_vmc(name, helpers.make_pair(name, seed=1)[2], pq.EnergyAccumulator)


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("name", ["h2o", "c2"])
def test_reference_rundmc_drives_device_objects(lib, name):
    """dmc.rundmc of the reference (VMC warm-up, T-moves, weights, branching, e_trial feedback)."""
    import pyqmc_b200 as pq

    run_reference_rundmc(name, helpers.make_pair(name, seed=1)[2], pq.EnergyAccumulator)


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("name", ["h2o", "open", "h2o_md", "h2o_3b", "diamond211"])
def test_reference_testwf_harness(lib, name):
    """pyqmc/wf/testwf.py on the device objects with the thresholds of tests/unit/test_wf_derivatives.py:40-72
    (1e-5 for finite differences and cache consistency, 1e-10 for the combined calls)."""
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    run_testwf_harness(name, wf, mol, pgradient=name in ("h2o", "h2o_md"))


@pytest.mark.gpu
@needs_reference
def test_reference_accumulator_contract(lib):
    """tests/unit/test_accumulators.py:72-83: keys() == shapes().keys() == avg().keys(), shapes match."""
    import pyqmc_b200 as pq

    refload.load()
    import pyqmc.method.mc as refmc

    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    to_opt = {k: np.ones(np.shape(v), dtype=bool) for k, v in wf.parameters.items() if "mo_coeff" not in k}
    accumulators = {"pgrad": pq.gradient_generator(mol, wf, to_opt)}
    accumulators["energy"] = accumulators["pgrad"].enacc
    np.random.seed(2)
    configs = refmc.initial_guess(mol, 100)
    wf.recompute(configs)
    for k, acc in accumulators.items():
        shapes, keys = acc.shapes(), acc.keys()
        assert shapes.keys() == keys
        avg = acc.avg(configs, wf)
        assert avg.keys() == keys
        for ka in keys:
            assert shapes[ka] == np.shape(avg[ka]), (ka, np.shape(avg[ka]))


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("name", ["ortho_3b", "diamond211_3b"])
def test_periodic_three_body_under_the_reference_driver(lib, name):
    """Periodic Slater x Jastrow x three-body: the reference's mc.vmc over the device objects (the per-call protocol)
    against the oracle loop over the oracle objects -- the reference's own three-body factor goes stale in driver
    order, so the oracle is the checker here."""
    import pyqmc_b200 as pq
    from oracle import vmc_driver
    from oracle.local_energy import EnergyOracle

    refload.load()
    import pyqmc.method.mc as refmc

    mol, mf, wf, orc = helpers.make_pair(name, seed=1)
    np.random.seed(4)
    configs = pq.initial_guess(mol, 9)
    oconfigs = helpers.to_oracle_walkers(configs)
    accepts = _spy_accepts(wf)
    np.random.seed(8)
    df, configs = refmc.vmc(wf, configs, nblocks=1, nsteps_per_block=2, accumulators={"energy": pq.EnergyAccumulator(mol, **EWALD)})
    record = []
    np.random.seed(8)
    odf, oconfigs = vmc_driver.vmc(orc, oconfigs, nblocks=1, nsteps_per_block=2, accumulators={"energy": EnergyOracle(mol, **EWALD)},
                                   record=record)
    assert len(accepts) == len(record) > 0, "the protocol loop of the reference driver must have run"
    assert all(np.array_equal(a, r["accept"]) for a, r in zip(accepts, record))
    assert np.abs(configs.configs - oconfigs.configs).max() < 1e-10 and np.array_equal(configs.wrap, oconfigs.wrap)
    for k in ("energytotal", "energyke", "energyecp", "energyee"):
        assert helpers.relerr(df[k], odf[k]) < TOL, k


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("name", ["ortho", "ortho_md", "h2o_md_3b"])
def test_device_resident_rundmc_equals_the_reference_driver_over_the_protocol(lib, name):
    """pyqmc_b200.rundmc (device-resident warm-up, propagation and variates, host branching) vs the reference's rundmc
    driving a second copy of the same device objects call by call: periodic cells and multi-determinant / three-body
    wave functions take the general DMC kernels (k_pbc_move_general<16, true>, k_vmc_move_coop<16, true>)."""
    import pyqmc_b200 as pq

    refload.load()
    import pyqmc.method.dmc as refdmc

    kw = dict(tstep=0.05, nblocks=3, nsteps_per_block=2, vmc_warmup=2)
    out = []
    for driver in (pq.rundmc, refdmc.rundmc):
        mol, mf, wf, _ = helpers.make_pair(name, seed=1)
        ekw = EWALD if hasattr(mol, "a") else {}
        np.random.seed(17)
        configs = pq.initial_guess(mol, 14)
        np.random.seed(18)
        out.append(driver(wf, configs, accumulators={"energy": pq.EnergyAccumulator(mol, **ekw)}, **kw))
    (df1, c1, w1), (df2, c2, w2) = out
    assert np.abs(c1.configs - c2.configs).max() < 1e-9
    if hasattr(c1, "wrap"):
        assert np.array_equal(c1.wrap, c2.wrap)
    assert helpers.relerr(w1, w2) < 1e-9
    for k in ("energytotal", "energyke", "energyecp", "weight", "acceptance", "tmove_acceptance", "e_trial", "e_est"):
        assert helpers.relerr(df1[k], df2[k]) < 1e-9, k


# ---- CPU self-check of the harness above: the same functions over the REFERENCE's own wave functions must
# reproduce the golden files (run in the build container; proves the test plumbing, not the product) ----------
@needs_reference
@pytest.mark.parametrize("name", ["he", "ortho"])
def test_harness_selfcheck_reference_objects(name):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden

    refload.load()
    from pyqmc.observables.accumulators import EnergyAccumulator

    mol, wf = make_golden.build_reference(name)
    run_reference_vmc(name, wf, EnergyAccumulator)
