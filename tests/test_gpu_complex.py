"""GPU parity for COMPLEX wave functions (csrc/cplx.cuh): complex orbital / determinant coefficients on open
boundaries (``h2o_cx``, ``h2o_md_cx``) and general twists -- complex Bloch phases, wrap phase exp(i k.R) -- on
periodic cells (``ortho_twist``, ``diamond211_twist``), against golden vectors the unmodified reference produced
for the same systems (tests/golden/make_golden.py): every protocol call, the internal arrays, ``pgradient``, the
energy accumulator (complex ECP / total), T-move tables, and a VMC run under the reference's own ``mc.vmc`` with
accept masks bit for bit.  Reference behaviour: slater.py:212-216, orbitals.py:38-39,160-165, energy.py:63-64,
eval_ecp.py:26."""
import numpy as np
import pytest

import golden_replay
import helpers
from test_gpu_reference_drivers import _ref_configs, _spy_accepts, needs_reference, refload

pytestmark = pytest.mark.gpu

OPEN = ["h2o_cx", "h2o_md_cx"]
PERIODIC = ["ortho_twist", "diamond211_twist"]
EWALD_GMAX = 10


def _configs(data, mol, key="configs0", wkey="wrap0"):
    import pyqmc_b200 as pq

    if not hasattr(mol, "a"):
        return pq.OpenConfigs(data[key].copy())
    c = pq.PeriodicConfigs(data[key].copy(), mol.lattice_vectors())
    c.configs = data[key].copy()
    c.wrap = data[wkey].copy()
    return c


def _energy(mol):
    import pyqmc_b200 as pq

    return pq.EnergyAccumulator(mol, **({"ewald_gmax": EWALD_GMAX} if hasattr(mol, "a") else {}))


def reference_vmc(wf, configs, accumulators):
    """The reference's own driver over the device objects (the per-call protocol)."""
    if not refload.available():
        pytest.skip("staged reference (oracle/_ref) absent")
    refload.load()
    import pyqmc.method.mc as refmc

    mol = wf.wf_factors[0]._mol
    rc = _ref_configs(mol, configs.configs, getattr(configs, "wrap", None))
    ne = rc.configs.shape[1]
    accepts = _spy_accepts(wf)
    df, rc = refmc.vmc(wf, rc, tstep=0.5, nblocks=2, nsteps_per_block=3, accumulators=accumulators)
    del wf.updateinternals
    return df, rc, np.array(accepts).reshape(2, 3, ne, -1)


def check_internal(wf, data):
    sl, ja = wf.wf_factors[:2]
    for s in (0, 1):
        assert np.iscomplexobj(sl._inverse[s])
        assert helpers.relerr(sl._inverse[s], data[f"inverse{s}"]) < 1e-9
        assert golden_replay.same_sign(sl._dets[s][0], data[f"dets{s}"][0])
        assert np.abs(sl._dets[s][1] - data[f"dets{s}"][1]).max() < 1e-10
    assert helpers.relerr(ja._a_partial, data["a_partial"]) < 1e-10
    assert helpers.relerr(ja._b_partial, data["b_partial"]) < 1e-10
    pg = wf.pgradient()
    for k in ("wf1det_coeff", "wf1mo_coeff_alpha", "wf1mo_coeff_beta", "wf2acoeff", "wf2bcoeff"):
        assert pg[k].shape == data["pgrad_" + k].shape, k
        assert np.iscomplexobj(pg[k]) == np.iscomplexobj(data["pgrad_" + k]), k
        assert helpers.relerr(pg[k], data["pgrad_" + k]) < 1e-9, k


@pytest.mark.parametrize("name", OPEN + PERIODIC)
def test_cuda_reproduces_reference_golden_complex(lib, name):
    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    assert wf.dtype == complex and wf.wf_factors[0].dtype == complex
    assert np.array_equal(wf.parameters["wf2acoeff"], data["acoeff"])
    golden_replay.replay(data, wf, _configs(data, mol), lambda: _energy(mol), reference_vmc, check_internal)


def device_vmc(wf, configs, accumulators):
    """Device-resident blocks of a complex wave function (k_cx_chain around the query kernels, qmcb_vmc_block)."""
    from pyqmc_b200 import mc

    rows, accepts = [], []
    for block in range(2):
        avg, configs, data = mc.vmc_block_device(wf, configs, 0.5, 3, accumulators, return_walker_data=True)
        rows.append(avg)
        accepts.append(data["accept"])
    df = {k: np.asarray([r[k] for r in rows]) for k in rows[0]}
    return df, configs, np.array(accepts)


@pytest.mark.parametrize("name", OPEN + PERIODIC)
def test_complex_device_resident_block_reproduces_reference_golden(lib, name):
    """The same golden replay with the VMC segment run as device-resident blocks: accept masks of the reference's own
    run bit for bit, walkers, wrap vectors, complex block energies."""
    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    golden_replay.replay(data, wf, _configs(data, mol), lambda: _energy(mol), device_vmc, check_internal)


@pytest.mark.parametrize("name", ["h2o_cx", "diamond211_twist"])
def test_complex_vmc_driver_runs_device_resident(lib, name):
    """pyqmc_b200.vmc on a complex wave function no longer delegates: block rows equal the reference driver's over the
    protocol calls (same seed), and kernels were launched without protocol calls in between."""
    import pyqmc_b200 as pq

    if not refload.available():
        pytest.skip("staged reference (oracle/_ref) absent")
    refload.load()
    import pyqmc.method.mc as refmc

    out = []
    for driver in (pq.vmc, refmc.vmc):
        mol, mf, wf, _ = helpers.make_pair(name, seed=1)
        np.random.seed(9)
        configs = pq.initial_guess(mol, 20)
        calls = _spy_accepts(wf)
        np.random.seed(10)
        df, configs = driver(wf, configs, nblocks=2, nsteps_per_block=2, accumulators={"energy": _energy(mol)})
        out.append((df, configs, len(calls)))
    (df1, c1, n1), (df2, c2, n2) = out
    assert n1 == 0 and n2 > 0, "pyqmc_b200.vmc must not go through updateinternals calls"
    assert np.abs(c1.configs - c2.configs).max() < 1e-10
    if hasattr(c1, "wrap"):
        assert np.array_equal(c1.wrap, c2.wrap)
    assert np.array_equal(df1["acceptance"], df2["acceptance"])
    for k in ("energytotal", "energyke", "energyecp", "energyee", "energyei", "energygrad2"):
        assert np.iscomplexobj(df1[k]) == np.iscomplexobj(df2[k]), k
        assert helpers.relerr(df1[k], df2[k]) < 1e-10, k


@pytest.mark.parametrize("name", OPEN + PERIODIC)
def test_slater_alone_and_unfused_product_complex(lib, name):
    """The Slater factor on its own context, and an unfused product (factors on separate contexts, combined on the
    host as multiplywf.py:116-129 does), give the fused result."""
    import pyqmc_b200 as pq

    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = _configs(data, mol)
    _, _, dets = helpers.make_system(name)
    slater = pq.Slater(mol, mf, determinants=dets)
    jast = wf.wf_factors[1]
    wf.recompute(configs)
    s1, l1 = slater.recompute(configs)
    assert np.iscomplexobj(s1) and np.abs(np.abs(s1) - 1).max() < 1e-12
    e = int(data["elist"][1])
    ep = configs.make_irreducible(e, data["q1_newpos"].copy())
    g, lap = wf.gradient_laplacian(e, ep)
    gs, ls = slater.gradient_laplacian(e, ep)
    gj, lj = jast.gradient_laplacian(e, ep)
    assert not np.iscomplexobj(gj)
    assert helpers.relerr(gs + gj, g) < 1e-12
    assert helpers.relerr(ls + lj + 2 * np.sum(gs * gj, axis=0), lap) < 1e-11
    v, _ = wf.testvalue(e, ep)
    vs, _ = slater.testvalue(e, ep)
    vj, _ = jast.testvalue(e, ep)
    assert helpers.relerr(vs * vj, v) < 1e-12


@pytest.mark.parametrize("name", ["h2o_cx", "ortho_twist"])
def test_tmoves_complex(lib, name):
    """compute_tmoves (eval_ecp.py:43-80): complex ratios, real weights."""
    data = golden_replay.load(name)
    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    configs = _configs(data, mol, "configs1", "wrap1")
    wf.recompute(configs)
    np.random.seed(22)
    tm = _energy(mol).nonlocal_tmoves(configs, wf, int(data["elist"][-1]), 0.02)
    assert np.iscomplexobj(tm["ratio"])
    assert helpers.relerr(tm["ratio"], data["tmove_ratio"]) < 1e-10
    assert helpers.relerr(tm["weight"], data["tmove_weight"]) < 1e-10
    assert np.abs(tm["configs"].configs - data["tmove_configs"]).max() < 1e-10


@needs_reference
@pytest.mark.parametrize("name", ["h2o_cx", "ortho_twist"])
def test_reference_harness_over_complex_device_wf(lib, name):
    """pyqmc/wf/testwf.py (finite-difference gradient / Laplacian / parameter-gradient checks, mask and
    updateinternals consistency) over the complex device objects, thresholds of tests/unit/test_wf_derivatives.py."""
    from test_gpu_reference_drivers import run_testwf_harness

    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    run_testwf_harness(name, wf, mol)


def test_twist_phase_across_the_cell_boundary(lib):
    """tests/integration/test_twist.py of the reference: moving every electron by lattice vectors multiplies the
    wave function by exp(i k.shift) and leaves the local energy unchanged."""
    import pyqmc_b200 as pq

    mol, mf, wf, _ = helpers.make_pair("ortho_twist", seed=1)
    kpt = np.asarray(mf.kpts)[0]
    np.random.seed(3)
    coords = pq.initial_guess(mol, 6)
    coords.wrap[...] = 0  # the reference's test re-creates the walkers inside the cell with zero wrap vectors
    L = np.random.RandomState(8).randint(10, size=coords.configs.shape) - 5
    shift = L @ mol.lattice_vectors()
    phase = np.exp(1j * np.einsum("ijk,k->ij", shift, kpt))
    moved = pq.PeriodicConfigs(coords.configs + shift, mol.lattice_vectors())
    assert np.abs(moved.configs - coords.configs).max() < 1e-10
    p0, v0 = wf.recompute(coords)
    np.random.seed(0)
    e0 = _energy(mol)(coords, wf)
    r = wf.testvalue(0, moved.electron(0))[0]
    assert np.abs(r - phase[:, 0]).max() < 1e-8
    p1, v1 = wf.recompute(moved)
    np.random.seed(0)
    e1 = _energy(mol)(moved, wf)
    assert np.abs(p0 * phase.prod(axis=1) - p1).max() < 1e-9
    assert np.abs(v0 - v1).max() < 1e-9
    for k in e0:
        assert np.abs(e0[k] - e1[k]).max() < 1e-8 * max(1.0, np.abs(e0[k]).max()), k
