"""CPU tests of the host layer: the C-ABI library loads and exports every symbol include/qmcb200.h
declares (no compute calls without a GPU), host tables equal the oracle's, RNG draw order of the
device-resident driver equals what the reference loop consumes."""
import os
import re

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(lib):
    from pyqmc_b200 import _lib

    header = open(os.path.join(ROOT, "include", "qmcb200.h")).read()
    declared = set(re.findall(r"\b(qmcb_[a-z_0-9]+)\s*\(", header))
    declared.discard("qmcb_ctx")
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in qmcb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_device(lib):
    """Creating a context without a CUDA device must fail loudly (never a silent CPU path)."""
    import ctypes

    if lib.qmcb_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.qmcb_create(0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.qmcb_last_error()
    import pyqmc_b200 as pq
    from pyqmc_b200._lib import QmcbError

    mol, mf, _ = helpers.make_system("he")
    wf = pq.Slater(mol, mf)
    with pytest.raises(QmcbError):
        wf.recompute(pq.OpenConfigs(np.zeros((2, 2, 3))))


def test_product_package_does_not_import_oracle():
    import subprocess
    import sys

    code = "import sys; import pyqmc_b200, pyqmc_b200.mc, pyqmc_b200.accumulators; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for fn in os.listdir(os.path.join(ROOT, "pyqmc_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "pyqmc_b200", fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


@pytest.mark.parametrize("name", ["he", "h2o", "c2"])
def test_shell_tables_equal_oracle(name):
    from oracle.gto import BasisTable
    from pyqmc_b200 import basis

    mol, mf, _ = helpers.make_system(name)
    t, o = basis.shell_tables(mol), BasisTable(mol)
    assert t["nao"] == o.nao
    assert np.array_equal(t["shell_atom"], o.shell_atom) and np.array_equal(t["shell_l"], o.shell_l)
    assert np.array_equal(t["prim_off"], o.prim_off)
    assert np.array_equal(t["exps"], o.exps)
    assert np.abs(t["coefs"] - o.coefs).max() < 1e-15 * np.abs(o.coefs).max()


@pytest.mark.parametrize("naip", [6, 12, 18, 26, 32, 50])
def test_quadrature_equals_oracle_and_integrates(naip):
    from oracle import local_energy
    from pyqmc_b200 import quadrature

    p, w = quadrature.grid(naip)
    po, wo = local_energy.quadrature(naip)
    assert np.array_equal(p, po) and np.array_equal(w, wo)
    assert p.shape == (naip, 3) and abs(w.sum() - 1) < 1e-14
    assert np.abs(np.linalg.norm(p, axis=1) - 1).max() < 1e-14
    assert np.abs(w @ p).max() < 1e-14  # l = 1 integrates to zero
    assert np.abs(w @ (p[:, 2] ** 2) - 1 / 3) < 1e-14  # <z^2> on the sphere


def test_flatten_ecp_column_order():
    from pyqmc_b200.accumulators import flatten_ecp

    mol, mf, _ = helpers.make_system("c2")
    t = flatten_ecp(mol)
    assert list(t["ecp_atom"]) == [0, 1] and list(t["naip"]) == [12, 12]
    assert list(t["chan_off"]) == [0, 3, 6]
    # channel columns: l=0, l=1, then the local channel (3 terms r^-1, r^0, r^1)
    nterm = np.diff(t["term_off"])
    assert list(nterm) == [1, 1, 3, 1, 1, 3]
    assert sorted(t["power"][t["term_off"][2]:t["term_off"][3]]) == [-1, 0, 1]


def test_block_variates_follow_reference_consumption_order():
    """draw_block_variates must leave the global stream exactly where vmc_worker + the energy
    accumulator would (mc.py:119,132; eval_ecp.py:145,263)."""
    import scipy.spatial.transform
    from pyqmc_b200 import mc
    from pyqmc_b200.accumulators import EnergyAccumulator

    mol, mf, _ = helpers.make_system("h2o")
    acc = EnergyAccumulator(mol)
    N, ne, nsteps, tstep = 5, 8, 2, 0.5
    np.random.seed(9)
    g, u, eu, er = mc.draw_block_variates(N, ne, tstep, nsteps, acc)
    after = np.random.random()
    np.random.seed(9)
    for step in range(nsteps):
        for e in range(ne):
            assert np.array_equal(g[step, e], np.random.normal(scale=np.sqrt(tstep), size=(N, 3)))
            assert np.array_equal(u[step, e], np.random.rand(N))
        for e in range(ne):
            for a in range(3):
                assert np.array_equal(eu[step, e, a], np.random.random(size=N))
                assert np.array_equal(er[step, e, a], scipy.spatial.transform.Rotation.random().as_matrix())
    assert after == np.random.random()


def test_parameters_view_and_factories():
    import pyqmc_b200 as pq

    mol, mf, _ = helpers.make_system("open")
    wf, to_opt = pq.generate_wf(mol, mf)
    keys = list(wf.parameters.keys())
    assert keys == ["wf1det_coeff", "wf1mo_coeff_alpha", "wf1mo_coeff_beta", "wf2bcoeff", "wf2acoeff"]
    assert set(to_opt) == set(keys) - {"wf1mo_coeff_alpha", "wf1mo_coeff_beta"}
    assert wf.parameters["wf1mo_coeff_alpha"].shape == (57, 3) and wf.parameters["wf1mo_coeff_beta"].shape == (57, 1)
    # hydrogens carry no ECP in this system -> electron-ion cusp term (wftools.py:118-146)
    assert wf.parameters["wf2acoeff"].shape == (3, 5, 2)
    assert np.array_equal(wf.parameters["wf2acoeff"][:, 0, 0], [0.0, 1.0, 1.0])
    assert np.array_equal(wf.parameters["wf2bcoeff"][0], [-0.25, -0.5, -0.25])
    wf.parameters["wf2bcoeff"] = np.ones((4, 3))
    assert wf.wf_factors[1].parameters["bcoeff"][2, 1] == 1.0
    assert wf.dtype == float


def test_initial_guess_matches_oracle_stream():
    import pyqmc_b200 as pq
    from oracle import vmc_driver

    for name in ("h2o", "open", "he"):
        mol, mf, _ = helpers.make_system(name)
        np.random.seed(4)
        a = pq.initial_guess(mol, 11).configs
        np.random.seed(4)
        b = vmc_driver.initial_guess(mol, 11).configs
        assert np.array_equal(a, b)


@pytest.mark.parametrize("producer", [False, True])
@pytest.mark.parametrize("seed", range(12))
def test_native_rng_is_bit_identical_to_numpy_legacy_stream(lib, seed, producer, monkeypatch):
    """qmcb_rng_vmc_block (csrc/legacy_rng.cpp) vs the numpy/scipy calls of the reference loop:
    same variates, same final MT19937 position and Gaussian cache, incl. odd sizes and a cached
    Gaussian carried in; with the recurrence inline (default) and on its producer thread."""
    from pyqmc_b200 import mc
    from pyqmc_b200.accumulators import EnergyAccumulator

    if producer:
        monkeypatch.setenv("QMCB_RNG_PRODUCER", "1")
        monkeypatch.setenv("QMCB_RNG_THREADS", "3")
    else:
        monkeypatch.delenv("QMCB_RNG_PRODUCER", raising=False)
    mol, mf, _ = helpers.make_system("c2" if seed % 2 else "h2o")
    acc = EnergyAccumulator(mol) if seed % 3 else None
    N = 1 + 37 * seed % 23
    np.random.seed(seed)
    if seed % 4 == 0:
        np.random.normal()  # leaves a cached Gaussian behind
    st0 = np.random.get_state()
    a = mc.draw_block_variates(N, 8, 0.3, 3, acc, native=False)
    sa = np.random.get_state()
    np.random.set_state(st0)
    b = mc.draw_block_variates(N, 8, 0.3, 3, acc, native=True)
    sb = np.random.get_state()
    for x, y in zip(a, b):
        assert (x is None and y is None) or np.array_equal(x, y)
    assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]


@pytest.mark.parametrize("producer", [False, True])
def test_native_rng_at_bench_size_matches_numpy(lib, producer, monkeypatch):
    """The C2 block shape of bench.py (4096 walkers, 8 electrons, 3 ECP atoms; ~1600 MT19937 state blocks for
    two steps): every variate and the final generator state equal numpy's, with 1 and 5 host threads, with the
    recurrence inline and on the producer thread, starting mid-block with a cached Gaussian."""
    from pyqmc_b200 import mc
    from pyqmc_b200.accumulators import EnergyAccumulator

    mol, mf, _ = helpers.make_system("h2o")
    acc = EnergyAccumulator(mol)
    if producer:
        monkeypatch.setenv("QMCB_RNG_PRODUCER", "1")
    else:
        monkeypatch.delenv("QMCB_RNG_PRODUCER", raising=False)
    np.random.seed(77)
    np.random.normal(size=301)  # odd count: mid-block position and a cached second Gaussian
    st0 = np.random.get_state()
    a = mc.draw_block_variates(4096, 8, 0.5, 2, acc, native=False)
    sa = np.random.get_state()
    for threads in ("1", "5"):
        monkeypatch.setenv("QMCB_RNG_THREADS", threads)
        np.random.set_state(st0)
        b = mc.draw_block_variates(4096, 8, 0.5, 2, acc, native=True)
        sb = np.random.get_state()
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]


def test_native_draw_program_is_bit_identical_to_numpy_for_the_dmc_block():
    """qmcb_rng_program (csrc/legacy_rng.cpp) on the DMC draw order (dmc.py:150-198): same numbers as
    the numpy / scipy calls, same final state of the global legacy generator, with a cached Gaussian
    carried in."""
    import numpy as np

    from pyqmc_b200 import dmc, systems
    from pyqmc_b200.accumulators import EnergyAccumulator

    mol, mf = systems.h2o_ccecp_pvtz()
    acc = EnergyAccumulator(mol)
    for nconf in (5, 48):
        np.random.seed(3)
        np.random.normal(size=1)
        a = dmc.draw_dmc_block_variates(nconf, 8, 0.02, 2, acc, native=True)
        ta = (np.random.rand(), np.random.normal())
        np.random.seed(3)
        np.random.normal(size=1)
        b = dmc.draw_dmc_block_variates(nconf, 8, 0.02, 2, acc, native=False)
        tb = (np.random.rand(), np.random.normal())
        assert ta == tb
        for k in a:
            assert np.array_equal(a[k], b[k]), k
