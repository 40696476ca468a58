"""CPU tests of the host logic around complex wave functions: when a Slater factor is complex (slater.py:212-216,
orbitals.py:160-165), and the parameter map against the reference's LinearTransform (accumulators.py:113-185) for
complex parameters."""
import os
import sys

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402


def test_dtype_follows_the_parameters_and_the_twist():
    import pyqmc_b200 as pq

    for name, want in (("h2o", float), ("h2o_cx", complex), ("ortho", float), ("ortho_twist", complex),
                       ("diamond211_twist", complex)):
        mol, mf, dets = helpers.make_system(name)
        wf = pq.Slater(mol, mf, determinants=dets)
        assert wf.dtype == want, name
        jast, _ = pq.generate_jastrow(mol)
        assert pq.MultiplyWF(wf, jast).dtype == want, name
    mol, mf, dets = helpers.make_system("h2o_md_cx")
    wf = pq.Slater(mol, mf, determinants=dets)
    assert wf.dtype == complex and np.iscomplexobj(wf.parameters["det_coeff"])


def test_orbitals_real_up_to_a_phase_stay_real():
    """complex128 coefficients whose columns are real up to a constant phase (what k-point mean-field objects hand out
    at Gamma) are rotated to real numbers; one genuinely complex column makes the whole wave function complex."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import systems

    mol, mf = systems.h2o_ccecp_pvtz()
    phases = np.exp(2j * np.pi * np.random.RandomState(0).rand(mf.mo_coeff.shape[-1]))
    rotated = systems.MF(mf.mo_coeff * phases[None, None, :], mf.mo_occ)
    wf = pq.Slater(mol, rotated)
    assert wf.dtype == float and not np.iscomplexobj(wf.parameters["mo_coeff_alpha"])
    assert np.allclose(np.abs(wf.parameters["mo_coeff_alpha"]), np.abs(np.asarray(mf.mo_coeff[0])[:, :4]))
    mixed = np.array(rotated.mo_coeff)
    mixed[0, :, 1] = mixed[0, :, 1] * (1 + 0.3j * np.arange(mixed.shape[1]) / mixed.shape[1])
    assert pq.Slater(mol, systems.MF(mixed, mf.mo_occ)).dtype == complex


@pytest.mark.skipif(not refload.available(), reason="staged reference (oracle/_ref) absent")
def test_parameter_map_equals_linear_transform_for_complex_parameters():
    refload.load()
    from pyqmc.observables.accumulators import LinearTransform

    from pyqmc_b200.sr import ParameterMap

    rng = np.random.RandomState(2)
    params = {"wf1det_coeff": rng.randn(5) + 1j * rng.randn(5), "wf1mo_coeff_alpha": rng.randn(6, 3) + 1j * rng.randn(6, 3),
              "wf2acoeff": rng.randn(3, 4, 2), "wf2bcoeff": rng.randn(4, 3)}
    to_opt = {k: rng.rand(*v.shape) > 0.4 for k, v in params.items()}
    to_opt["wf2bcoeff"][...] = False

    class WF:
        parameters = params

    mine, ref = ParameterMap(params, to_opt), LinearTransform(params, to_opt)
    assert mine.nparams == ref.nparams
    v_mine, v_ref = mine.serialize_parameters(params), ref.serialize_parameters(params)
    assert np.array_equal(v_mine, v_ref) and not np.iscomplexobj(v_mine)
    pgrad = {k: rng.randn(7, *v.shape) + (1j * rng.randn(7, *v.shape) if np.iscomplexobj(v) else 0) for k, v in params.items()}
    assert np.array_equal(mine.serialize_gradients(pgrad), np.asarray(ref.serialize_gradients(pgrad)))
    step = v_ref + 0.1 * rng.randn(len(v_ref))
    d_mine, d_ref = mine.deserialize(WF, step), ref.deserialize(WF, step)
    assert set(d_mine) == set(d_ref)
    for k in d_ref:
        assert np.array_equal(d_mine[k], d_ref[k]) and d_mine[k].dtype == d_ref[k].dtype, k


def test_block_averages_of_a_complex_block_follow_the_reference_rolling_mean():
    """Device-resident blocks of a complex wave function return 8 energy rows per step (6 real + Im ecp, Im total);
    the block dictionary must equal what vmc_worker accumulates from accumulator.avg (mc.py:139-147): complex ecp /
    total, real ke / ee / ei / grad2."""
    from types import SimpleNamespace

    from pyqmc_b200 import mc
    from pyqmc_b200.accumulators import KEYS

    rng = np.random.RandomState(2)
    nsteps, nconf, nelec = 3, 17, 4
    energy = rng.randn(nsteps, 8, nconf)
    nacc = rng.randint(0, nconf, size=(nsteps, nelec))
    got = mc._block_averages(SimpleNamespace(energy=energy, nacc=nacc), nsteps, nconf, nelec, "energy", object())
    for i, k in enumerate(KEYS):
        per_walker = energy[:, i].astype(complex) if k in ("ecp", "total") else energy[:, i]
        if k == "ecp":
            per_walker = per_walker + 1j * energy[:, 6]
        if k == "total":
            per_walker = per_walker + 1j * energy[:, 7]
        want = None
        for step in range(nsteps):  # the reference: block_avg[k] = res / nsteps, then += res / nsteps
            res = np.mean(per_walker[step], axis=0)
            want = res / nsteps if want is None else want + res / nsteps
        assert np.iscomplexobj(got["energy" + k]) == (k in ("ecp", "total")), k
        assert got["energy" + k] == want, k
    real = mc._block_averages(SimpleNamespace(energy=energy[:, :6], nacc=nacc), nsteps, nconf, nelec, "energy", object())
    assert all(not np.iscomplexobj(v) for v in real.values())
    assert real["energyke"] == got["energyke"] and real["acceptance"] == got["acceptance"]
