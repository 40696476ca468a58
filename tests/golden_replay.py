"""Replays the call sequence recorded in tests/golden/<system>.npz (generated from the reference
by tests/golden/make_golden.py) on any wave-function implementation and compares."""
import os

import numpy as np

from helpers import relerr

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))


def _close(a, b, what, tol=TOL):
    err = relerr(a, b)
    assert err < tol, f"{what}: relative error {err:.3e}"


def same_sign(s, ref):
    """Signs of real wave functions are compared exactly, unit phases of complex ones to TOL."""
    if np.iscomplexobj(ref):
        return np.iscomplexobj(s) and np.abs(s - ref).max() < TOL
    return (not np.iscomplexobj(s)) and np.array_equal(s, ref)


def replay(data, wf, configs, make_energy, vmc_fn, check_internal=None):
    """wf: implementation under test; configs: walker container holding data['configs0'];
    make_energy(): energy accumulator; vmc_fn(wf, configs, accumulators) -> (df, configs, accepts)."""
    N = configs.configs.shape[0]
    ne = configs.configs.shape[1]
    s, l = wf.recompute(configs)
    assert same_sign(s, data["recompute_sign"])
    assert np.abs(l - data["recompute_log"]).max() < TOL * max(1.0, np.abs(l).max())
    for i, e in enumerate(data["elist"]):
        e = int(e)
        newpos, mask, aux = data[f"q{i}_newpos"], data[f"q{i}_mask"], data[f"q{i}_aux"]
        ep = configs.make_irreducible(e, newpos.copy())
        _close(wf.gradient(e, ep), data[f"q{i}_gradient"], "gradient")
        g, v, saved = wf.gradient_value(e, ep)
        _close(g, data[f"q{i}_gv_grad"], "gradient_value grad")
        _close(v, data[f"q{i}_gv_val"], "gradient_value val")
        g, lap = wf.gradient_laplacian(e, ep)
        _close(g, data[f"q{i}_gl_grad"], "gradient_laplacian grad")
        _close(lap, data[f"q{i}_gl_lap"], "gradient_laplacian lap")
        _close(wf.testvalue(e, ep)[0], data[f"q{i}_testvalue"], "testvalue")
        _close(wf.testvalue(e, ep, mask)[0], data[f"q{i}_testvalue_mask"], "testvalue mask")
        _close(wf.testvalue(e, configs.make_irreducible(e, aux.copy()), mask)[0], data[f"q{i}_testvalue_aux"],
               "testvalue aux")
        _close(wf.testvalue_many(np.arange(ne), ep), data[f"q{i}_testvalue_many"], "testvalue_many")
        g, v, saved = wf.gradient_value(e, ep)
        wf.updateinternals(e, ep, configs, mask=mask, saved_values=saved)
        configs.move(e, ep, mask)
        s, l = wf.value()
        assert same_sign(s, data[f"q{i}_value_sign"])
        assert np.abs(l - data[f"q{i}_value_log"]).max() < TOL * max(1.0, np.abs(l).max())
    assert np.abs(configs.configs - data["configs1"]).max() == 0.0
    if "wrap1" in data:
        assert np.array_equal(configs.wrap, data["wrap1"])
    if check_internal is not None:
        check_internal(wf, data)
    np.random.seed(21)
    en = make_energy()(configs, wf)
    for k in ("ke", "ee", "ei", "ecp", "grad2", "total"):
        assert np.iscomplexobj(en[k]) == np.iscomplexobj(data["energy_" + k]), k
        _close(en[k], data["energy_" + k], "energy " + k)
    if "vmc_accept" not in data:
        return
    np.random.seed(31)
    df, configs, accepts = vmc_fn(wf, configs, {"energy": make_energy()})
    if accepts is not None:
        assert np.array_equal(accepts, data["vmc_accept"]), "accept masks differ from the reference"
    assert np.array_equal(df["acceptance"], data["vmc_acceptance"])
    assert np.abs(configs.configs - data["vmc_configs"]).max() < TOL
    if "vmc_wrap" in data:
        assert np.array_equal(configs.wrap, data["vmc_wrap"])
    for k in ("energytotal", "energyke", "energyecp", "energyee", "energyei", "energygrad2"):
        assert np.abs(df[k] - data["vmc_" + k]).max() <= TOL * max(1.0, np.abs(data["vmc_" + k]).max()), k


def check_dmc(data, out, configs, weights):
    assert np.abs(configs.configs - data["dmc_configs"]).max() < 1e-9
    if "dmc_wrap" in data:
        assert np.array_equal(configs.wrap, data["dmc_wrap"])
    assert relerr(weights, data["dmc_weights"]) < 1e-9
    for k in ("energytotal", "energyke", "energyecp", "energygrad2", "weight", "acceptance", "tmove_acceptance"):
        assert abs(out[k] - data["dmc_" + k]) <= 1e-9 * max(1.0, abs(data["dmc_" + k])), k
