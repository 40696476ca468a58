"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle needs hours
there): incremental state == fresh recompute after device-resident blocks, walker-sharding invariance
(bit-exact: a shard run with its slice of the variates reproduces its slice of the full run),
move-and-return round trips of the Sherman-Morrison / Jastrow updates, and the stand-alone
Sherman-Morrison kernel at the C4 shape (131072 matrices, n = 32)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _slice_variates(v, sl):
    g, u, eu, er = v
    return (np.ascontiguousarray(g[:, :, sl]), np.ascontiguousarray(u[:, :, sl]),
            None if eu is None else np.ascontiguousarray(eu[:, :, :, sl]), er)


def _block_properties(name, nwalk, nsteps, ewald_kw):
    import pyqmc_b200 as pq
    from pyqmc_b200 import mc

    mol, mf, wf, _ = helpers.make_pair(name, seed=1)
    acc = pq.EnergyAccumulator(mol, **ewald_kw)
    np.random.seed(17)
    configs = pq.initial_guess(mol, nwalk)
    ne = configs.configs.shape[1]
    start = configs.copy()
    variates = mc.draw_block_variates(nwalk, ne, 0.5, nsteps, acc)
    avg, configs, data = mc.vmc_block_device(wf, configs, 0.5, nsteps, {"energy": acc}, variates=variates,
                                             return_walker_data=True)
    assert 0.2 < avg["acceptance"] < 0.9 and np.all(np.isfinite(data["energy"]))
    # (1) the incrementally updated state equals a fresh recompute of the final walkers
    sign_inc, log_inc = wf.value()
    inv_inc = wf.wf_factors[0]._inverse
    sign_new, log_new = wf.recompute(configs)
    assert np.array_equal(sign_inc, sign_new)
    assert np.abs(log_inc - log_new).max() < 1e-9 * max(1.0, np.abs(log_new).max())
    for a, b in zip(inv_inc, wf.wf_factors[0]._inverse):
        assert helpers.relerr(a, b) < 1e-7
    # (2) sharding invariance, bit for bit (walkers are independent; coord.py:72-80)
    half = nwalk // 2
    for sl in (slice(0, half), slice(half, nwalk)):
        mol2, mf2, wf2, _ = helpers.make_pair(name, seed=1)
        part = start.copy()  # (slicing through PeriodicConfigs.mask would re-wrap: not idempotent to the last bit)
        part.configs = start.configs[sl].copy()
        if hasattr(start, "wrap"):
            part.wrap = start.wrap[sl].copy()
        avg2, part, data2 = mc.vmc_block_device(wf2, part, 0.5, nsteps, {"energy": pq.EnergyAccumulator(mol2, **ewald_kw)},
                                                variates=_slice_variates(variates, sl), return_walker_data=True)
        assert np.array_equal(data2["accept"], data["accept"][:, :, sl])
        assert np.array_equal(part.configs, configs.configs[sl])
        assert np.array_equal(data2["energy"], data["energy"][:, :, sl])
    return mol, wf, configs


def test_c2_h2o_4096_walkers_block_properties(lib):
    mol, wf, configs = _block_properties("h2o", 4096, 3, {})
    # (3) move-and-return round trip of the masked updates at full size
    rng = np.random.RandomState(3)
    before = wf.value()[1].copy()
    for e in (1, 6):
        old = configs.configs[:, e].copy()
        new = old + 0.2 * rng.randn(len(old), 3)
        mask = rng.rand(len(old)) > 0.3
        r1, saved = wf.testvalue(e, configs.make_irreducible(e, new))
        wf.updateinternals(e, configs.make_irreducible(e, new), configs, mask=mask, saved_values=saved)
        configs.move(e, configs.make_irreducible(e, new), mask)
        r2, saved = wf.testvalue(e, configs.make_irreducible(e, old))
        assert np.abs(r1[mask] * r2[mask] - 1.0).max() < 1e-9
        wf.updateinternals(e, configs.make_irreducible(e, old), configs, mask=mask, saved_values=saved)
        configs.move(e, configs.make_irreducible(e, old), mask)
    assert np.abs(wf.value()[1] - before).max() < 1e-9 * max(1.0, np.abs(before).max())


def test_c4_diamond_1024_walkers_block_properties(lib):
    mol, wf, configs = _block_properties("diamond222", 1024, 1, {"ewald_gmax": 10})
    frac = configs.configs @ np.linalg.inv(mol.lattice_vectors())
    assert frac.min() >= -1e-12 and frac.max() < 1 + 1e-12, "walkers must stay wrapped into the simulation cell"
    assert np.array_equal(configs.wrap, np.round(configs.wrap))


def test_c5_dmc_2048_walkers_properties(lib):
    import pyqmc_b200 as pq
    from pyqmc_b200 import dmc

    mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
    acc = {"energy": pq.EnergyAccumulator(mol)}
    np.random.seed(23)
    configs = pq.initial_guess(mol, 2048)
    df, configs = pq.vmc(wf, configs, nblocks=1, nsteps_per_block=5, accumulators=acc)
    e0 = float(df["energytotal"][-1])
    weights = np.ones(2048)
    out, configs, weights = dmc.dmc_propagate(wf, configs, weights, 0.02, 10.0, e0, e0, nsteps=5, accumulators=acc)
    assert np.all(np.isfinite(weights)) and np.all(weights > 0)
    assert 0.9 < out["acceptance"] <= 1.0 and 0.0 <= out["tmove_acceptance"] < 0.05
    sign_inc, log_inc = wf.value()
    sign_new, log_new = wf.recompute(configs)
    assert np.array_equal(sign_inc, sign_new)  # fixed node: no walker crossed
    assert np.abs(log_inc - log_new).max() < 1e-9 * max(1.0, np.abs(log_new).max())


def test_sherman_morrison_c4_shape_round_trip(lib):
    """131072 matrices of 32x32 (1 GiB of inverses): replace row e, then put the old row back."""
    import ctypes

    import torch

    n, nmat, e = 32, 1 << 17, 13
    g = torch.Generator(device="cuda").manual_seed(5)
    mat = torch.randn(nmat, n, n, dtype=torch.float64, device="cuda", generator=g) + 4.0 * torch.eye(n, dtype=torch.float64, device="cuda")
    inv = torch.linalg.inv(mat).contiguous()
    orig = inv.clone()
    newrow = (mat[:, e, :] + 0.5 * torch.randn(nmat, n, dtype=torch.float64, device="cuda", generator=g)).contiguous()
    oldrow = mat[:, e, :].contiguous()
    r1 = torch.empty(nmat, dtype=torch.float64, device="cuda")
    r2 = torch.empty(nmat, dtype=torch.float64, device="cuda")
    vp = ctypes.c_void_p
    torch.cuda.synchronize()
    assert lib.qmcb_sm_update_device(n, e, nmat, vp(inv.data_ptr()), vp(newrow.data_ptr()), None, vp(r1.data_ptr()), None) == 0
    torch.cuda.synchronize()
    mat2 = mat.clone()
    mat2[:, e, :] = newrow
    # spot check against torch on a slice, identity residual on everything
    assert (inv[:64] - torch.linalg.inv(mat2[:64])).abs().max().item() < 1e-10
    resid = torch.bmm(mat2[::257], inv[::257]) - torch.eye(n, dtype=torch.float64, device="cuda")
    assert resid.abs().max().item() < 1e-9
    assert lib.qmcb_sm_update_device(n, e, nmat, vp(inv.data_ptr()), vp(oldrow.data_ptr()), None, vp(r2.data_ptr()), None) == 0
    torch.cuda.synchronize()
    assert ((r1 * r2) - 1.0).abs().max().item() < 1e-9
    assert (inv - orig).abs().max().item() < 1e-8 * orig.abs().max().item()
