"""The device-resident generator (csrc/device_rng.cuh) against numpy's global legacy generator, bit for bit:
values AND the generator state afterwards (np.random.get_state()), for draw programs that exercise the cached
Gaussian (odd counts), block boundaries, carried-in state, rotations (scipy Rotation.random) and the exact
program of a VMC block; and the public driver with the device generator against the host generator."""
import ctypes

import numpy as np
import pytest
import scipy.spatial.transform

import helpers

pytestmark = pytest.mark.gpu
U32P = ctypes.POINTER(ctypes.c_uint32)


def _ctx(lib):
    h = ctypes.c_void_p()
    assert lib.qmcb_create(0, ctypes.byref(h)) == 0, lib.qmcb_last_error()
    return h


def _set_state_from_numpy(lib, h):
    st = np.random.get_state()
    key = np.ascontiguousarray(st[1], dtype=np.uint32)
    assert lib.qmcb_devrng_set_state(h, key.ctypes.data_as(U32P), int(st[2]), int(st[3]), float(st[4])) == 0


def _get_state(lib, h):
    key = np.empty(624, dtype=np.uint32)
    pos, has, cached = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
    assert lib.qmcb_devrng_get_state(h, key.ctypes.data_as(U32P), ctypes.byref(pos), ctypes.byref(has), ctypes.byref(cached)) == 0, \
        lib.qmcb_last_error()
    return key, pos.value, has.value, cached.value


def _run_program(lib, h, ops):
    """ops: list of (kind, count, scale).  Returns the device outputs and numpy's for the same stream."""
    import torch

    outs = [torch.full((9 if k == 2 else n,), float("nan"), dtype=torch.float64, device="cuda") for k, n, s in ops]
    kind = np.array([k for k, n, s in ops], dtype=np.int32)
    count = np.array([n for k, n, s in ops], dtype=np.int64)
    dst = np.array([o.data_ptr() for o in outs], dtype=np.uint64)
    scale = np.array([s for k, n, s in ops], dtype=np.float64)
    assert lib.qmcb_devrng_program(h, len(ops), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                   count.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                   dst.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                   scale.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0, lib.qmcb_last_error()
    state = _get_state(lib, h)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in outs], state


def _numpy_program(ops):
    out = []
    for k, n, s in ops:
        if k == 0:
            out.append(np.random.random(size=n))
        elif k == 1:
            out.append(np.random.normal(scale=s, size=n))
        else:
            out.append(scipy.spatial.transform.Rotation.random().as_matrix().reshape(-1))
    return out


def _check(lib, ops, seed, predraw=0, pregauss=0):
    h = _ctx(lib)
    try:
        np.random.seed(seed)
        if predraw:
            np.random.random(size=predraw)   # start somewhere inside a state block
        if pregauss:
            np.random.normal(size=pregauss)  # odd: leaves a cached Gaussian in the state
        _set_state_from_numpy(lib, h)
        dev, (key, pos, has, cached) = _run_program(lib, h, ops)
        ref = _numpy_program(ops)
        for i, (a, b) in enumerate(zip(dev, ref)):
            assert np.array_equal(a, b), f"op {i} {ops[i]}: {np.sum(a != b)} of {a.size} values differ"
        st = np.random.get_state()
        assert np.array_equal(key, st[1]) and pos == st[2], (pos, st[2])
        assert has == st[3] and cached == st[4]
    finally:
        lib.qmcb_destroy(h)


def test_uniform_and_normal_draws_match_numpy(lib):
    _check(lib, [(0, 7, 1.0), (1, 10, 1.0), (0, 1000, 1.0), (1, 4096 * 3, np.sqrt(0.5)), (0, 4096, 1.0)], seed=1)


def test_odd_counts_carry_the_cached_gaussian(lib):
    _check(lib, [(1, 5, 2.0), (0, 3, 1.0), (1, 1, 1.0), (1, 1, 3.0), (1, 7, 1.0), (1, 12289, 0.7), (1, 2, 1.0), (0, 5, 1.0)], seed=2)


def test_state_carried_in_mid_block_with_cached_value(lib):
    _check(lib, [(1, 9, 1.0), (0, 700, 1.0), (1, 3, 1.0)], seed=3, predraw=155, pregauss=3)
    _check(lib, [(0, 11, 1.0)], seed=4, predraw=312)  # 312 doubles = exactly one state block: position 624


def test_rotations_match_scipy(lib):
    _check(lib, [(2, 4, 1.0), (0, 5, 1.0), (2, 4, 1.0), (1, 3, 1.0), (2, 4, 1.0), (2, 4, 1.0)], seed=5)


def test_consecutive_programs_chain_through_the_device_state(lib):
    h = _ctx(lib)
    try:
        np.random.seed(6)
        _set_state_from_numpy(lib, h)
        for ops in ([(1, 101, 1.0), (0, 50, 1.0)], [(1, 101, 1.0), (0, 50, 1.0)], [(2, 4, 1.0), (1, 33, 0.3)]):
            dev, state = _run_program(lib, h, ops)
            for a, b in zip(dev, _numpy_program(ops)):
                assert np.array_equal(a, b)
        st = np.random.get_state()
        assert np.array_equal(state[0], st[1]) and state[1:] == (st[2], st[3], st[4])
    finally:
        lib.qmcb_destroy(h)


def test_vmc_block_program_matches_host_generator(lib):
    """The exact draw program of a C2-shaped block (4096 walkers, 8 electrons, 3 ECP atoms, 2 steps)."""
    nsteps, ne, N, necp, sigma = 2, 8, 4096, 3, float(np.sqrt(0.5))
    ops = []
    for _ in range(nsteps):
        for _ in range(ne):
            ops += [(1, 3 * N, sigma), (0, N, 1.0)]
        for _ in range(ne * necp):
            ops += [(0, N, 1.0), (2, 4, 1.0)]
    _check(lib, ops, seed=7, predraw=17)


def test_segmented_generation_with_jump_ahead(lib):
    """Programs long enough for the generator to split the stream into concurrent segments (k_mt_jump: start blocks
    from the jump-ahead polynomial x^J mod phi, tools/mt_jump_poly.py): a full C2 block (10 steps, ~8400 state blocks =
    nine segments), twice in a row through the device-resident state, from a position inside a state block with a
    cached Gaussian -- every value and the final np.random state equal numpy's."""
    nsteps, ne, N, necp, sigma = 10, 8, 4096, 3, float(np.sqrt(0.5))
    ops = []
    for _ in range(nsteps):
        for _ in range(ne):
            ops += [(1, 3 * N, sigma), (0, N, 1.0)]
        for _ in range(ne * necp):
            ops += [(0, N, 1.0), (2, 4, 1.0)]
    h = _ctx(lib)
    try:
        np.random.seed(11)
        np.random.random(size=77)
        np.random.normal(size=5)
        _set_state_from_numpy(lib, h)
        for _ in range(2):
            dev, state = _run_program(lib, h, ops)
            for i, (a, b) in enumerate(zip(dev, _numpy_program(ops))):
                assert np.array_equal(a, b), f"op {i}: {np.sum(a != b)} of {a.size} values differ"
            st = np.random.get_state()
            assert np.array_equal(state[0], st[1]) and state[1:] == (st[2], st[3], st[4])
    finally:
        lib.qmcb_destroy(h)


def test_one_huge_uniform_draw_across_many_segments(lib):
    """A single draw of 3 million doubles (9600 state blocks): every word of every segment is consumed."""
    _check(lib, [(0, 3_000_000, 1.0), (1, 9, 1.0)], seed=12, predraw=5)


def test_public_vmc_device_generator_equals_host_generator(lib, monkeypatch):
    """pyqmc_b200.vmc: same accept counts, energies, walkers and final np.random state with either generator."""
    import pyqmc_b200 as pq
    from pyqmc_b200 import mc

    assert mc.device_rng_usable()
    results = []
    for host in (False, True):
        if host:
            monkeypatch.setenv("QMCB_HOST_RNG", "1")
        mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
        np.random.seed(9)
        configs = pq.initial_guess(mol, 257)
        np.random.seed(10)
        df, configs = pq.vmc(wf, configs, nblocks=4, nsteps_per_block=3, accumulators={"energy": pq.EnergyAccumulator(mol)})
        results.append((df, configs.configs.copy(), np.random.get_state()))
    (d1, c1, s1), (d2, c2, s2) = results
    assert np.array_equal(c1, c2)
    for k in ("energytotal", "energyecp", "energyke", "acceptance"):
        assert np.array_equal(d1[k], d2[k]), k
    assert np.array_equal(s1[1], s2[1]) and s1[2:] == s2[2:]


def test_rundmc_device_generator_equals_host_generator(lib, monkeypatch):
    """pyqmc_b200.rundmc (T-moves, weights, branching draw between blocks): identical results and final np.random
    state whether the legacy stream is continued on the device or drawn by the host generator."""
    import pyqmc_b200 as pq

    results = []
    for host in (False, True):
        if host:
            monkeypatch.setenv("QMCB_HOST_RNG", "1")
        mol, mf, wf, _ = helpers.make_pair("h2o", seed=1)
        np.random.seed(19)
        configs = pq.initial_guess(mol, 63)
        np.random.seed(20)
        df, configs, weights = pq.rundmc(wf, configs, tstep=0.02, nblocks=3, nsteps_per_block=2, vmc_warmup=2,
                                         accumulators={"energy": pq.EnergyAccumulator(mol)})
        results.append((df, configs.configs.copy(), weights.copy(), np.random.get_state()))
    (d1, c1, w1, s1), (d2, c2, w2, s2) = results
    assert np.array_equal(c1, c2) and np.array_equal(w1, w2)
    for k in ("energytotal", "weight", "acceptance", "tmove_acceptance", "e_trial", "max branches"):
        assert np.array_equal(d1[k], d2[k]), k
    assert np.array_equal(s1[1], s2[1]) and s1[2:] == s2[2:]
