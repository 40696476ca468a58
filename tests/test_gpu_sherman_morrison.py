"""Sherman-Morrison kernel vs numpy det/inv, restating the reference's own unit test
(tests/unit/test_sherman_morrison.py:20-82: well-conditioned random matrices, tolerance 1e-13),
for every kernel variant (thread-per-matrix n <= 8, warp-per-matrix n <= 32), plus masks."""
import numpy as np
import pytest

from pyqmc_b200 import _lib

pytestmark = pytest.mark.gpu


def construct_mat(rng, nmat, n):
    u, _, v = np.linalg.svd(rng.randn(n, n))
    svals = (rng.rand(nmat, n) + 1) * rng.choice([-1, 1], (nmat, n))
    return np.einsum("ij,hj,jk->hik", u, svals, v)


def construct_vec(rng, matrix, e):
    nmat, n, _ = matrix.shape
    coef = rng.randn(nmat, n - 1)
    not_e = np.arange(n) != e
    vec = np.einsum("ij,ijk->ik", coef, matrix[:, not_e, :])
    proj = (rng.random_sample(nmat) - 1) * 2
    proj += np.sign(proj) * 0.5
    return vec + matrix[:, e, :] * proj[:, None]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 16, 17, 32])
def test_sherman_morrison(lib, n):
    rng = np.random.RandomState(n)
    nmat = 77
    e = n // 2
    matrix = construct_mat(rng, nmat, n)
    inv = np.linalg.inv(matrix)
    vec = construct_vec(rng, matrix, e)
    new = matrix.copy()
    new[:, e, :] = vec
    mask = (rng.rand(nmat) > 0.3).astype(np.uint8)
    for use_mask in (False, True):
        work = np.ascontiguousarray(inv.copy())
        ratio = np.zeros(nmat)
        m = mask if use_mask else None
        _lib.check(lib.qmcb_sm_update(n, e, nmat, _lib.dptr(work), _lib.dptr(np.ascontiguousarray(vec)),
                                      _lib.u8ptr(m), _lib.dptr(ratio)))
        sel = mask.astype(bool) if use_mask else np.ones(nmat, dtype=bool)
        npratio = np.linalg.det(new) / np.linalg.det(matrix)
        npinv = np.linalg.inv(new)
        assert np.abs(ratio[sel] - npratio[sel]).max() < 1e-13 * max(1.0, np.abs(npratio).max())
        assert np.abs(work[sel] - npinv[sel]).max() < 1e-12
        if use_mask:
            assert np.array_equal(work[~sel], inv[~sel]), "masked-out matrices must not change"


def test_rejects_bad_arguments(lib):
    a = np.zeros((1, 40, 40))
    v = np.zeros((1, 40))
    r = np.zeros(1)
    assert lib.qmcb_sm_update(40, 0, 1, _lib.dptr(a), _lib.dptr(v), None, _lib.dptr(r)) != 0
    assert b"n <= 32" in lib.qmcb_last_error()


@pytest.mark.parametrize("use_mask", [False, True])
def test_bulk_copy_staged_kernel_for_large_batches_of_32x32(lib, use_mask, monkeypatch):
    """Batches of >= 1776 matrices with n = 32 run k_sm_tma32 (matrices staged in shared memory by cp.async.bulk, updated
    in place, written back by cp.async.bulk): same answers as numpy and -- bit for bit -- as the register kernel
    k_sm_warp<32> (QMCB_SM_NO_TMA=1), with masked-out matrices untouched."""
    n, nmat, e = 32, 4099, 13
    rng = np.random.RandomState(5)
    matrix = construct_mat(rng, nmat, n)
    inv = np.linalg.inv(matrix)
    vec = np.ascontiguousarray(construct_vec(rng, matrix, e))
    mask = (rng.rand(nmat) > 0.4).astype(np.uint8) if use_mask else None
    out = {}
    for variant in ("tma", "registers"):
        if variant == "registers":
            monkeypatch.setenv("QMCB_SM_NO_TMA", "1")
        work, ratio = np.ascontiguousarray(inv.copy()), np.zeros(nmat)
        _lib.check(lib.qmcb_sm_update(n, e, nmat, _lib.dptr(work), _lib.dptr(vec), _lib.u8ptr(mask), _lib.dptr(ratio)))
        out[variant] = (work, ratio)
    sel = mask.astype(bool) if use_mask else np.ones(nmat, dtype=bool)
    assert np.array_equal(out["tma"][0], out["registers"][0]) and np.array_equal(out["tma"][1][sel], out["registers"][1][sel])
    new = matrix.copy()
    new[:, e, :] = vec
    assert np.abs(out["tma"][0][sel] - np.linalg.inv(new)[sel]).max() < 1e-12
    assert np.abs(out["tma"][1][sel] - (np.linalg.det(new) / np.linalg.det(matrix))[sel]).max() < 1e-12
    if use_mask:
        assert np.array_equal(out["tma"][0][~sel], inv[~sel])
