"""ctypes binding of libqmcb200.so (C ABI in include/qmcb200.h).

There is no CPU fallback: importing this module fails loudly when the shared library has
not been built (``python -c 'import __graft_entry__ as g; g.build()'``), and creating a
context fails loudly when no CUDA device is present.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqmcb200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int32)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_double = ctypes.c_double
c_i64 = ctypes.c_int64

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header
SIGNATURES = {
    "qmcb_last_error": (ctypes.c_char_p, []),
    "qmcb_device_count": (c_int, []),
    "qmcb_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "qmcb_destroy": (None, [c_void_p]),
    "qmcb_set_atoms": (c_int, [c_void_p, c_int, c_double_p, c_double_p]),
    "qmcb_set_basis": (c_int, [c_void_p, c_int, c_int_p, c_int_p, c_int_p, c_double_p, c_double_p]),
    "qmcb_set_slater": (c_int, [c_void_p, c_int, c_int, c_int, c_double_p, c_int, c_double_p, c_int,
                                c_int_p, c_int, c_int_p, c_int, c_int_p, c_int_p, c_double_p]),
    "qmcb_set_slater_cx": (c_int, [c_void_p, c_int, c_int, c_int, c_double_p, c_double_p, c_int, c_double_p, c_double_p,
                                   c_int, c_int_p, c_int, c_int_p, c_int, c_int_p, c_int_p, c_double_p, c_double_p]),
    "qmcb_is_complex": (c_int, [c_void_p]),
    "qmcb_set_pbc_phases_imag": (c_int, [c_void_p, c_int, c_int, c_double_p]),
    "qmcb_set_jastrow": (c_int, [c_void_p, c_int, c_int, c_int, c_int_p, c_double_p, c_double, c_int,
                                 c_int_p, c_double_p, c_double, c_double_p, c_double_p]),
    "qmcb_set_jastrow3": (c_int, [c_void_p, c_int, c_int, c_int, c_int_p, c_double_p, c_double, c_int,
                                  c_int_p, c_double_p, c_double, c_double_p]),
    "qmcb_set_ecp": (c_int, [c_void_p, c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_double_p,
                             c_double_p, c_int_p, c_double_p, c_double]),
    "qmcb_set_lattice": (c_int, [c_void_p, c_double_p, c_int, c_double_p]),
    "qmcb_set_pbc_orbitals": (c_int, [c_void_p, c_int, c_double_p, c_double_p, c_double_p, c_int, c_double_p, c_int,
                                      c_double_p, c_int_p, c_double_p, c_int, c_double_p, c_double_p, c_int, c_int_p,
                                      c_int, c_int_p, c_int]),
    "qmcb_set_ewald": (c_int, [c_void_p, c_double, c_int, c_double_p, c_int, c_double_p, c_double_p, c_double_p,
                               c_double_p, c_double, c_double, c_double, c_double]),
    "qmcb_set_point_wrap": (c_int, [c_void_p, c_double_p, c_i64]),
    "qmcb_recompute_pbc": (c_int, [c_void_p, c_int, c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "qmcb_recompute": (c_int, [c_void_p, c_int, c_int, c_double_p, c_double_p, c_double_p]),
    "qmcb_recompute_resident": (c_int, [c_void_p, c_int]),
    "qmcb_recompute_resident_on": (c_int, [c_void_p, c_int, c_void_p]),
    "qmcb_value": (c_int, [c_void_p, c_int, c_double_p, c_double_p]),
    "qmcb_gradient": (c_int, [c_void_p, c_int, c_int, c_double_p, c_double_p]),
    "qmcb_gradient_value": (c_int, [c_void_p, c_int, c_int, c_double_p, c_double_p, c_double_p, c_i64_p]),
    "qmcb_gradient_laplacian": (c_int, [c_void_p, c_int, c_int, c_double_p, c_double_p, c_double_p]),
    "qmcb_testvalue": (c_int, [c_void_p, c_int, c_int, c_double_p, c_int, c_u8_p, c_double_p, c_i64_p]),
    "qmcb_testvalue_many": (c_int, [c_void_p, c_int, c_int, c_int_p, c_double_p, c_u8_p, c_double_p]),
    "qmcb_updateinternals": (c_int, [c_void_p, c_int, c_int, c_double_p, c_u8_p, c_i64]),
    "qmcb_pgradient": (c_int, [c_void_p, ctypes.c_char_p, c_double_p]),
    "qmcb_get_state": (c_int, [c_void_p, ctypes.c_char_p, c_double_p]),
    "qmcb_energy": (c_int, [c_void_p, c_double_p, c_double_p, c_double_p]),
    "qmcb_tmoves": (c_int, [c_void_p, c_int, c_double, c_double_p, c_double_p, c_double_p, c_double_p,
                            c_double_p]),
    "qmcb_vmc_block": (c_int, [c_void_p, c_int, c_double, c_int, c_double_p, c_double_p, c_double_p,
                               c_double_p, c_double_p, c_u8_p, c_double_p, c_double_p, c_i64_p]),
    "qmcb_vmc_block_device": (c_int, [c_void_p, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "qmcb_vmc_upload": (c_int, [c_void_p, c_int, c_int, c_int, c_i64, c_int, c_double_p, c_double_p, c_double_p,
                                c_double_p]),
    "qmcb_vmc_block_slot": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_double_p, c_u8_p, c_double_p,
                                    c_double_p, c_i64_p]),
    "qmcb_vmc_block_slot_begin": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_double_p, c_double_p, c_i64_p]),
    "qmcb_vmc_block_slot_end": (c_int, [c_void_p, c_int]),
    "qmcb_comb_indices": (c_int, [c_i64, c_double_p, c_double, c_i64_p, c_double_p]),
    "qmcb_fp64_peak": (c_int, [c_int, c_double_p]),
    "qmcb_gemm_tn": (c_int, [c_int, c_i64, c_int, c_double_p, c_double_p, c_double_p, c_int, c_int, c_double_p]),
    "qmcb_orbitals_at_points": (c_int, [c_void_p, c_i64, c_double_p, c_int, c_double_p, c_double_p]),
    "qmcb_kernel_launches": (c_int, [c_void_p, c_i64_p]),
    "qmcb_sr_avg": (c_int, [c_void_p, c_int, c_int_p, c_i64_p, c_double_p, c_double_p, c_double_p, c_double, c_double_p,
                            c_double_p, c_double_p, c_double_p]),
    "qmcb_dmc_block": (c_int, [c_void_p, c_int, c_double, c_double, c_double, c_double, c_double_p, c_double_p,
                               c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                               c_double_p, c_double_p, c_i64_p, c_i64_p]),
    "qmcb_pinned_alloc": (c_int, [c_i64, ctypes.POINTER(c_void_p)]),
    "qmcb_pinned_free": (c_int, [c_void_p]),
    "qmcb_sm_update_device": (c_int, [c_int, c_int, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "qmcb_rng_vmc_block": (c_int, [ctypes.POINTER(ctypes.c_uint32), c_int_p, c_int_p, c_double_p, c_int, c_int, c_i64,
                                   c_int, c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_int]),
    "qmcb_rng_program": (c_int, [ctypes.POINTER(ctypes.c_uint32), c_int_p, c_int_p, c_double_p, c_i64, c_int_p, c_i64_p,
                                 ctypes.POINTER(ctypes.c_uint64), c_double_p, c_int]),
    "qmcb_rng_plan_create": (c_void_p, []),
    "qmcb_rng_plan_destroy": (None, [c_void_p]),
    "qmcb_rng_phase_a": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_uint32), c_int_p, c_int_p, c_double_p, c_int, c_int,
                                 c_i64, c_int, c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_int]),
    "qmcb_rng_phase_b": (c_int, [c_void_p, c_int]),
    "qmcb_sm_update": (c_int, [c_int, c_int, c_i64, c_double_p, c_double_p, c_u8_p, c_double_p]),
    "qmcb_devrng_set_state": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_uint32), c_int, c_int, c_double]),
    "qmcb_devrng_get_state": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_uint32), c_int_p, c_int_p, c_double_p]),
    "qmcb_devrng_program": (c_int, [c_void_p, c_i64, c_int_p, c_i64_p, ctypes.POINTER(ctypes.c_uint64), c_double_p]),
    "qmcb_devrng_vmc_block": (c_int, [c_void_p, c_int, c_int, c_int, c_i64, c_int, c_double]),
    "qmcb_devrng_dmc_block": (c_int, [c_void_p, c_int, c_int, c_int, c_i64, c_int, c_double, c_int, c_int]),
    "qmcb_dmc_block_slot": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_double_p, c_double_p,
                                    c_double_p, c_i64_p, c_i64_p, c_double_p]),
    "qmcb_glibc_log_mismatches": (c_i64, [c_i64, ctypes.c_uint64]),
    "qmcb_devrng_generator_timing": (c_int, [c_void_p, c_i64_p]),
}

_lib = None


class QmcbError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'). pyqmc_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise QmcbError(load().qmcb_last_error().decode())


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def u8ptr(a):
    return None if a is None else a.ctypes.data_as(c_u8_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _PinnedBlock:
    """Owner of one cudaMallocHost allocation: freed when the last numpy view of it is gone."""

    def __init__(self, nbytes):
        p = c_void_p()
        check(load().qmcb_pinned_alloc(max(nbytes, 1), ctypes.byref(p)))
        self.ptr = p

    def __del__(self):
        try:
            if self.ptr:
                load().qmcb_pinned_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class PinnedArray:
    """numpy view of a page-locked host buffer (cudaMallocHost).  The allocation belongs to the VIEW chain, not to
    this object: ``array.base`` is a ctypes buffer that references the owning block, so any array (or slice of it)
    handed out -- to a prefetch thread, a caller-held variates dictionary -- keeps the memory alive after the
    buffer set that created it was replaced."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        count = int(np.prod(self.shape, dtype=np.int64))
        nbytes = count * self.dtype.itemsize
        block = _PinnedBlock(nbytes)
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(block.ptr.value)
        buf._pinned_block = block  # ndarray.base -> buf -> block
        self.array = np.frombuffer(buf, dtype=self.dtype, count=count).reshape(self.shape)
