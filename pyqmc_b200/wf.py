"""B200 wave-function objects with the reference's ``pyqmc.wf`` object protocol.

``Slater``, ``JastrowSpin`` and ``MultiplyWF`` keep the method names, argument meaning and
return shapes of ``pyqmc/wf/slater.py:97-542``, ``pyqmc/wf/jastrowspin.py:20-464`` and
``pyqmc/wf/multiplywf.py:71-132`` (protocol summary: ``doc/source/wavefunction.rst:1-37``),
so ``pyqmc.method.mc.vmc``, ``pyqmc.method.dmc.rundmc`` and the accumulators drive them
unchanged.  All arithmetic runs in ``libqmcb200.so`` (hand-written sm_100a kernels) through
ctypes; walker state lives on the device, every call takes and returns host numpy arrays.

A ``MultiplyWF`` of one ``Slater`` and one ``JastrowSpin`` shares a single device context:
each protocol call is then ONE fused kernel launch and one device->host copy.
"""
import ctypes
import os

import numpy as np

from . import _lib, basis as _basis
from .func3d import CutoffCuspFunction, PolyPadeFunction  # noqa: F401  (re-export)

SLATER = 1
JASTROW = 2
JASTROW3 = 4


def default_device():
    if "PYQMC_B200_DEVICE" in os.environ:
        return int(os.environ["PYQMC_B200_DEVICE"])
    if "LOCAL_RANK" in os.environ:
        n = _lib.load().qmcb_device_count()
        return int(os.environ["LOCAL_RANK"]) % max(n, 1)
    return 0


class SavedSlot:
    """Opaque ``saved_values`` token: the MO row / position stay on the device.

    The saved row is a function of (electron, position) only, so two tokens stand for equal saved
    values exactly when those agree; ``shape`` and ``-`` let the reference's harness compare the tokens
    of ``testvalue`` and ``gradient_value`` (``testwf.py:248-256``) as it compares saved arrays."""

    __slots__ = ("token", "key")
    shape = ()

    def __init__(self, token, key=None):
        self.token = int(token)
        self.key = key

    def __sub__(self, other):
        same = isinstance(other, SavedSlot) and self.key is not None and self.key == other.key
        return 0.0 if same else float("inf")


def _point_key(e, positions):
    import zlib

    return int(e), positions.shape, zlib.crc32(positions)


def _token(saved):
    if isinstance(saved, SavedSlot):
        return saved.token
    if isinstance(saved, (tuple, list)):  # MultiplyWF-style tuple of per-factor tokens
        for s in saved:
            if isinstance(s, SavedSlot):
                return s.token
    return -1


class DeviceContext:
    """Owns one ``qmcb_ctx`` (device tables + walker state)."""

    def __init__(self, mol, device=None):
        self.lib = _lib.load()
        self.device = default_device() if device is None else int(device)
        h = ctypes.c_void_p()
        _lib.check(self.lib.qmcb_create(self.device, ctypes.byref(h)))
        self.h = h
        self.mol = mol
        xyz = _lib.f64(mol.atom_coords())
        chg = _lib.f64(mol.atom_charges())
        _lib.check(self.lib.qmcb_set_atoms(h, len(chg), _lib.dptr(xyz), _lib.dptr(chg)))
        self.natom = len(chg)
        self.nconf = 0
        self.epoch = 0  # bumped by every call that changes the walker state (recompute, updateinternals)
        self._resident = None  # (BlockBuffers whose newconf a device block returned, epoch) -- mc.vmc_block_device
        self.nelec = tuple(int(x) for x in mol.nelec)
        self.ecp_key = None
        self.has_basis = False
        self.cplx = False  # set by a complex Slater factor (qmcb_set_slater_cx): Slater-valued outputs are complex128
        self.periodic = hasattr(mol, "a")
        if self.periodic:
            from . import pbc

            lat = _lib.f64(mol.lattice_vectors())
            mode, shifts = pbc.minimal_image_tables(lat)
            _lib.check(self.lib.qmcb_set_lattice(h, _lib.dptr(lat), mode, _lib.dptr(_lib.f64(shifts))))

    def __del__(self):
        try:
            if self.h:
                self.lib.qmcb_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_basis(self):
        if self.has_basis:
            return
        # periodic systems: the orbitals live on the primitive cell (orbitals.py:141, 201)
        t = _basis.shell_tables(self.mol.original_cell if self.periodic else self.mol)
        self.nao = t["nao"]
        _lib.check(self.lib.qmcb_set_basis(self.h, len(t["shell_l"]), _lib.iptr(t["shell_atom"]),
                                           _lib.iptr(t["shell_l"]), _lib.iptr(t["prim_off"]),
                                           _lib.dptr(t["exps"]), _lib.dptr(t["coefs"])))
        self.has_basis = True

    def _dt(self, which):
        """dtype of the wave-function-valued outputs of a call on the factors ``which`` (include/qmcb200.h,
        qmcb_set_slater_cx): complex when the context's Slater factor is complex and takes part."""
        return complex if (self.cplx and (which & SLATER)) else float

    # ---- protocol calls -----------------------------------------------------------------
    def recompute(self, which, configs, wrap=None):
        self.epoch += 1
        c = _lib.f64(configs)
        n = c.shape[0]
        sign, logv = np.empty(n, dtype=self._dt(which)), np.empty(n)
        if self.periodic:
            wr = None if wrap is None else _lib.f64(wrap)
            _lib.check(self.lib.qmcb_recompute_pbc(self.h, which, n, _lib.dptr(c), _lib.dptr(wr), _lib.dptr(sign),
                                                   _lib.dptr(logv)))
        else:
            _lib.check(self.lib.qmcb_recompute(self.h, which, n, _lib.dptr(c), _lib.dptr(sign), _lib.dptr(logv)))
        self.nconf = n
        return sign, logv

    def value(self, which):
        sign, logv = np.empty(self.nconf, dtype=self._dt(which)), np.empty(self.nconf)
        _lib.check(self.lib.qmcb_value(self.h, which, _lib.dptr(sign), _lib.dptr(logv)))
        return sign, logv

    def _epos(self, epos, naip=1):
        p = _lib.f64(epos.configs)
        want = (self.nconf, 3) if naip == 1 and p.ndim == 2 else (self.nconf, naip, 3)
        if p.shape != want:
            raise ValueError(f"electron positions have shape {p.shape}, expected {want}")
        if self.periodic:  # PeriodicElectron.wrap travels with the positions (coord.py:115-134)
            wrap = getattr(epos, "wrap", None)
            if wrap is not None:
                wr = _lib.f64(wrap)
                if wr.shape != p.shape:
                    raise ValueError(f"wrap vectors have shape {wr.shape}, expected {p.shape}")
                _lib.check(self.lib.qmcb_set_point_wrap(self.h, _lib.dptr(wr), wr.size // 3))
            else:
                _lib.check(self.lib.qmcb_set_point_wrap(self.h, None, 0))
        return p

    def gradient(self, which, e, epos):
        p = self._epos(epos)
        g = np.empty((3, self.nconf), dtype=self._dt(which))
        _lib.check(self.lib.qmcb_gradient(self.h, which, int(e), _lib.dptr(p), _lib.dptr(g)))
        return g

    def gradient_value(self, which, e, epos):
        p = self._epos(epos)
        g, v = np.empty((3, self.nconf), dtype=self._dt(which)), np.empty(self.nconf, dtype=self._dt(which))
        slot = ctypes.c_int64(-1)
        _lib.check(self.lib.qmcb_gradient_value(self.h, which, int(e), _lib.dptr(p), _lib.dptr(g),
                                                _lib.dptr(v), ctypes.byref(slot)))
        return g, v, SavedSlot(slot.value, _point_key(e, p))

    def gradient_laplacian(self, which, e, epos):
        p = self._epos(epos)
        g, lap = np.empty((3, self.nconf), dtype=self._dt(which)), np.empty(self.nconf, dtype=self._dt(which))
        _lib.check(self.lib.qmcb_gradient_laplacian(self.h, which, int(e), _lib.dptr(p), _lib.dptr(g),
                                                    _lib.dptr(lap)))
        return g, lap

    @staticmethod
    def _mask(mask, n):
        if mask is None:
            return None, n
        m = np.ascontiguousarray(np.asarray(mask, dtype=bool)).view(np.uint8)
        if m.shape != (n,):
            raise ValueError(f"mask has shape {m.shape}, expected ({n},)")
        return m, int(m.sum())

    def testvalue(self, which, e, epos, mask=None):
        aux = np.ndim(epos.configs) == 3
        naip = epos.configs.shape[1] if aux else 1
        p = self._epos(epos, naip)
        m, nm = self._mask(mask, self.nconf)
        out = np.empty((nm, naip), dtype=self._dt(which))
        slot = ctypes.c_int64(-1)
        _lib.check(self.lib.qmcb_testvalue(self.h, which, int(e), _lib.dptr(p), naip, _lib.u8ptr(m),
                                           _lib.dptr(out), ctypes.byref(slot)))
        return (out if aux else out[:, 0]), SavedSlot(slot.value, _point_key(e, p))

    def testvalue_many(self, which, e, epos, mask=None):
        el = _lib.i32(np.asarray(e))
        p = self._epos(epos)
        m, nm = self._mask(mask, self.nconf)
        out = np.empty((nm, len(el)), dtype=self._dt(which))
        _lib.check(self.lib.qmcb_testvalue_many(self.h, which, len(el), _lib.iptr(el), _lib.dptr(p),
                                                _lib.u8ptr(m), _lib.dptr(out)))
        return out

    def recompute_resident(self, which):
        """recompute from the coordinates the device already holds (block driver, see mc.vmc_block_device)"""
        _lib.check(self.lib.qmcb_recompute_resident(self.h, which))

    def updateinternals(self, which, e, epos, mask=None, saved_values=None):
        self.epoch += 1
        p = self._epos(epos)
        m, _ = self._mask(mask, self.nconf)
        _lib.check(self.lib.qmcb_updateinternals(self.h, which, int(e), _lib.dptr(p), _lib.u8ptr(m),
                                                 _token(saved_values)))

    def get_state(self, name, shape, dtype=float):
        out = np.empty(shape, dtype=dtype)
        _lib.check(self.lib.qmcb_get_state(self.h, name.encode(), _lib.dptr(out)))
        return out

    def kernel_launches(self):
        n = ctypes.c_int64(0)
        _lib.check(self.lib.qmcb_kernel_launches(self.h, ctypes.byref(n)))
        return n.value


# ---- determinant bookkeeping (determinant_tools.py:39-71, pyscftools.py:176-219) -------------
def _single_determinant(mf):
    try:
        mfu = mf.to_uhf()
    except TypeError:
        mfu = mf.to_uhf(mf)
    return [(1.0, [list(np.nonzero(np.asarray(o) > 0.5)[0]) for o in mfu.mo_occ])]


def _realify_orbitals(mo, what="orbital"):
    """Real orbital coefficients from complex ones when every orbital is real up to a constant phase, else ``None``.

    Mean-field codes hand out complex128 coefficients even where the orbitals can be chosen real (pyscf k-point
    objects at Gamma or at time-reversal-invariant k-points).  Each column is rotated by the
    phase of its largest coefficient; if what remains has a negligible imaginary part the real part is returned
    (the wave function changes by a constant global phase, which no observable of this path depends on) and the
    wave function stays on the real kernels.  Genuinely complex orbitals return ``None``: the caller keeps the
    complex coefficients (``dtype = complex``, the kernels of ``csrc/cplx.cuh``)."""
    mo = np.asarray(mo)
    if not np.iscomplexobj(mo):
        return np.array(mo, dtype=float)
    out = np.empty(mo.shape, dtype=float)
    for j in range(mo.shape[1]):
        col = mo[:, j]
        big = col[np.argmax(np.abs(col))] if len(col) else 1.0
        rot = col * (np.conj(big) / abs(big)) if abs(big) > 0 else col
        if np.abs(rot.imag).max(initial=0.0) > 1e-9 * max(np.abs(rot).max(initial=0.0), 1e-300):
            return None
        out[:, j] = rot.real
    return out


def _orbital_coefficients(blocks):
    """Real coefficient blocks when EVERY block can be realified, else all of them as complex128 (one dtype per
    wave function, slater.py:212-216)."""
    real = [_realify_orbitals(b) for b in blocks]
    if all(r is not None for r in real):
        return real
    return [np.array(b, dtype=complex) for b in blocks]


def _pack_determinants(determinants, tol):
    coeff, occ, dmap = [], [[], []], [[], []]
    for w, spin_occ in determinants:
        if abs(w) <= tol:
            continue
        coeff.append(w if np.iscomplexobj(w) and np.imag(w) != 0 else float(np.real(w)))
        for s in (0, 1):
            o = [int(i) for i in spin_occ[s]]
            if o not in occ[s]:
                occ[s].append(o)
            dmap[s].append(occ[s].index(o))
    return np.array(coeff), occ, np.array(dmap, dtype=np.int32)


class _Seed:
    """Walkers read back from a device context (what a copied wave function starts from)."""

    def __init__(self, configs, wrap):
        self.configs = configs
        if wrap is not None:
            self.wrap = wrap


def _resident_walkers(ctx):
    if ctx is None or ctx.nconf == 0:
        return None
    shape = (ctx.nconf, sum(ctx.nelec), 3)
    return _Seed(ctx.get_state("configs", shape), ctx.get_state("wrap", shape) if ctx.periodic else None)


class _DeviceFactor:
    """Shared plumbing of the two device-resident factors."""

    _which = 0
    dtype = float

    def _ensure_ctx(self):
        if self._ctx is None:
            self._ctx = DeviceContext(self._mol, self._device)
            self._push_static(self._ctx)
            self._dirty = True
        return self._ctx

    def _bind(self, ctx):
        """Move this factor into a shared context (MultiplyWF fusion)."""
        self._ctx = ctx
        self._push_static(ctx)
        self._dirty = True

    def _sync(self):
        ctx = self._ensure_ctx()
        self._push_parameters(ctx)
        return ctx

    # pickling / copying: device handles are rebuilt lazily (mc.py:160-163 pickles wf objects to
    # workers; testwf.py:44,77,108 uses copy.copy)
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_ctx"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._ctx = None

    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__getstate__())
        new.parameters = self._copy_parameters()
        new._seed = _resident_walkers(self._ctx)
        return new

    def __deepcopy__(self, memo):
        return self.__copy__()

    def _live(self):
        """The device context holding this object's walker state.  A copy owns its own context, created on
        first use from the walkers the original held when it was copied: the reference's harness queries
        ``copy.copy(wf)`` without a recompute of its own (testwf.py:44-54), as a shallow copy of numpy-backed
        objects allows."""
        if (self._ctx is None or self._ctx.nconf == 0) and getattr(self, "_seed", None) is not None:
            seed, self._seed = self._seed, None
            self.recompute(seed)
        if self._ctx is None:
            raise RuntimeError("wf.recompute(configs) must be called first")
        return self._ctx

    # ---- the wf protocol ------------------------------------------------------------------
    def recompute(self, configs):
        self._seed = None
        return self._sync().recompute(self._which, configs.configs, getattr(configs, "wrap", None))

    def value(self):
        return self._live().value(self._which)

    def gradient(self, e, epos):
        return self._live().gradient(self._which, e, epos)

    def gradient_value(self, e, epos):
        return self._live().gradient_value(self._which, e, epos)

    def gradient_laplacian(self, e, epos):
        return self._live().gradient_laplacian(self._which, e, epos)

    def testvalue(self, e, epos, mask=None):
        return self._live().testvalue(self._which, e, epos, mask)

    def testvalue_many(self, e, epos, mask=None):
        return self._live().testvalue_many(self._which, e, epos, mask)

    def updateinternals(self, e, epos, configs, mask=None, saved_values=None):
        self._live().updateinternals(self._which, e, epos, mask, saved_values)


class Slater(_DeviceFactor):
    """Multi-determinant Slater wave function (reference: ``pyqmc/wf/slater.py:97-542``).

    ``determinants`` uses the reference format ``[(weight, [occ_up, occ_dn]), ...]``
    (slater.py:167-179); without it a single determinant is taken from ``mf.mo_occ``.
    """

    _which = SLATER

    def __init__(self, mol, mf, mc=None, tol=None, twist=0, determinants=None,
                 eval_gto_precision=None, evaluate_orbitals_with="b200", device=None):
        if mc is not None and determinants is None:
            raise NotImplementedError("pass determinants=[(weight, [occ_up, occ_dn]), ...]; "
                                      "reading pyscf CI objects needs pyscf")
        self.tol = -1 if tol is None else tol
        self._mol = mol
        self._nelec = tuple(int(x) for x in mol.nelec)
        self._device = device
        self._ctx = None
        self._pbc = None
        if hasattr(mol, "a"):
            self._init_periodic(mol, mf, determinants, twist, eval_gto_precision)
            return
        if determinants is None:
            determinants = _single_determinant(mf)
        try:
            mfu = mf.to_uhf()
        except TypeError:
            mfu = mf.to_uhf(mf)
        top = [0, 0]
        for _, d in determinants:
            for s in (0, 1):
                if len(d[s]) > 0:
                    top[s] = max(top[s], int(np.max(d[s])) + 1)
        coeff, self._det_occup, self._det_map = _pack_determinants(determinants, self.tol)
        for s in (0, 1):
            for o in self._det_occup[s]:
                if len(o) != self._nelec[s]:
                    raise AssertionError(
                        f"disagreement between number of electrons and number of orbitals: "
                        f"{self._nelec[s]} electrons and {len(o)} orbitals")
        mo = _orbital_coefficients([np.asarray(mfu.mo_coeff[0])[:, : top[0]], np.asarray(mfu.mo_coeff[1])[:, : top[1]]])
        self.parameters = {"det_coeff": coeff, "mo_coeff_alpha": mo[0], "mo_coeff_beta": mo[1]}

    def _init_periodic(self, mol, mf, determinants, twist, eval_gto_precision):
        """Bloch orbitals on a supercell: k-points of the twist, per-k MO blocks concatenated,
        determinant lists flattened over k (pyscftools.py:140-186, determinant_tools.py:92-106) and the
        image / cutoff / phase tables of the reference's in-tree evaluator (pbcgto.py:594-621)."""
        from . import pbc

        if not hasattr(mol, "original_cell"):
            mol = pbc.get_supercell(mol, np.eye(3, dtype=int))
            self._mol = mol
        try:
            mfu = mf.to_uhf()
        except TypeError:
            mfu = mf.to_uhf(mf)
        kinds = pbc.create_supercell_twists(mol, mfu)["primitive_ks"][twist]
        if len(kinds) != mol.scale:
            raise ValueError(f"Found {len(kinds)} k-points but should have found {mol.scale}.")
        if determinants is None:
            determinants = [(1.0, [[list(np.nonzero(np.asarray(k) > 0.5)[0]) for k in s] for s in mfu.mo_occ])]

        def f_max_orb(a):
            return int(np.max(a, initial=0)) + 1 if len(a) > 0 else 0

        max_orb = np.amax([[[f_max_orb(k) for k in s] for s in det] for _, det in determinants], axis=0)
        blocks = [[np.asarray(mfu.mo_coeff[s][k])[:, 0:max_orb[s][k]] for k in kinds] for s in (0, 1)]
        offs = np.cumsum(max_orb[:, kinds], axis=1)
        offs = np.pad(offs[:, :-1], ((0, 0), (1, 0)))
        flat = []
        for wt, det in determinants:
            fd = [list(np.concatenate([np.asarray(det_s[k], dtype=int) + off_s[ki] for ki, k in enumerate(kinds)]).astype(int))
                  for det_s, off_s in zip(det, offs)]
            flat.append((wt, fd))
        coeff, self._det_occup, self._det_map = _pack_determinants(flat, self.tol)
        for s in (0, 1):
            for o in self._det_occup[s]:
                if len(o) != self._nelec[s]:
                    raise AssertionError(
                        f"disagreement between number of electrons and number of orbitals: "
                        f"{self._nelec[s]} electrons and {len(o)} orbitals")
        kpts = np.asarray(mfu.kpts)[kinds].reshape(-1, 3)
        tables = pbc.image_tables(mol.original_cell, kpts, eval_gto_precision)
        flatb = blocks[0] + blocks[1]
        if np.iscomplexobj(tables["phases"]):  # general twist: complex Bloch phases -> a complex wave function
            flatb = [np.array(b, dtype=complex) for b in flatb]
        else:
            flatb = _orbital_coefficients(flatb)
        blocks = [flatb[:len(kinds)], flatb[len(kinds):]]
        mo = [np.concatenate(blocks[s], axis=1) for s in (0, 1)]
        self._pbc = dict(
            kpts=kpts, tables=tables, isgamma=bool(np.abs(kpts).sum() < 1e-9),
            mo_k=[np.concatenate([np.full(b.shape[1], ki, dtype=np.int32) for ki, b in enumerate(blocks[s])]
                                 or [np.zeros(0, dtype=np.int32)]) for s in (0, 1)])
        self.parameters = {"det_coeff": coeff, "mo_coeff_alpha": np.array(mo[0]), "mo_coeff_beta": np.array(mo[1])}

    def _copy_parameters(self):
        return {k: np.array(v) for k, v in self.parameters.items()}

    def _push_static(self, ctx):
        ctx.set_basis()
        if self._pbc is not None:
            cell = self._mol.original_cell
            t = self._pbc["tables"]
            bxyz, lprim = _lib.f64(cell.atom_coords()), _lib.f64(cell.lattice_vectors())
            smat, kpts = _lib.f64(self._mol.S), _lib.f64(self._pbc["kpts"])
            Ls, ncut, acut = _lib.f64(t["Ls"]), _lib.i32(t["num_Ls"]), _lib.f64(t["atom_cutoff"])
            lcut, ph = _lib.f64(t["l_cutoff"]), _lib.f64(np.real(t["phases"]))
            mk = [_lib.i32(m) for m in self._pbc["mo_k"]]
            _lib.check(ctx.lib.qmcb_set_pbc_orbitals(
                ctx.h, len(bxyz), _lib.dptr(bxyz), _lib.dptr(lprim), _lib.dptr(smat), len(kpts), _lib.dptr(kpts),
                len(Ls), _lib.dptr(Ls), _lib.iptr(ncut), _lib.dptr(acut), len(lcut), _lib.dptr(lcut), _lib.dptr(ph),
                len(mk[0]), _lib.iptr(mk[0]), len(mk[1]), _lib.iptr(mk[1]), 1 if self._pbc["isgamma"] else 0))
            if np.iscomplexobj(t["phases"]):
                phi = _lib.f64(np.imag(t["phases"]))
                _lib.check(ctx.lib.qmcb_set_pbc_phases_imag(ctx.h, phi.shape[0], phi.shape[1], _lib.dptr(phi)))

    @property
    def dtype(self):
        """complex when any parameter or the Bloch phase table is complex (slater.py:212-216)."""
        cx = any(np.iscomplexobj(v) for v in self.parameters.values())
        if self._pbc is not None and np.iscomplexobj(self._pbc["tables"]["phases"]):
            cx = True
        return complex if cx else float

    def _push_parameters(self, ctx):
        p = self.parameters
        occ = [_lib.i32(np.asarray(self._det_occup[s]).reshape(len(self._det_occup[s]), -1)) for s in (0, 1)]
        m0, m1 = _lib.i32(self._det_map[0]), _lib.i32(self._det_map[1])
        if self.dtype == complex:
            arr = [np.asarray(p[k], dtype=complex) for k in ("mo_coeff_alpha", "mo_coeff_beta", "det_coeff")]
            re, im = [_lib.f64(a.real) for a in arr], [_lib.f64(a.imag) for a in arr]
            _lib.check(ctx.lib.qmcb_set_slater_cx(
                ctx.h, self._nelec[0], self._nelec[1], re[0].shape[1], _lib.dptr(re[0]), _lib.dptr(im[0]),
                re[1].shape[1], _lib.dptr(re[1]), _lib.dptr(im[1]),
                len(self._det_occup[0]), _lib.iptr(occ[0]), len(self._det_occup[1]), _lib.iptr(occ[1]),
                len(re[2]), _lib.iptr(m0), _lib.iptr(m1), _lib.dptr(re[2]), _lib.dptr(im[2])))
            ctx.cplx = True
            return
        cu, cd = _lib.f64(p["mo_coeff_alpha"]), _lib.f64(p["mo_coeff_beta"])
        dc = _lib.f64(p["det_coeff"])
        _lib.check(ctx.lib.qmcb_set_slater(
            ctx.h, self._nelec[0], self._nelec[1], cu.shape[1], _lib.dptr(cu), cd.shape[1], _lib.dptr(cd),
            len(self._det_occup[0]), _lib.iptr(occ[0]), len(self._det_occup[1]), _lib.iptr(occ[1]),
            len(dc), _lib.iptr(m0), _lib.iptr(m1), _lib.dptr(dc)))
        ctx.cplx = False

    def pgradient(self):
        """d ln Psi / d parameters: det_coeff (N, D), mo_coeff_alpha/beta (N, A, nmo_s)
        (slater.py:462-542), evaluated on the device from the stored inverses."""
        ctx = self._ctx
        N = ctx.nconf
        p = self.parameters
        shapes = {"det_coeff": (N, len(p["det_coeff"])),
                  "mo_coeff_alpha": (N,) + p["mo_coeff_alpha"].shape,
                  "mo_coeff_beta": (N,) + p["mo_coeff_beta"].shape}
        out = {}
        for k, shape in shapes.items():
            if int(np.prod(shape)) == 0:
                continue
            arr = np.empty(shape, dtype=self.dtype)
            _lib.check(ctx.lib.qmcb_pgradient(ctx.h, k.encode(), _lib.dptr(arr)))
            out[k] = arr
        return out

    # read-back of the reference's internal arrays (tests)
    @property
    def _inverse(self):
        N = self._ctx.nconf
        out = []
        for s, name in ((0, "inverse_up"), (1, "inverse_dn")):
            n = self._nelec[s]
            out.append(self._ctx.get_state(name, (N, len(self._det_occup[s]), n, n), self.dtype))
        return out

    @property
    def _dets(self):
        N = self._ctx.nconf
        return [self._ctx.get_state(name, (2, N, len(self._det_occup[s])), self.dtype)
                for s, name in ((0, "dets_up"), (1, "dets_dn"))]


class JastrowSpin(_DeviceFactor):
    """One- and two-body Jastrow factor (reference: ``pyqmc/wf/jastrowspin.py:20-464``)."""

    _which = JASTROW

    def __init__(self, mol, a_basis, b_basis, device=None):
        self._mol = mol
        self._nelec = tuple(int(x) for x in mol.nelec)
        self._device = device
        self._ctx = None
        self.a_basis = list(a_basis)
        self.b_basis = list(b_basis)
        for bas in (self.a_basis, self.b_basis):
            for f in bas:
                assert f.parameters["rcut"] == bas[0].parameters["rcut"]  # func3d.py:289-291
        self.parameters = {
            "bcoeff": np.zeros((len(self.b_basis), 3)),
            "acoeff": np.zeros((mol.natm, len(self.a_basis), 2)),
        }

    def _copy_parameters(self):
        return {k: np.array(v) for k, v in self.parameters.items()}

    def _push_static(self, ctx):
        pass

    def _push_parameters(self, ctx):
        ak = _lib.i32([f.kind for f in self.a_basis])
        ap = _lib.f64([f.shape_parameter for f in self.a_basis])
        bk = _lib.i32([f.kind for f in self.b_basis])
        bp = _lib.f64([f.shape_parameter for f in self.b_basis])
        ra = float(self.a_basis[0].parameters["rcut"]) if self.a_basis else 1.0
        rb = float(self.b_basis[0].parameters["rcut"]) if self.b_basis else 1.0
        ac, bc = _lib.f64(self.parameters["acoeff"]), _lib.f64(self.parameters["bcoeff"])
        _lib.check(ctx.lib.qmcb_set_jastrow(ctx.h, self._nelec[0], self._nelec[1], len(ak), _lib.iptr(ak),
                                            _lib.dptr(ap), ra, len(bk), _lib.iptr(bk), _lib.dptr(bp), rb,
                                            _lib.dptr(ac), _lib.dptr(bc)))

    def pgradient(self):
        N = self._ctx.nconf
        return {
            "bcoeff": self._ctx.get_state("bvalues", (N, len(self.b_basis), 3)),
            "acoeff": self._ctx.get_state("avalues", (N, self._mol.natm, len(self.a_basis), 2)),
        }

    @property
    def _a_partial(self):
        ne = sum(self._nelec)
        return self._ctx.get_state("a_partial", (ne, self._ctx.nconf, self._mol.natm, len(self.a_basis)))

    @property
    def _b_partial(self):
        ne = sum(self._nelec)
        return self._ctx.get_state("b_partial", (ne, self._ctx.nconf, len(self.b_basis), 2))


class ThreeBodyJastrow(_DeviceFactor):
    """Electron-electron-ion Jastrow factor (reference: ``pyqmc/wf/three_body_jastrow.py``).

    Cache updates use the factor's own copy of the walker coordinates (as ``JastrowSpin`` does),
    i.e. the behaviour of the reference when ``updateinternals`` is called before ``configs.move``
    (its test harness, ``testwf.py:116-119``).  The reference's drivers move ``configs`` first
    (``mc.py:135-136``), after which its ``P_i`` cache no longer matches a fresh ``recompute``;
    that inconsistency is deliberately not reproduced."""

    _which = JASTROW3

    def __init__(self, mol, a_basis, b_basis, device=None):
        # periodic systems: every displacement goes through the minimal image (distance.py:83-159), in the per-call
        # protocol kernels and in the device-resident blocks (k_pbc_move_general + k_jastrow3_update_coop)
        self._mol = mol
        self._nelec = tuple(int(x) for x in mol.nelec)
        self._device = device
        self._ctx = None
        self.a_basis = list(a_basis)
        self.b_basis = list(b_basis)
        self.parameters = {"ccoeff": np.zeros((mol.natm, len(a_basis), len(a_basis), len(b_basis), 3))}

    def _copy_parameters(self):
        return {k: np.array(v) for k, v in self.parameters.items()}

    def _push_static(self, ctx):
        pass

    def _push_parameters(self, ctx):
        ak = _lib.i32([f.kind for f in self.a_basis])
        ap = _lib.f64([f.shape_parameter for f in self.a_basis])
        bk = _lib.i32([f.kind for f in self.b_basis])
        bp = _lib.f64([f.shape_parameter for f in self.b_basis])
        ra = float(self.a_basis[0].parameters["rcut"]) if self.a_basis else 1.0
        rb = float(self.b_basis[0].parameters["rcut"]) if self.b_basis else 1.0
        cc = _lib.f64(self.parameters["ccoeff"])
        _lib.check(ctx.lib.qmcb_set_jastrow3(ctx.h, self._nelec[0], self._nelec[1], len(ak), _lib.iptr(ak),
                                             _lib.dptr(ap), ra, len(bk), _lib.iptr(bk), _lib.dptr(bp), rb,
                                             _lib.dptr(cc)))

    def pgradient(self):
        N = self._ctx.nconf
        out = np.empty((N,) + self.parameters["ccoeff"].shape)
        _lib.check(self._ctx.lib.qmcb_pgradient(self._ctx.h, b"ccoeff", _lib.dptr(out)))
        return {"ccoeff": out}

    @property
    def P_i(self):
        return self._ctx.get_state("P_i", (sum(self._nelec), self._ctx.nconf))

    @property
    def a_values(self):
        return self._ctx.get_state("a3_values", (sum(self._nelec), self._ctx.nconf, self._mol.natm, len(self.a_basis)))


class Parameters:
    """``wfN``-prefixed view of the factors' parameter dictionaries (multiplywf.py:18-68)."""

    def __init__(self, dicts):
        self.data = {f"wf{i + 1}": d for i, d in enumerate(dicts)}
        self.wf_count = len(dicts)

    def __getitem__(self, idx):
        return self.data[idx[:3]][idx[3:]]

    def __setitem__(self, idx, value):
        self.data[idx[:3]][idx[3:]] = value

    def __delitem__(self, idx):
        del self.data[idx[:3]][idx[3:]]

    def __contains__(self, idx):
        return idx[:3] in self.data and idx[3:] in self.data[idx[:3]]

    def keys(self):
        for i in range(self.wf_count):
            k1 = f"wf{i + 1}"
            for k2 in self.data[k1].keys():
                yield k1 + k2

    __iter__ = keys

    def items(self):
        for k in self.keys():
            yield k, self[k]

    def values(self):
        for k in self.keys():
            yield self[k]

    def __len__(self):
        return sum(len(d) for d in self.data.values())

    def __repr__(self):
        return "Parameters(" + repr(self.data) + ")"


class MultiplyWF:
    """Product of wave-function factors (reference: ``pyqmc/wf/multiplywf.py:71-132``).

    Any product of at most one ``Slater``, one ``JastrowSpin`` and one ``ThreeBodyJastrow`` of this
    package is fused into a single device context.  Any other combination falls back to combining the factors' results on the host
    exactly as the reference does (each factor still evaluates on the device).
    """

    def __init__(self, *wf_factors):
        self.wf_factors = list(wf_factors)
        self.parameters = Parameters([wf.parameters for wf in self.wf_factors])
        kinds = [type(wf) for wf in self.wf_factors]
        device_kinds = (Slater, JastrowSpin, ThreeBodyJastrow)
        self._fused = (len(kinds) >= 2 and all(k in device_kinds for k in kinds) and len(set(kinds)) == len(kinds)
                       and all(wf._mol is self.wf_factors[0]._mol for wf in self.wf_factors))
        self._ctx = None
        self._which = 0
        if self._fused:
            for wf in self.wf_factors:
                self._which |= wf._which

    @property
    def dtype(self):
        return complex if any(wf.dtype == complex for wf in self.wf_factors) else float

    def _ensure_ctx(self):
        if self._ctx is None:
            f0 = self.wf_factors[0]
            self._ctx = DeviceContext(f0._mol, f0._device)
            for f in self.wf_factors:
                f._bind(self._ctx)
        return self._ctx

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_ctx"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._ctx = None

    def __copy__(self):
        import copy as _copy

        new = MultiplyWF(*[_copy.copy(f) for f in self.wf_factors])
        if self._fused:
            for f in new.wf_factors:
                f._seed = None
            new._seed = _resident_walkers(self._ctx)
        return new

    def __deepcopy__(self, memo):
        return self.__copy__()

    def _live(self):
        """See ``_DeviceFactor._live``: a copy starts from the walkers its original held."""
        if (self._ctx is None or self._ctx.nconf == 0) and getattr(self, "_seed", None) is not None:
            seed, self._seed = self._seed, None
            self.recompute(seed)
        if self._ctx is None:
            raise RuntimeError("wf.recompute(configs) must be called first")
        return self._ctx

    def recompute(self, configs):
        self._seed = None
        if self._fused:
            ctx = self._ensure_ctx()
            for f in self.wf_factors:
                f._push_parameters(ctx)
            return ctx.recompute(self._which, configs.configs, getattr(configs, "wrap", None))
        signs, vals = np.ones(len(configs.configs)), np.zeros(len(configs.configs))
        for wf in self.wf_factors:
            s, v = wf.recompute(configs)
            signs, vals = signs * s, vals + v
        return signs, vals

    def value(self):
        if self._fused:
            return self._live().value(self._which)
        res = np.array([wf.value() for wf in self.wf_factors])
        return np.prod(res[:, 0, :], axis=0), np.sum(res[:, 1, :], axis=0)

    def updateinternals(self, e, epos, configs, mask=None, saved_values=None):
        if self._fused:
            return self._live().updateinternals(self._which, e, epos, mask, saved_values)
        if saved_values is None or isinstance(saved_values, SavedSlot):
            saved_values = [saved_values] * len(self.wf_factors)
        for wf, sv in zip(self.wf_factors, saved_values):
            wf.updateinternals(e, epos, configs, mask=mask, saved_values=sv)

    def gradient(self, e, epos):
        if self._fused:
            return self._live().gradient(self._which, e, epos)
        return np.sum([wf.gradient(e, epos) for wf in self.wf_factors], axis=0)

    def gradient_value(self, e, epos):
        if self._fused:
            return self._live().gradient_value(self._which, e, epos)
        grads, vals, saved = zip(*[wf.gradient_value(e, epos) for wf in self.wf_factors])
        return np.sum(grads, axis=0), np.prod(vals, axis=0), saved

    def gradient_laplacian(self, e, epos):
        if self._fused:
            return self._live().gradient_laplacian(self._which, e, epos)
        grads, laps = zip(*[wf.gradient_laplacian(e, epos) for wf in self.wf_factors])
        cross = np.zeros(laps[0].shape, dtype=self.dtype)
        for i in range(len(grads)):
            for j in range(i + 1, len(grads)):
                cross += np.sum(grads[i] * grads[j], axis=0)
        return np.sum(grads, axis=0), np.sum(laps, axis=0) + cross * 2

    def testvalue(self, e, epos, mask=None):
        if self._fused:
            return self._live().testvalue(self._which, e, epos, mask)
        vals, saved = zip(*[wf.testvalue(e, epos, mask=mask) for wf in self.wf_factors])
        return np.prod(vals, axis=0), saved

    def testvalue_many(self, e, epos, mask=None):
        if self._fused:
            return self._live().testvalue_many(self._which, e, epos, mask)
        return np.prod([wf.testvalue_many(e, epos, mask=mask) for wf in self.wf_factors], axis=0)

    def pgradient(self):
        return Parameters([wf.pgradient() for wf in self.wf_factors])
