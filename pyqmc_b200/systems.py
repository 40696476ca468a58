"""Duck-typed molecule / mean-field objects and the synthetic benchmark systems.

The reference builds its wave functions from pyscf ``Mole``/``SCF`` objects
(``pyqmc/pyscftools.py:105-191``).  pyscf is not available on the GPU boxes, so the
B200 path (and the oracle, and the golden-vector generator that drives the real
reference) use these small stand-ins.  They expose exactly the attributes the
reference hot path reads:

* ``_atom, _basis, _ecp, nelec, natm, cart`` and ``atom_coords(), atom_charges(),
  atom_symbol(i), atom_pure_symbol(i), has_ecp()``
  (consumed at ``pyqmc/wf/numba/gto.py:438-457``, ``pyqmc/wf/jastrowspin.py:49-53``,
  ``pyqmc/observables/eval_ecp.py:27-32,153``, ``pyqmc/observables/energy.py:40``,
  ``pyqmc/method/mc.py:42-57``, ``pyqmc/wftools.py:118-126``);
* ``mf.mo_coeff (2,A,nmo)``, ``mf.mo_occ (2,nmo)``, ``mf.to_uhf()``
  (``pyqmc/pyscftools.py:140-143,206-219``).

A real pyscf ``Mole`` satisfies the same protocol, so the B200 objects accept either.

The systems are SYNTHETIC: correct shell structure and ECP
functional form, even-tempered exponents, seeded random orthonormal MOs.
"""
import numpy as np


class Mol:
    """Minimal pyscf-``Mole`` look-alike (open boundary conditions, Bohr units)."""

    cart = False

    def __init__(self, atoms, basis, ecp, nelec, charges):
        # atoms: list of (symbol, (x, y, z)) in Bohr
        self._atom = [[str(s), tuple(float(v) for v in xyz)] for s, xyz in atoms]
        self._basis = basis
        self._ecp = ecp
        self.nelec = (int(nelec[0]), int(nelec[1]))
        self.natm = len(self._atom)
        self._charges = np.asarray(charges, dtype=float)

    def atom_coords(self):
        return np.array([a[1] for a in self._atom], dtype=float)

    def atom_charges(self):
        return self._charges.copy()

    def atom_symbol(self, i):
        return self._atom[i][0]

    def atom_pure_symbol(self, i):
        return self._atom[i][0]

    def has_ecp(self):
        return len(self._ecp) > 0

    @property
    def nao(self):
        n = 0
        for sym, _ in self._atom:
            n += sum(2 * sh[0] + 1 for sh in self._basis[sym])
        return n


class MF:
    """Minimal mean-field look-alike already in UHF layout."""

    def __init__(self, mo_coeff, mo_occ):
        self.mo_coeff = np.asarray(mo_coeff)
        self.mo_occ = np.asarray(mo_occ)

    def to_uhf(self, *args):
        return self


def _even_tempered(a0, ratio, n):
    return [a0 * ratio**i for i in range(n)]


def _contracted(l, exps, seed):
    """One contracted shell ``[l, [exp, coef], ...]`` with smooth synthetic coefficients."""
    rng = np.random.RandomState(seed)
    n = len(exps)
    coefs = np.exp(-0.5 * ((np.arange(n) - 0.6 * n) / (0.35 * n)) ** 2)
    coefs *= 1.0 + 0.2 * rng.uniform(-1, 1, n)
    return [l] + [[float(e), float(c)] for e, c in zip(exps, coefs)]


def _single(l, e):
    return [l, [float(e), 1.0]]


def _ecp_entry(ncore, zeff, a_loc, nonlocal_channels):
    """ccECP-shaped semi-local ECP in pyscf ``_ecp`` layout.

    ``[ncore, [[l, [terms r^-2, terms r^-1, terms r^0, terms r^1, ...]], ...]]`` with
    each term ``[alpha, c]`` (layout read at ``pyqmc/observables/eval_ecp.py:160-179``).
    The local channel is ``Z/r e^{-a r^2} + a Z r e^{-b r^2} + g e^{-d r^2}`` which
    cancels the bare ``-Z/r`` at the origin, as in the ccECP construction.
    """
    a1, a2, a3, g3 = a_loc
    local = [-1, [[], [[a1, zeff]], [[a3, g3]], [[a2, a1 * zeff]], [], [], []]]
    chans = [local]
    for l, (alpha, c) in enumerate(nonlocal_channels):
        chans.append([l, [[], [], [[alpha, c]], [], [], [], []]])
    return [ncore, chans]


def _random_orthonormal_mos(nao, nmo, seed):
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rng.randn(nao, nao))
    return q[:, :nmo].copy()


def _random_unitary_mos(nao, nmo, seed):
    """Genuinely complex orthonormal orbitals (complex wave functions: slater.py:212-216)."""
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rng.randn(nao, nao) + 1j * rng.randn(nao, nao))
    return q[:, :nmo].copy()


def he_ccecp_pvdz(seed=7):
    """Config C1: He atom, Z_eff = 2, [2s1p] basis (A = 5), 1 up + 1 down electron."""
    basis = {
        "He": [
            _contracted(0, _even_tempered(0.35, 2.9, 5), seed),
            _single(0, 0.30),
            _single(1, 1.10),
        ]
    }
    ecp = {"He": _ecp_entry(0, 2.0, (32.0, 32.5, 33.0, -27.7), [(1.7, 0.63)])}
    mol = Mol([("He", (0.0, 0.0, 0.0))], basis, ecp, (1, 1), [2.0])
    nao = mol.nao
    c = _random_orthonormal_mos(nao, nao, seed)
    # make the occupied orbital s-like and nodeless: mostly the contracted s shell
    c[:, 0] = 0.0
    c[0, 0], c[1, 0] = 0.8, 0.3
    mo_coeff = np.array([c, c])
    occ = np.zeros((2, nao))
    occ[:, 0] = 1
    return mol, MF(mo_coeff, occ)


H2O_GEOM = [
    ("O", (0.0, 0.0, 0.0)),
    ("H", (0.0, -2.757, 2.587)),
    ("H", (0.0, 2.757, 2.587)),
]  # Bohr; geometry of benchmarks/h2o_benchmark.py:11


def h2o_basis_ecp(seed=11):
    basis = {
        "O": [
            _contracted(0, _even_tempered(0.12, 2.45, 9), seed),
            _single(0, 0.95),
            _single(0, 0.29),
            _contracted(1, _even_tempered(0.09, 2.35, 9), seed + 1),
            _single(1, 0.70),
            _single(1, 0.21),
            _single(2, 2.30),
            _single(2, 0.66),
            _single(3, 1.40),
        ],
        "H": [
            _contracted(0, _even_tempered(0.10, 2.8, 5), seed + 2),
            _single(0, 0.65),
            _single(0, 0.14),
            _single(1, 1.40),
            _single(1, 0.39),
            _single(2, 1.05),
        ],
    }
    ecp = {
        "O": _ecp_entry(2, 6.0, (12.30997, 14.76962, 13.71419, -47.876), [(13.65512, 85.86406)]),
        "H": _ecp_entry(0, 1.0, (21.24359508, 21.24359508, 21.77696655, -10.85192405), [(1.0, 0.0)]),
    }
    return basis, ecp


def h2o_ccecp_pvtz(seed=11, nmo=None, ncas=None):
    """Configs C2/C3/C5: H2O, ccECP-cc-pVTZ-shaped basis, A = 57, 4 up + 4 down.

    ``nmo`` columns of a seeded random orthonormal matrix are returned as MOs
    (default 4 = occupied only; C3 uses 8 for the CAS(8e,8o) determinant list).
    The O ECP numbers follow the published ccECP functional form; the H s-channel has a
    zero coefficient, as in ccECP, so only O contributes non-local work ... but the
    stochastic mask is still drawn for every (electron, atom) pair, as in the reference.
    """
    basis, ecp = h2o_basis_ecp(seed)
    mol = Mol(H2O_GEOM, basis, ecp, (4, 4), [6.0, 1.0, 1.0])
    nao = mol.nao
    assert nao == 57, nao
    nmo = 4 if nmo is None else nmo
    c = _random_orthonormal_mos(nao, nao, seed)
    # bias the orbitals towards the compact shells so the walkers stay near the molecule
    mo_coeff = np.array([c, c])
    occ = np.zeros((2, nao))
    occ[:, :4] = 1
    return mol, MF(mo_coeff, occ)


def cas_determinants(nelec_s, norb, seed=3, ncore=0):
    """Full CAS list ``[(weight, [occ_up, occ_dn]), ...]`` with seeded weights (config C3).

    Format of ``determinants=`` at ``pyqmc/wf/slater.py:167-179``.
    """
    import itertools

    rng = np.random.RandomState(seed)
    core = list(range(ncore))
    strings = [core + [ncore + i for i in c] for c in itertools.combinations(range(norb), nelec_s)]
    dets = []
    for iu, up in enumerate(strings):
        for idn, dn in enumerate(strings):
            w = 1.0 if (iu == 0 and idn == 0) else 0.05 * rng.randn() / (1 + 0.2 * (iu + idn))
            dets.append((float(w), [list(up), list(dn)]))
    return dets


def c2_probe(seed=5):
    """8-electron, 2-atom probe with d functions (s,p,d shells; both atoms carry ECPs with
    two non-local channels => naip = 12).  Exercises l_max = 2 channels of the ECP code."""
    basis = {
        "C": [
            _contracted(0, _even_tempered(0.11, 2.5, 7), seed),
            _single(0, 0.22),
            _contracted(1, _even_tempered(0.10, 2.4, 6), seed + 1),
            _single(1, 0.18),
            _single(2, 0.55),
        ]
    }
    ecp = {
        "C": _ecp_entry(
            2, 4.0, (14.43502, 8.39889, 7.38188, -19.25), [(7.76079, 52.13345), (8.1, -3.2)]
        )
    }
    atoms = [("C", (0.0, 0.0, -1.17)), ("C", (0.0, 0.0, 1.17))]
    mol = Mol(atoms, basis, ecp, (4, 4), [4.0, 4.0])
    nao = mol.nao
    c = _random_orthonormal_mos(nao, nao, seed)
    mo_coeff = np.array([c, c])
    occ = np.zeros((2, nao))
    occ[:, :4] = 1
    return mol, MF(mo_coeff, occ)


def open_shell_probe(seed=9):
    """3 up + 1 down electrons on H2O geometry without ECP on the hydrogens (exercises
    n_up != n_dn, all-electron atoms with the electron-ion cusp term, atoms without ECP)."""
    basis, ecp = h2o_basis_ecp(seed)
    ecp = {"O": ecp["O"]}
    mol = Mol(H2O_GEOM, basis, ecp, (3, 1), [6.0, 1.0, 1.0])
    nao = mol.nao
    c = _random_orthonormal_mos(nao, nao, seed)
    d = _random_orthonormal_mos(nao, nao, seed + 1)
    mo_coeff = np.array([c, d])
    occ = np.zeros((2, nao))
    occ[0, :3] = 1
    occ[1, :1] = 1
    return mol, MF(mo_coeff, occ)


def h_atom_like(seed=13):
    """One spin-up electron, no ECP, all-electron cusp term: exercises n_dn = 0 (empty determinant),
    an empty ECP list and Jastrow sums without partners."""
    basis = {"H": [_contracted(0, _even_tempered(0.12, 3.0, 4), seed), _single(0, 0.3), _single(1, 0.8)]}
    mol = Mol([("H", (0.1, -0.2, 0.3))], basis, {}, (1, 0), [1.0])
    nao = mol.nao
    c = _random_orthonormal_mos(nao, nao, seed)
    c[:, 0] = 0.0
    c[0, 0], c[1, 0] = 0.9, 0.2
    occ = np.zeros((2, nao))
    occ[0, 0] = 1
    return mol, MF(np.array([c, c]), occ)


def high_l_probe(seed=17):
    """He-like two-electron probe whose basis reaches l = 5 (s, p, d, f, g, h shells): exercises every branch of the
    solid-harmonic dispatch the reference has (COMPUTE_SPH_L0..L5, gto.py:107-118)."""
    basis = {
        "He": [
            _contracted(0, _even_tempered(0.35, 2.9, 4), seed),
            _single(0, 0.30),
            _single(1, 1.10),
            _single(2, 0.90),
            _single(3, 0.80),
            _single(4, 0.75),
            _single(5, 0.70),
        ]
    }
    ecp = {"He": _ecp_entry(0, 2.0, (32.0, 32.0, 33.7, -27.7), [(1.0, 0.0)])}
    mol = Mol([("He", (0.05, -0.1, 0.15))], basis, ecp, (1, 1), [2.0])
    nao = mol.nao
    assert nao == 2 + 3 + 5 + 7 + 9 + 11, nao
    rng = np.random.RandomState(seed)
    c = 0.25 * rng.randn(nao, nao)  # an occupied orbital with weight on every shell
    c[0, 0] += 1.0
    occ = np.zeros((2, nao))
    occ[:, :1] = 1
    return mol, MF(np.array([c, c]), occ)


def h2o_complex(seed=11):
    """The H2O system with complex orbital coefficients (what tests/integration/test_complex_linemin.py of the
    reference builds from a real mean field): every protocol output becomes complex."""
    mol, mf = h2o_ccecp_pvtz(seed)
    c = _random_unitary_mos(mol.nao, mol.nao, seed + 100)
    return mol, MF(np.array([c, c]), mf.mo_occ)


SYSTEMS = {
    "h2o_cx": h2o_complex,
    "hatom": h_atom_like,
    "high_l": high_l_probe,
    "he": he_ccecp_pvdz,
    "h2o": h2o_ccecp_pvtz,
    "c2": c2_probe,
    "open": open_shell_probe,
}
