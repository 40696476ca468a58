"""Synthetic periodic systems (SURVEY.md section 8d, config C4 and smaller probes).

Diamond: fcc primitive cell, a = 3.5668 Angstrom, two C atoms with a ccECP-shaped pseudopotential
(s non-local channel + local) and a [2s2p1d] basis whose exponents stay above 0.3 (the reference's
diamond fixture discards more diffuse primitives, ``exp_to_discard=0.3``).  MOs are seeded random
orthonormal matrices per k-point; the k-points are the primitive-cell points that fold onto the
supercell Gamma point (``pyqmc/pbc/supercell.py:18-30``), so all phases are real.
"""
import numpy as np

from . import pbc
from .systems import _contracted, _ecp_entry, _even_tempered, _random_orthonormal_mos, _random_unitary_mos, _single

ANG = 1.0 / 0.52917721092
A_DIAMOND = 3.5668 * ANG


def _carbon_basis_ecp(seed):
    basis = {
        "C": [
            _contracted(0, _even_tempered(0.32, 2.3, 5), seed),
            _single(0, 0.45),
            _contracted(1, _even_tempered(0.31, 2.2, 4), seed + 1),
            _single(1, 0.42),
            _single(2, 0.60),
        ]
    }
    ecp = {"C": _ecp_entry(2, 4.0, (14.43502, 8.39889, 7.38188, -19.25), [(7.76079, 52.13345)])}
    return basis, ecp


def diamond_primitive(seed=17):
    basis, ecp = _carbon_basis_ecp(seed)
    a = A_DIAMOND
    lat = 0.5 * a * np.array([[0.0, 1.0, 1.0], [1.0, 0.0, 1.0], [1.0, 1.0, 0.0]])
    atoms = [("C", (0.0, 0.0, 0.0)), ("C", (0.25 * a, 0.25 * a, 0.25 * a))]
    return pbc.Cell(atoms, basis, ecp, (4, 4), [4.0, 4.0], lat)


def _kmf(cell, supercell, nocc, seed, twist=None):
    """``twist``: fractional shift (supercell reciprocal vectors) of the whole k-mesh -- a general twist, for which
    the Bloch phases exp(i k.L) and the orbitals are genuinely complex (pyqmc/pbc/twists.py:19-25)."""
    kpts = pbc.get_supercell_kpts(supercell)
    make_mos = _random_orthonormal_mos
    if twist is not None:
        kpts = kpts + np.asarray(twist, dtype=float) @ supercell.reciprocal_vectors()
        make_mos = _random_unitary_mos
    nao = cell.nao
    mo, occ = [[], []], [[], []]
    for s in (0, 1):
        for k in range(len(kpts)):
            mo[s].append(make_mos(nao, nao, seed + 31 * k))  # same orbitals for both spins
            o = np.zeros(nao)
            o[:nocc[s]] = 1
            occ[s].append(o)
    return pbc.KMF(kpts, np.array(mo), np.array(occ))


def diamond(S=None, seed=17, twist=None):
    """(supercell, mf): diamond ``S`` supercell (default 2x2x2 = config C4: 16 atoms, 32+32
    electrons, 8 k-points with 4 occupied orbitals each per spin)."""
    S = 2 * np.eye(3, dtype=int) if S is None else np.asarray(S, dtype=int)
    cell = diamond_primitive(seed)
    sc = pbc.get_supercell(cell, S)
    return sc, _kmf(cell, sc, cell.nelec, seed, twist)


def _probe_cell(lat, seed, nelec=(2, 2), twist=None):
    """Two pseudo-carbon atoms in an arbitrary cell, 2+2 electrons (minimal-image mode probes)."""
    basis, ecp = _carbon_basis_ecp(seed)
    lat = np.asarray(lat, dtype=float)
    frac = np.array([[0.1, 0.15, 0.2], [0.55, 0.6, 0.45]])
    atoms = [("C", tuple(f @ lat)) for f in frac]
    cell = pbc.Cell(atoms, basis, ecp, nelec, [2.0, 2.0], lat)
    sc = pbc.get_supercell(cell, np.eye(3, dtype=int))
    return sc, _kmf(cell, sc, nelec, seed, twist)


TWIST = (0.21, -0.13, 0.37)  # a general (non time-reversal-invariant) twist


def orthorhombic_probe(seed=23, twist=None):
    """Diagonal lattice: the per-axis minimal image of distance.py:152-159."""
    return _probe_cell(np.diag([7.1, 7.9, 8.6]), seed, twist=twist)


def rotated_cubic_probe(seed=29):
    """Orthogonal but not axis-aligned lattice: the fractional-wrap minimal image (143-150)."""
    c, s = np.cos(0.4), np.sin(0.4)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return _probe_cell(np.diag([7.4, 7.4, 8.1]) @ R, seed)


PBC_SYSTEMS = {
    "diamond111": lambda: diamond(np.eye(3, dtype=int)),
    "diamond211": lambda: diamond(np.diag([2, 1, 1])),
    "diamond222": diamond,
    "ortho": orthorhombic_probe,
    # complex wave functions: the same cells at a general twist (complex Bloch phases, wrap phase exp(i k.R))
    "ortho_twist": lambda: orthorhombic_probe(twist=TWIST),
    "diamond211_twist": lambda: diamond(np.diag([2, 1, 1]), twist=TWIST),
    "rotcubic": rotated_cubic_probe,
}
