"""TEST INFRASTRUCTURE (oracle) -- real solid harmonics S_lm(x,y,z) as explicit polynomials.

Restates the functions tabulated in the reference at
``pyqmc/wf/numba/spherical_harmonics.py:40-300`` (COMPUTE_SPH_L0..L5 and their
derivative macros): orthonormal real spherical harmonics multiplied by r^l, flattened
index ``l*l + b``.  Ordering inside a shell follows the reference: l = 1 is (x, y, z)
(``spherical_harmonics.py:57-67``), every other l runs m = -l..l
(``spherical_harmonics.py:88-104`` for l = 2).

Instead of the reference's hand-factored expressions, each function is kept as a
polynomial ``{(i, j, k): coefficient}`` in x^i y^j z^k; gradients are obtained by
differentiating the polynomial.  The same tables generate the CUDA device code
(``pyqmc_b200/csrc/gen_sph.py``), so kernel and oracle share one definition that is
itself pinned against the reference functions in ``tests/test_oracle_vs_reference.py``.
"""
from math import comb, factorial, pi, sqrt

import numpy as np

LMAX = 5


def _poly(*terms):
    p = {}
    for c, i, j, k in terms:
        p[(i, j, k)] = p.get((i, j, k), 0.0) + c
    return p


def _scale(p, s):
    return {m: c * s for m, c in p.items()}


def _build_tables():
    t = {}
    # l = 0
    t[0] = [_poly((0.5 / sqrt(pi), 0, 0, 0))]
    # l = 1 : reference order x, y, z
    c1 = sqrt(3.0 / (4.0 * pi))
    t[1] = [_poly((c1, 1, 0, 0)), _poly((c1, 0, 1, 0)), _poly((c1, 0, 0, 1))]
    # l = 2 : xy, yz, 2z2-x2-y2, xz, x2-y2
    a = 0.5 * sqrt(15.0 / pi)
    b = 0.25 * sqrt(5.0 / pi)
    c = 0.25 * sqrt(15.0 / pi)
    t[2] = [
        _poly((a, 1, 1, 0)),
        _poly((a, 0, 1, 1)),
        _poly((2 * b, 0, 0, 2), (-b, 2, 0, 0), (-b, 0, 2, 0)),
        _poly((a, 1, 0, 1)),
        _poly((c, 2, 0, 0), (-c, 0, 2, 0)),
    ]
    # l = 3
    a3 = 0.25 * sqrt(35.0 / (2.0 * pi))
    b3 = 0.5 * sqrt(105.0 / pi)
    c3 = 0.25 * sqrt(21.0 / (2.0 * pi))
    d3 = 0.25 * sqrt(7.0 / pi)
    e3 = 0.25 * sqrt(105.0 / pi)
    t[3] = [
        _poly((3 * a3, 2, 1, 0), (-a3, 0, 3, 0)),  # y(3x2-y2)
        _poly((b3, 1, 1, 1)),  # xyz
        _poly((4 * c3, 0, 1, 2), (-c3, 2, 1, 0), (-c3, 0, 3, 0)),  # y(4z2-x2-y2)
        _poly((2 * d3, 0, 0, 3), (-3 * d3, 2, 0, 1), (-3 * d3, 0, 2, 1)),  # z(2z2-3x2-3y2)
        _poly((4 * c3, 1, 0, 2), (-c3, 3, 0, 0), (-c3, 1, 2, 0)),  # x(4z2-x2-y2)
        _poly((e3, 2, 0, 1), (-e3, 0, 2, 1)),  # z(x2-y2)
        _poly((a3, 3, 0, 0), (-3 * a3, 1, 2, 0)),  # x(x2-3y2)
    ]
    # l = 4
    a4 = 0.75 * sqrt(35.0 / pi)
    b4 = 0.75 * sqrt(35.0 / (2.0 * pi))
    c4 = 0.75 * sqrt(5.0 / pi)
    d4 = 0.75 * sqrt(5.0 / (2.0 * pi))
    e4 = (3.0 / 16.0) * sqrt(1.0 / pi)
    f4 = (3.0 / 8.0) * sqrt(5.0 / pi)
    g4 = (3.0 / 16.0) * sqrt(35.0 / pi)
    t[4] = [
        _poly((a4, 3, 1, 0), (-a4, 1, 3, 0)),  # xy(x2-y2)
        _poly((3 * b4, 2, 1, 1), (-b4, 0, 3, 1)),  # yz(3x2-y2)
        _poly((6 * c4, 1, 1, 2), (-c4, 3, 1, 0), (-c4, 1, 3, 0)),  # xy(6z2-x2-y2)
        _poly((4 * d4, 0, 1, 3), (-3 * d4, 2, 1, 1), (-3 * d4, 0, 3, 1)),  # yz(4z2-3x2-3y2)
        # 35z4 - 30 z2 r2 + 3 r4 = 8z4 - 24 z2(x2+y2) + 3(x2+y2)^2
        _poly((8 * e4, 0, 0, 4), (-24 * e4, 2, 0, 2), (-24 * e4, 0, 2, 2),
              (3 * e4, 4, 0, 0), (6 * e4, 2, 2, 0), (3 * e4, 0, 4, 0)),
        _poly((4 * d4, 1, 0, 3), (-3 * d4, 3, 0, 1), (-3 * d4, 1, 2, 1)),  # xz(4z2-3x2-3y2)
        # (x2-y2)(6z2-x2-y2)
        _poly((6 * f4, 2, 0, 2), (-6 * f4, 0, 2, 2), (-f4, 4, 0, 0), (f4, 0, 4, 0)),
        _poly((b4, 3, 0, 1), (-3 * b4, 1, 2, 1)),  # xz(x2-3y2)
        _poly((g4, 4, 0, 0), (-6 * g4, 2, 2, 0), (g4, 0, 4, 0)),  # x4-6x2y2+y4
    ]
    # l = 5 (COMPUTE_SPH_L5, spherical_harmonics.py:270-300): from the closed form; the same closed form
    # reproduces the tabulated l = 2..4 above (tests/test_oracle_golden.py) and the reference's SPH5
    t[5] = [closed_form(5, m) for m in range(-5, 6)]
    return t


def _mul(p, q):
    out = {}
    for (a, b, c), u in p.items():
        for (d, e, f), v in q.items():
            k = (a + d, b + e, c + f)
            out[k] = out.get(k, 0.0) + u * v
    return out


def closed_form(l, m):
    """S_lm = N_lm Pi_l^|m|(z, r^2) {Re, Im}(x + iy)^|m| as a polynomial (real solid harmonic, orthonormal on
    the sphere, m >= 0 takes the real part, m < 0 the imaginary part)."""
    am = abs(m)
    xy = {}
    for q in range(am + 1):
        if (q % 2 == 0) == (m >= 0):
            xy[(am - q, q, 0)] = xy.get((am - q, q, 0), 0.0) + (-1) ** (q // 2) * comb(am, q)
    r2 = {(2, 0, 0): 1.0, (0, 2, 0): 1.0, (0, 0, 2): 1.0}
    radial = {}
    for k in range((l - am) // 2 + 1):
        c = (-1) ** k * 2.0 ** (-l) * comb(l, k) * comb(2 * l - 2 * k, l) * factorial(l - 2 * k) / factorial(l - 2 * k - am)
        term = {(0, 0, l - 2 * k - am): c}
        for _ in range(k):
            term = _mul(term, r2)
        for key, v in term.items():
            radial[key] = radial.get(key, 0.0) + v
    norm = sqrt((2 * l + 1) / (4 * pi)) if m == 0 else sqrt((2 * l + 1) / (2 * pi))
    norm *= sqrt(factorial(l - am) / factorial(l + am))
    return {k: norm * v for k, v in _mul(radial, xy).items() if abs(v) > 1e-300}


TABLES = _build_tables()


def poly_derivative(p, axis):
    d = {}
    for (i, j, k), c in p.items():
        e = (i, j, k)[axis]
        if e == 0:
            continue
        m = [i, j, k]
        m[axis] -= 1
        m = tuple(m)
        d[m] = d.get(m, 0.0) + c * e
    return d


def flat_polys(lmax):
    """List of polynomials in flattened (l*l + b) order up to lmax."""
    out = []
    for l in range(lmax + 1):
        out.extend(TABLES[l])
    return out


def _eval_poly(p, pw):
    acc = 0.0
    for (i, j, k), c in p.items():
        acc = acc + c * pw[0][i] * pw[1][j] * pw[2][k]
    return acc


def _powers(v, lmax):
    pw = [np.ones_like(v)]
    for _ in range(lmax):
        pw.append(pw[-1] * v)
    return pw


def evaluate(lmax, x, y, z, deriv=False):
    """S[(lmax+1)^2, ...] and (if deriv) dS[3, (lmax+1)^2, ...] at the points (x, y, z)."""
    if lmax > LMAX:
        raise NotImplementedError(f"solid harmonics only tabulated to l={LMAX}")
    pw = (_powers(x, lmax), _powers(y, lmax), _powers(z, lmax))
    polys = flat_polys(lmax)
    shape = np.shape(x)
    S = np.zeros((len(polys),) + shape)
    for n, p in enumerate(polys):
        S[n] = _eval_poly(p, pw)
    if not deriv:
        return S
    dS = np.zeros((3, len(polys)) + shape)
    for n, p in enumerate(polys):
        for ax in range(3):
            dS[ax, n] = _eval_poly(poly_derivative(p, ax), pw)
    return S, dS
