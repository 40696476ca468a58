"""Generates ``sph_gen.cuh``: real solid harmonics S_lm (l <= 5) and their gradients as
straight-line FP64 device code, one function per l so a shell only pays for its own l.

The functions are the ones the reference tabulates in
``pyqmc/wf/numba/spherical_harmonics.py:40-300`` (orthonormal real spherical harmonics times
r^l; l = 1 ordered x, y, z; other l ordered m = -l..l).  They are written here as
polynomials in x, y, z with closed-form normalisation constants; gradients are the
polynomial derivatives.  Run ``python gen_sph.py`` to regenerate the header.
"""
from math import comb, factorial, pi, sqrt
import os

LMAX = 5


def _poly(*terms):
    p = {}
    for c, i, j, k in terms:
        p[(i, j, k)] = p.get((i, j, k), 0.0) + c
    return p


def tables():
    t = {}
    t[0] = [_poly((0.5 / sqrt(pi), 0, 0, 0))]
    c1 = sqrt(3.0 / (4.0 * pi))
    t[1] = [_poly((c1, 1, 0, 0)), _poly((c1, 0, 1, 0)), _poly((c1, 0, 0, 1))]
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
    a3 = 0.25 * sqrt(35.0 / (2.0 * pi))
    b3 = 0.5 * sqrt(105.0 / pi)
    c3 = 0.25 * sqrt(21.0 / (2.0 * pi))
    d3 = 0.25 * sqrt(7.0 / pi)
    e3 = 0.25 * sqrt(105.0 / pi)
    t[3] = [
        _poly((3 * a3, 2, 1, 0), (-a3, 0, 3, 0)),
        _poly((b3, 1, 1, 1)),
        _poly((4 * c3, 0, 1, 2), (-c3, 2, 1, 0), (-c3, 0, 3, 0)),
        _poly((2 * d3, 0, 0, 3), (-3 * d3, 2, 0, 1), (-3 * d3, 0, 2, 1)),
        _poly((4 * c3, 1, 0, 2), (-c3, 3, 0, 0), (-c3, 1, 2, 0)),
        _poly((e3, 2, 0, 1), (-e3, 0, 2, 1)),
        _poly((a3, 3, 0, 0), (-3 * a3, 1, 2, 0)),
    ]
    a4 = 0.75 * sqrt(35.0 / pi)
    b4 = 0.75 * sqrt(35.0 / (2.0 * pi))
    c4 = 0.75 * sqrt(5.0 / pi)
    d4 = 0.75 * sqrt(5.0 / (2.0 * pi))
    e4 = (3.0 / 16.0) * sqrt(1.0 / pi)
    f4 = (3.0 / 8.0) * sqrt(5.0 / pi)
    g4 = (3.0 / 16.0) * sqrt(35.0 / pi)
    t[4] = [
        _poly((a4, 3, 1, 0), (-a4, 1, 3, 0)),
        _poly((3 * b4, 2, 1, 1), (-b4, 0, 3, 1)),
        _poly((6 * c4, 1, 1, 2), (-c4, 3, 1, 0), (-c4, 1, 3, 0)),
        _poly((4 * d4, 0, 1, 3), (-3 * d4, 2, 1, 1), (-3 * d4, 0, 3, 1)),
        _poly((8 * e4, 0, 0, 4), (-24 * e4, 2, 0, 2), (-24 * e4, 0, 2, 2),
              (3 * e4, 4, 0, 0), (6 * e4, 2, 2, 0), (3 * e4, 0, 4, 0)),
        _poly((4 * d4, 1, 0, 3), (-3 * d4, 3, 0, 1), (-3 * d4, 1, 2, 1)),
        _poly((6 * f4, 2, 0, 2), (-6 * f4, 0, 2, 2), (-f4, 4, 0, 0), (f4, 0, 4, 0)),
        _poly((b4, 3, 0, 1), (-3 * b4, 1, 2, 1)),
        _poly((g4, 4, 0, 0), (-6 * g4, 2, 2, 0), (g4, 0, 4, 0)),
    ]
    for l in range(5, LMAX + 1):  # closed form (checked against the tabulated l = 2..4 in self_check)
        t[l] = [solid_harmonic(l, m) for m in range(-l, l + 1)]
    return t


def _mul(p, q):
    out = {}
    for (a, b, c), u in p.items():
        for (d, e, f), v in q.items():
            k = (a + d, b + e, c + f)
            out[k] = out.get(k, 0.0) + u * v
    return out


def solid_harmonic(l, m):
    """Orthonormal real spherical harmonic times r^l as a polynomial {(i, j, k): c} in x^i y^j z^k:
    S_lm = N_lm Pi_l^|m|(z, r^2) A_|m|(x, y)  (m >= 0)  or  ... B_|m|(x, y)  (m < 0), with
    A_m + i B_m = (x + i y)^m and Pi_l^m = sqrt((l-m)!/(l+m)!) sum_k (-1)^k 2^-l C(l,k) C(2l-2k,l) (l-2k)!/(l-2k-m)!
    r^2k z^(l-2k-m); N = sqrt((2l+1)/4pi) (m = 0), sqrt((2l+1)/2pi) otherwise."""
    am = abs(m)
    xy = {}
    for q in range(am + 1):
        if (q % 2 == 0) == (m >= 0):
            sign = (-1) ** (q // 2)
            xy[(am - q, q, 0)] = xy.get((am - q, q, 0), 0.0) + sign * comb(am, q)
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
    p = _mul(radial, xy)
    return {k: norm * v for k, v in p.items() if abs(v) > 1e-300}


def self_check():
    """The closed form reproduces the tabulated polynomials for l = 2..4 (same ordering m = -l..l and signs)."""
    t = tables()
    for l in (2, 3, 4):
        for i, m in enumerate(range(-l, l + 1)):
            a, b = t[l][i], solid_harmonic(l, m)
            keys = set(a) | set(b)
            assert all(abs(a.get(k, 0.0) - b.get(k, 0.0)) < 1e-13 for k in keys), (l, m)


def deriv(p, axis):
    d = {}
    for (i, j, k), c in p.items():
        e = (i, j, k)[axis]
        if e == 0:
            continue
        m = [i, j, k]
        m[axis] -= 1
        d[tuple(m)] = d.get(tuple(m), 0.0) + c * e
    return d


def _pw(name, e):
    if e == 0:
        return None
    return name if e == 1 else f"{name}{e}"


def expr(p):
    if not p:
        return "0.0"
    parts = []
    for (i, j, k), c in sorted(p.items(), reverse=True):
        fac = [f for f in (_pw("x", i), _pw("y", j), _pw("z", k)) if f]
        term = repr(float(c))
        if fac:
            term += "*" + "*".join(fac)
        parts.append(term)
    return " + ".join(parts).replace("+ -", "- ")


def generate():
    t = tables()
    self_check()
    out = [f"// GENERATED by gen_sph.py -- do not edit.  Real solid harmonics, l <= {LMAX}.",
           "#pragma once", ""]
    for l in range(LMAX + 1):
        n = 2 * l + 1
        out.append(f"template <bool D> __device__ __forceinline__ void sph_l{l}(double x, double y, double z,")
        out.append(f"    double (&s)[{n}], double (&gx)[{n}], double (&gy)[{n}], double (&gz)[{n}]) {{")
        for e in range(2, l + 1):
            for v in "xyz":
                prev = v if e == 2 else f"{v}{e - 1}"
                out.append(f"  const double {v}{e} = {prev}*{v};")
        for m, p in enumerate(t[l]):
            out.append(f"  s[{m}] = {expr(p)};")
        out.append("  if (D) {")
        for m, p in enumerate(t[l]):
            for ax, g in enumerate(("gx", "gy", "gz")):
                out.append(f"    {g}[{m}] = {expr(deriv(p, ax))};")
        out.append("  }")
        out.append("}")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "sph_gen.cuh"), "w") as f:
        f.write(generate())
    print("wrote sph_gen.cuh")
