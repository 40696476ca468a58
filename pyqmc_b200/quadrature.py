"""Spherical quadrature rules for the semi-local ECP integration.

Octahedral (6, 18, 26, 50 points) and icosahedral (12, 32 points) rules of Mitas, Shirley and
Ceperley, J. Chem. Phys. 95, 3467 (1991), in the point order of the reference tables
(``generate_quadrature_grids``, ``pyqmc/observables/eval_ecp.py:278-336``) -- the order matters
because T-move candidates are indexed by quadrature point.
"""
import functools

import numpy as np


@functools.lru_cache(maxsize=None)
def grid(naip):
    if naip in (6, 18, 26, 50):
        cube = np.mgrid[-1:2, -1:2, -1:2].reshape(3, -1).T
        nz = np.count_nonzero(cube, axis=1)
        face = cube[nz == 1].astype(float)
        edge = cube[nz == 2] / np.sqrt(2.0)
        corner = cube[nz == 3] / np.sqrt(3.0)
        d1 = corner * np.sqrt(3.0 / 11.0)
        d1[:, 2] *= 3.0
        extra = np.concatenate([np.roll(d1, i, axis=1) for i in range(3)])
        rules = {
            6: ([face], [1 / 6]),
            18: ([face, edge], [1 / 30, 1 / 15]),
            26: ([face, edge, corner], [1 / 21, 4 / 105, 27 / 840]),
            50: ([face, edge, corner, extra], [4 / 315, 64 / 2835, 27 / 1280, 14641 / 725760]),
        }
    elif naip in (12, 32):
        def on_sphere(theta, phi):
            return np.transpose([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])

        k = np.arange(10)
        root5 = 5.0 ** 0.5
        tb = np.arctan(2.0)
        tc1 = np.arccos((2 + root5) / (15 + 6 * root5) ** 0.5)
        tc2 = np.arccos(1 / (15 + 6 * root5) ** 0.5)
        poles = on_sphere(np.array([0.0, np.pi]), np.zeros(2))
        ring = on_sphere(np.tile([tb, np.pi - tb], 5), k * np.pi / 5)
        caps = on_sphere(np.concatenate([np.tile([np.pi - tc1, tc1], 5), np.tile([np.pi - tc2, tc2], 5)]),
                         np.tile(k * np.pi / 5, 2))
        rules = {12: ([poles, ring], [1 / 12, 1 / 12]), 32: ([poles, ring, caps], [5 / 168, 5 / 168, 27 / 840])}
    else:
        raise ValueError(f"Possible AIPs are one of (6, 12, 18, 26, 32, 50), not {naip}")
    pts, wts = rules[naip]
    points = np.concatenate(pts, axis=0)
    weights = np.concatenate([np.full(len(p), w) for p, w in zip(pts, wts)])
    return points, weights
