"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the reference local-energy accumulator.

Follows
  * ``EnergyAccumulator`` (``pyqmc/observables/accumulators.py:45-95``, use_old_ecp=True path);
  * open-boundary Coulomb + kinetic (``pyqmc/observables/energy.py:28-65``);
  * semi-local ECP (``pyqmc/observables/eval_ecp.py``): ``ecp`` 21-40, ``compute_tmoves`` 43-80,
    ``ecp_ea`` 83-132, ``ecp_mask`` 135-146, ``get_v_l``/``rnExp`` 149-200, ``P_l`` 203-225,
    ``get_P_l`` 228-252, ``get_rot`` 255-275, quadrature tables 278-336
    (Mitas, Shirley, Ceperley, J. Chem. Phys. 95, 3467 (1991)).

RNG order is the reference's and uses the global legacy ``np.random`` stream: per
(electron, ECP atom) first ``np.random.random(N)`` for the stochastic channel mask
(eval_ecp.py:145) and then one ``scipy Rotation.random()`` shared by all walkers
(eval_ecp.py:263) -- drawn even if no walker passes the mask.
"""
import numpy as np
import scipy.spatial.transform


def quadrature(naip):
    """Points (naip, 3) and weights (naip,) of the octahedral / icosahedral rules."""
    if naip in (6, 18, 26, 50):
        grid = np.mgrid[-1:2, -1:2, -1:2].reshape(3, -1).T
        nz = np.count_nonzero(grid, axis=1)
        A = grid[nz == 1].astype(float)
        B = grid[nz == 2] / np.sqrt(2.0)
        C = grid[nz == 3] / np.sqrt(3.0)
        d1 = C * np.sqrt(3.0 / 11.0)
        d1[:, 2] *= 3.0
        D = np.concatenate([np.roll(d1, i, axis=1) for i in range(3)])
        sets = {6: ([A], [1 / 6]), 18: ([A, B], [1 / 30, 1 / 15]),
                26: ([A, B, C], [1 / 21, 4 / 105, 27 / 840]),
                50: ([A, B, C, D], [4 / 315, 64 / 2835, 27 / 1280, 14641 / 725760])}
    elif naip in (12, 32):
        def sphere(t, p):
            return np.transpose([np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)])

        k = np.arange(10)
        b1 = np.arctan(2.0)
        s5 = 5.0**0.5
        c1 = np.arccos((2 + s5) / (15 + 6 * s5) ** 0.5)
        c2 = np.arccos(1 / (15 + 6 * s5) ** 0.5)
        A = sphere(np.array([0.0, np.pi]), np.zeros(2))
        B = sphere(np.tile([b1, np.pi - b1], 5), k * np.pi / 5)
        C = sphere(np.concatenate([np.tile([np.pi - c1, c1], 5), np.tile([np.pi - c2, c2], 5)]),
                   np.tile(k * np.pi / 5, 2))
        sets = {12: ([A, B], [1 / 12, 1 / 12]), 32: ([A, B, C], [5 / 168, 5 / 168, 27 / 840])}
    else:
        raise ValueError(f"no quadrature rule with {naip} points")
    pts, wts = sets[naip]
    points = np.concatenate(pts, axis=0)
    weights = np.concatenate([np.full(len(p), w) for p, w in zip(pts, wts)])
    return points, weights


def legendre(l, x):
    if l == 0:
        return np.ones_like(x)
    if l == 1:
        return x
    if l == 2:
        return 0.5 * (3 * x * x - 1)
    if l == 3:
        return 0.5 * (5 * x * x * x - 3 * x)
    if l == 4:
        return 0.125 * (35 * x * x * x * x - 30 * x * x + 3)
    raise NotImplementedError(l)


class EcpChannels:
    """Radial ECP channels of one species: v_l(r) = sum_t c r^(n-2) exp(-a r^2).

    Column order follows the reference's negative indexing (eval_ecp.py:154-157, 248-252):
    columns 0..lmax are the non-local channels, the LAST column is the local channel l=-1.
    """

    def __init__(self, ecp_entry):
        chans = {}
        for l, expand in ecp_entry[1]:
            terms = []
            for n, lines in enumerate(expand):
                for alpha, c in lines:
                    terms.append((n - 2, alpha, c))
            chans[int(l)] = terms
        self.nl = len(chans)
        self.lmax = self.nl - 2
        assert sorted(chans) == list(range(-1, self.lmax + 1)), sorted(chans)
        self.columns = [chans[l] for l in range(self.lmax + 1)] + [chans[-1]]

    def v_l(self, r):
        out = np.zeros((len(r), self.nl))
        for col, terms in enumerate(self.columns):
            for n, alpha, c in terms:
                out[:, col] += r ** n * c * np.exp(-alpha * r * r)
        return out


def ecp_electron_atom(channels, apos, configs, wf, e, threshold, naip=None):
    """eval_ecp.py:83-132 for one (electron, atom) pair."""
    N = configs.configs.shape[0]
    rvec = configs.configs[:, e, :] - apos
    if getattr(configs, "dist", None) is not None:  # periodic: minimal image (eval_ecp.py:94)
        rvec = configs.dist(rvec)
    r = np.linalg.norm(rvec, axis=-1)
    v = channels.v_l(r)
    # stochastic channel mask (eval_ecp.py:135-146)
    if threshold > 0:
        lodd = 2 * np.arange(channels.nl - 1) + 1
        prob = np.minimum(1.0, np.abs(v[:, :-1]) @ (threshold * (2 * lodd + 1)))
    else:
        prob = np.ones(N)
    mask = prob > np.random.random(size=N)
    mv = v[mask]
    mv[:, :-1] /= prob[mask, None]
    if naip is None:
        naip = 6 if channels.nl <= 2 else 12
    rot = scipy.spatial.transform.Rotation.random().as_matrix()
    points, weights = quadrature(naip)
    rm, rvm = r[mask], rvec[mask]
    unit = (rot @ points.T).T  # (naip, 3)
    disp = rm[:, None, None] * unit[None]  # (Nm, naip, 3)
    cosang = np.einsum("ik,ijk->ij", rvm, disp)
    cosang /= rm[:, None] * np.linalg.norm(disp, axis=-1)
    P = np.zeros((len(rm), naip, channels.nl))
    for l in range(channels.lmax + 1):
        P[:, :, l] = (2 * l + 1) * legendre(l, cosang) * weights[None]
    # column -1 (local) stays zero, as P_l(x, -1) = 0
    epos = np.repeat(configs.configs[:, e, :][:, None, :], naip, axis=1)
    epos[mask] = (configs.configs[mask, e, :] - rvm)[:, None] + disp
    epos = configs.make_irreducible(e, epos, mask)
    if np.any(mask):
        ratio = wf.testvalue(e, epos, mask)[0]
    else:
        ratio = np.zeros((0, naip))
    total = np.zeros(N, dtype=wf.dtype)  # eval_ecp.py:90
    total[mask] = np.einsum("ij,ik,ijk->i", ratio, mv, P)
    total += v[:, -1]
    return {"total": total, "v_l": mv, "local": v[:, -1], "P_l": P, "ratio": ratio,
            "epos": epos, "mask": mask}


class EnergyOracle:
    """Restatement of EnergyAccumulator (open boundary conditions, old ECP path)."""

    def __init__(self, mol, threshold=10, naip=None, **ewald_kwargs):
        self.mol = mol
        self.threshold = threshold
        self.naip = naip
        self.atoms = np.asarray(mol.atom_coords(), dtype=float)
        self.charges = np.asarray(mol.atom_charges(), dtype=float)
        self.ecp_atoms = [
            (i, EcpChannels(mol._ecp[sym]))
            for i, (sym, _) in enumerate(mol._atom) if sym in mol._ecp
        ]
        ii = 0.0
        for i in range(len(self.atoms)):
            for j in range(i + 1, len(self.atoms)):
                ii += self.charges[i] * self.charges[j] / np.linalg.norm(self.atoms[i] - self.atoms[j])
        self.ii = ii
        self.ewald = None
        if hasattr(mol, "a"):  # accumulators.py:52-55
            from .pbc import EwaldOracle

            self.ewald = EwaldOracle(mol, **ewald_kwargs)

    def ee(self, configs):
        c = configs.configs
        ne = c.shape[1]
        out = np.zeros(c.shape[0])
        for i in range(ne):
            for j in range(i + 1, ne):
                out += 1.0 / np.linalg.norm(c[:, i] - c[:, j], axis=-1)
        return out

    def ei(self, configs):
        c = configs.configs
        out = np.zeros(c.shape[0])
        for z, pos in zip(self.charges, self.atoms):
            out += -z * np.sum(1.0 / np.linalg.norm(c - pos, axis=-1), axis=1)
        return out

    def ecp(self, configs, wf):
        N, ne = configs.configs.shape[:2]
        tot = np.zeros(N, dtype=wf.dtype)  # eval_ecp.py:26
        for e in range(ne):
            per_e = np.zeros(N, dtype=wf.dtype)
            for i, ch in self.ecp_atoms:
                per_e += ecp_electron_atom(ch, self.atoms[i], configs, wf, e, self.threshold,
                                           self.naip)["total"]
            tot += per_e
        return tot

    def kinetic(self, configs, wf):
        N, ne = configs.configs.shape[:2]
        ke, grad2 = np.zeros(N), np.zeros(N)
        for e in range(ne):
            g, lap = wf.gradient_laplacian(e, configs.electron(e))
            ke += -0.5 * np.real(lap)
            grad2 += np.sum(np.abs(g) ** 2, axis=0)
        return ke, grad2

    def __call__(self, configs, wf):
        if self.ewald is not None:
            ee, ei, ii = self.ewald.energy(configs)
        else:
            ee, ei, ii = self.ee(configs), self.ei(configs), self.ii
        ecp = self.ecp(configs, wf)
        ke, grad2 = self.kinetic(configs, wf)
        return {"ke": ke, "ee": ee, "ei": ei, "ecp": ecp, "grad2": grad2,
                "total": ke + ee + ei + ecp + ii}

    def avg(self, configs, wf):
        return {k: np.mean(v, axis=0) for k, v in self(configs, wf).items()}

    def nonlocal_tmoves(self, configs, wf, e, tau):
        """eval_ecp.py:43-80."""
        N = configs.configs.shape[0]
        if not self.ecp_atoms:
            return {"ratio": np.ones((N, 0)), "weight": np.zeros((N, 0))}
        ratios, weights, positions = [], [], []
        for i, ch in self.ecp_atoms:
            d = ecp_electron_atom(ch, self.atoms[i], configs, wf, e, self.threshold, None)  # accumulators.py:80-81: no naip
            npts = d["ratio"].shape[1]
            w = np.zeros((N, npts))
            r = np.ones((N, npts), dtype=d["ratio"].dtype)
            w[d["mask"]] = np.einsum("ik,ijk->ij", np.exp(-tau * d["v_l"]) - 1, d["P_l"])
            r[d["mask"]] = d["ratio"]
            ratios.append(r)
            weights.append(w)
            positions.append(d["epos"].configs)
        return {"ratio": np.concatenate(ratios, axis=1), "weight": np.concatenate(weights, axis=1),
                "configs": np.concatenate(positions, axis=1)}

    def has_nonlocal_moves(self):
        return len(self.ecp_atoms) > 0

    def keys(self):
        return {"ke", "ee", "ei", "ecp", "total", "grad2"}

    def shapes(self):
        return {k: () for k in self.keys()}
