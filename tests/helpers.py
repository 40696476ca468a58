"""Shared builders for the parity tests: a pyqmc_b200 wave function and the numpy oracle with
identical seeded parameters."""
import numpy as np

from pyqmc_b200 import pbc_systems, systems


def jastrow_coefficients(shape_a, shape_b, has_cusp, seed, scale=0.1):
    rng = np.random.RandomState(seed)
    a0 = 1 if has_cusp else 0
    ac = scale * rng.randn(*shape_a)
    bc = scale * rng.randn(*shape_b)
    return a0, ac, bc


def make_system(name):
    if name.endswith("_3b"):
        return make_system(name[:-3])
    if name == "h2o_md":
        mol, mf = systems.h2o_ccecp_pvtz()
        dets = systems.cas_determinants(4, 6, seed=3)[:40]
        return mol, mf, dets
    if name == "h2o_md_cx":  # complex orbitals AND complex determinant coefficients
        mol, mf = systems.h2o_complex()
        dets = systems.cas_determinants(4, 6, seed=3)[:12]
        ph = np.exp(2j * np.pi * np.random.RandomState(4).rand(len(dets)))
        return mol, mf, [(w * p, occ) for (w, occ), p in zip(dets, ph)]
    if name == "h2o_cas":
        mol, mf = systems.h2o_ccecp_pvtz()
        return mol, mf, systems.cas_determinants(4, 8, seed=3)
    if name in ("ortho_md", "diamond211_md"):  # periodic multi-determinant: occupations per spin and per k-point
        mol, mf = pbc_systems.PBC_SYSTEMS[name[:-3]]()
        if name == "ortho_md":  # 2 + 2 electrons, one k-point
            dets = [(0.8, [[[0, 1]], [[0, 1]]]), (0.3, [[[0, 2]], [[0, 1]]]), (-0.25, [[[0, 1]], [[1, 3]]]),
                    (0.2, [[[1, 2]], [[0, 2]]]), (0.1, [[[2, 3]], [[2, 3]]])]
        else:  # 8 + 8 electrons, two k-points with 4 occupied orbitals each
            g = [0, 1, 2, 3]
            dets = [(0.85, [[g, g], [g, g]]), (0.3, [[[0, 1, 2, 4], g], [g, g]]), (-0.2, [[g, g], [g, [0, 1, 3, 5]]]),
                    (0.15, [[[0, 1, 2, 4], g], [[0, 2, 3, 4], g]])]
        return mol, mf, dets
    if name in pbc_systems.PBC_SYSTEMS:
        mol, mf = pbc_systems.PBC_SYSTEMS[name]()
        return mol, mf, None
    mol, mf = systems.SYSTEMS[name]()
    return mol, mf, None


def three_body_coefficients(shape, seed=2, scale=0.02):
    return scale * np.random.RandomState(seed).randn(*shape)


def make_pair(name, seed=1, jastrow=True, slater=True, three_body=None):
    """Returns (mol, mf, b200 wf, oracle wf) with the same parameters.  Systems named ``*_3b`` get a
    three-body Jastrow factor as third factor."""
    import pyqmc_b200 as pq
    from oracle.jastrow2 import JastrowOracle
    from oracle.product import ProductOracle
    from oracle.slater_det import SlaterOracle

    mol, mf, dets = make_system(name)
    factors, ofactors = [], []
    if slater:
        factors.append(pq.Slater(mol, mf, determinants=dets))
        if hasattr(mol, "a"):
            from oracle.pbc import SlaterPbcOracle

            ofactors.append(SlaterPbcOracle(mol, mf, determinants=dets))
        else:
            ofactors.append(SlaterOracle(mol, mf, determinants=dets))
    if jastrow:
        j, _ = pq.generate_jastrow(mol)
        oj = JastrowOracle.default(mol)
        has_cusp = len(j.a_basis) > 4
        a0, ac, bc = jastrow_coefficients(j.parameters["acoeff"].shape, j.parameters["bcoeff"].shape, has_cusp, seed)
        for obj in (j, oj):
            obj.parameters["acoeff"][:, a0:, :] = ac[:, a0:, :]
            obj.parameters["bcoeff"][1:, :] = bc[1:, :]
        assert np.array_equal(j.parameters["acoeff"], oj.parameters["acoeff"])
        assert np.array_equal(j.parameters["bcoeff"], oj.parameters["bcoeff"])
        factors.append(j)
        ofactors.append(oj)
    if three_body is None:
        three_body = name.endswith("_3b")
    if three_body:
        from oracle.jastrow3 import Jastrow3Oracle

        j3, _ = pq.generate_jastrow3(mol)
        oj3 = Jastrow3Oracle.default(mol)
        cc = three_body_coefficients(j3.parameters["ccoeff"].shape)
        j3.parameters["ccoeff"][...] = cc
        oj3.parameters["ccoeff"][...] = cc
        factors.append(j3)
        ofactors.append(oj3)
    if len(factors) >= 2:
        return mol, mf, pq.MultiplyWF(*factors), ProductOracle(*ofactors)
    return mol, mf, factors[0], ofactors[0]


def to_oracle_walkers(configs):
    from oracle.walkers import Walkers

    if hasattr(configs, "wrap"):
        from oracle.pbc import PeriodicWalkers

        w = PeriodicWalkers(configs.configs.copy(), configs.lvecs)
        w.wrap = configs.wrap.copy()
        return w
    return Walkers(configs.configs.copy())


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if not (np.iscomplexobj(a) or np.iscomplexobj(b)):
        a, b = a.astype(float), b.astype(float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)
