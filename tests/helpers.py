"""Shared builders for the parity tests: a pyqmc_b200 wave function and the numpy oracle with
identical seeded parameters."""
import numpy as np

from pyqmc_b200 import systems


def jastrow_coefficients(shape_a, shape_b, has_cusp, seed, scale=0.1):
    rng = np.random.RandomState(seed)
    a0 = 1 if has_cusp else 0
    ac = scale * rng.randn(*shape_a)
    bc = scale * rng.randn(*shape_b)
    return a0, ac, bc


def make_system(name):
    if name == "h2o_md":
        mol, mf = systems.h2o_ccecp_pvtz()
        dets = systems.cas_determinants(4, 6, seed=3)[:40]
        return mol, mf, dets
    if name == "h2o_cas":
        mol, mf = systems.h2o_ccecp_pvtz()
        return mol, mf, systems.cas_determinants(4, 8, seed=3)
    mol, mf = systems.SYSTEMS[name]()
    return mol, mf, None


def make_pair(name, seed=1, jastrow=True, slater=True):
    """Returns (mol, mf, b200 wf, oracle wf) with the same parameters."""
    import pyqmc_b200 as pq
    from oracle.jastrow2 import JastrowOracle
    from oracle.product import ProductOracle
    from oracle.slater_det import SlaterOracle

    mol, mf, dets = make_system(name)
    factors, ofactors = [], []
    if slater:
        factors.append(pq.Slater(mol, mf, determinants=dets))
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
    if len(factors) == 2:
        return mol, mf, pq.MultiplyWF(*factors), ProductOracle(*ofactors)
    return mol, mf, factors[0], ofactors[0]


def to_oracle_walkers(configs):
    from oracle.walkers import Walkers

    return Walkers(configs.configs.copy())


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)
