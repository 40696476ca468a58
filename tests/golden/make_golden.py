"""Generates the committed golden vectors by running the UNMODIFIED reference
(/root/reference, numba GTO evaluator) on the synthetic systems of ``pyqmc_b200.systems``.

    python tests/golden/make_golden.py        # writes tests/golden/<system>.npz

Only this script (and refload.py) touch /root/reference; the fixtures travel to the GPU box.
Each file holds the inputs (walkers, trial positions, masks, Jastrow coefficients, RNG seeds)
and the reference's outputs for every wf-protocol call, the energy accumulator and a short
VMC run.  The reference's numba kernels use fastmath, so outputs are reproducible to rounding
(~1e-13 relative), not bit-for-bit across LLVM versions -- tests compare at 1e-10.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import helpers  # noqa: E402
import refload  # noqa: E402

NCONF = 12
SYSTEMS = ["he", "h2o", "open", "c2", "h2o_md", "h2o_3b", "h2o_md_3b", "h2o_cas", "h2o_cas_3b", "high_l",
           "h2o_cx", "h2o_md_cx"]  # *_cx: complex orbital (and determinant) coefficients
PBC_SYSTEMS = ["diamond211", "ortho", "rotcubic", "diamond211_3b", "ortho_3b",
               "ortho_twist", "diamond211_twist",  # *_twist: general twist = complex Bloch phases
               "ortho_md", "diamond211_md"]  # periodic multi-determinant expansions (occupations per k-point)
EWALD_GMAX = 10  # the reference enumerates (2 gmax + 1)^3 / 2 reciprocal points: keep the fixture run small


def build_reference(name):
    mol, mf, dets = helpers.make_system(name)
    three_body = name.endswith("_3b")
    wf = refload.build_reference_wf(mol, mf, determinants=dets, seed=0, three_body=three_body)
    if three_body:
        j3 = wf.wf_factors[2]
        j3.parameters["ccoeff"][...] = helpers.three_body_coefficients(j3.parameters["ccoeff"].shape)
    # same seeded Jastrow coefficients as helpers.make_pair
    jast = wf.wf_factors[1]
    has_cusp = len(jast.a_basis) > 4
    a0, ac, bc = helpers.jastrow_coefficients(jast.parameters["acoeff"].shape, jast.parameters["bcoeff"].shape,
                                              has_cusp, 1)
    jast.parameters["acoeff"][:, a0:, :] = ac[:, a0:, :]
    jast.parameters["bcoeff"][1:, :] = bc[1:, :]
    return mol, wf


def generate(name):
    import pyqmc.method.mc as mc
    from pyqmc.observables.accumulators import EnergyAccumulator

    periodic = name in PBC_SYSTEMS
    mol, wf = build_reference(name)
    ekw = {"ewald_gmax": EWALD_GMAX} if periodic else {}
    out = {}
    np.random.seed(3)
    configs = mc.initial_guess(mol, NCONF)
    out["configs0"] = configs.configs.copy()
    if periodic:
        out["wrap0"] = configs.wrap.copy()
    out["acoeff"] = np.array(wf.parameters["wf2acoeff"])
    out["bcoeff"] = np.array(wf.parameters["wf2bcoeff"])
    s, l = wf.recompute(configs)
    out["recompute_sign"], out["recompute_log"] = s, l
    ne = configs.configs.shape[1]
    rng = np.random.RandomState(5)
    elist = sorted({0, ne // 2, ne - 1})
    out["elist"] = np.array(elist)
    for i, e in enumerate(elist):
        newpos = configs.configs[:, e] + 0.3 * rng.randn(NCONF, 3)
        mask = rng.rand(NCONF) > 0.4
        aux = configs.configs[:, e][:, None, :] + 0.2 * rng.randn(NCONF, 6, 3)
        ep = configs.make_irreducible(e, newpos)
        out[f"q{i}_newpos"], out[f"q{i}_mask"], out[f"q{i}_aux"] = newpos, mask, aux
        out[f"q{i}_gradient"] = wf.gradient(e, ep)
        g, v, saved = wf.gradient_value(e, ep)
        out[f"q{i}_gv_grad"], out[f"q{i}_gv_val"] = g, v
        g, lap = wf.gradient_laplacian(e, ep)
        out[f"q{i}_gl_grad"], out[f"q{i}_gl_lap"] = g, lap
        out[f"q{i}_testvalue"] = wf.testvalue(e, ep)[0]
        out[f"q{i}_testvalue_mask"] = wf.testvalue(e, ep, mask)[0]
        out[f"q{i}_testvalue_aux"] = wf.testvalue(e, configs.make_irreducible(e, aux), mask)[0]
        out[f"q{i}_testvalue_many"] = wf.testvalue_many(np.arange(ne), ep)
        g, v, saved = wf.gradient_value(e, ep)
        # update BEFORE moving configs (order of the reference's harness, testwf.py:116-119): the
        # reference's three-body factor reads the old position of electron e from `configs`
        wf.updateinternals(e, ep, configs, mask=mask, saved_values=saved)
        configs.move(e, ep, mask)
        s, l = wf.value()
        out[f"q{i}_value_sign"], out[f"q{i}_value_log"] = s, l
    out["configs1"] = configs.configs.copy()
    sl, ja = wf.wf_factors[0], wf.wf_factors[1]
    for spin in (0, 1):
        out[f"inverse{spin}"] = np.array(sl._inverse[spin])
        out[f"dets{spin}"] = np.array(sl._dets[spin])
    out["a_partial"], out["b_partial"] = np.array(ja._a_partial), np.array(ja._b_partial)
    if len(wf.wf_factors) > 2:
        out["P_i"], out["a3_values"] = np.array(wf.wf_factors[2].P_i), np.array(wf.wf_factors[2].a_values)
    pg = wf.pgradient()
    for k in pg.keys():
        out["pgrad_" + k] = np.array(pg[k])
    if periodic:
        out["wrap1"] = configs.wrap.copy()
    np.random.seed(21)
    en = EnergyAccumulator(mol, **ekw)(configs, wf)
    for k, v in en.items():
        out["energy_" + k] = np.asarray(v)
    np.random.seed(22)
    tm = EnergyAccumulator(mol, **ekw).nonlocal_tmoves(configs, wf, elist[-1], 0.02)
    out["tmove_ratio"], out["tmove_weight"] = tm["ratio"], tm["weight"]
    out["tmove_configs"] = tm["configs"].configs
    if len(wf.wf_factors) > 2:
        # the reference's drivers call updateinternals after configs.move (mc.py:135-136), which
        # leaves its three-body cache inconsistent; no VMC golden for these systems
        return out
    # short VMC run (recording the accept masks through a thin wrapper around updateinternals)
    accepts = []
    orig = wf.updateinternals

    def spy(e, epos, cfg, mask=None, saved_values=None):
        accepts.append(np.array(mask))
        return orig(e, epos, cfg, mask=mask, saved_values=saved_values)

    wf.updateinternals = spy
    np.random.seed(31)
    df, configs = mc.vmc(wf, configs, tstep=0.5, nblocks=2, nsteps_per_block=3,
                         accumulators={"energy": EnergyAccumulator(mol, **ekw)})
    wf.updateinternals = orig
    out["vmc_accept"] = np.array(accepts).reshape(2, 3, ne, NCONF)
    out["vmc_configs"] = configs.configs.copy()
    if periodic:
        out["vmc_wrap"] = configs.wrap.copy()
    for k in ("energytotal", "energyke", "energyecp", "energyee", "energyei", "energygrad2", "acceptance"):
        out["vmc_" + k] = df[k]
    if name in ("h2o", "c2", "open", "h2o_md", "ortho", "diamond211", "ortho_md", "rotcubic", "diamond211_md"):
        # DMC propagation with T-moves through the reference's own dmc_propagate (dmc.py:123-221)
        import pyqmc.method.dmc as dmc

        out["dmc_configs0"] = configs.configs.copy()
        if periodic:
            out["dmc_wrap0"] = configs.wrap.copy()
        weights = np.ones(NCONF)
        np.random.seed(41)
        dret, configs, weights = dmc.dmc_propagate(wf, configs, weights, 0.02, 10.0, 1.5, 1.7, nsteps=3,
                                                  accumulators={"energy": EnergyAccumulator(mol, **ekw)})
        out["dmc_configs"] = configs.configs.copy()
        if periodic:
            out["dmc_wrap"] = configs.wrap.copy()
        out["dmc_weights"] = weights.copy()
        for k, v in dret.items():
            out["dmc_" + k] = np.asarray(v)
    return out


SR_SYSTEMS = ["h2o", "h2o_md"]


def sr_inputs(parameters, nconf):
    """Deterministic optimisation mask (everything but the cusp row of bcoeff) and walker weights."""
    to_opt = {}
    for k in parameters.keys():
        m = np.ones(np.shape(parameters[k]), dtype=bool)
        if k.endswith("bcoeff"):
            m[0, :] = False
        if "mo_coeff" in k:  # a slice of the orbital coefficients keeps the P x P fixture small
            m[:] = False
            m[3:9, :] = True
        to_opt[k] = m
    return to_opt, 0.5 + np.random.RandomState(4).rand(nconf)


def generate_sr(name):
    """StochasticReconfiguration.avg of the reference (stochastic_reconfiguration.py:85-118) on the walkers
    `configs1` of the main golden file -> tests/golden/sr_<name>.npz."""
    import pyqmc.configurations.coord as coord
    from pyqmc.observables.accumulators import EnergyAccumulator, LinearTransform
    from pyqmc.observables.stochastic_reconfiguration import StochasticReconfiguration

    mol, wf = build_reference(name)
    data = dict(np.load(os.path.join(HERE, f"{name}.npz")))
    configs = coord.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    to_opt, weights = sr_inputs(wf.parameters, len(configs.configs))
    acc = StochasticReconfiguration(EnergyAccumulator(mol), LinearTransform(wf.parameters, to_opt))
    np.random.seed(77)
    d = acc.avg(configs, wf, weights=weights)
    return {"sr_" + k: np.asarray(v) for k, v in d.items()}


RUNDMC_SYSTEMS = ["h2o", "c2"]


def generate_rundmc(name, nconf=16):
    """The reference's whole DMC driver (dmc.py:413-591: VMC warm-up without accumulators, energy
    reference, propagation with T-moves, branching, e_trial feedback) -> tests/golden/rundmc_<name>.npz."""
    import pyqmc.method.dmc as dmc
    import pyqmc.method.mc as mc
    from pyqmc.observables.accumulators import EnergyAccumulator

    mol, wf = build_reference(name)
    np.random.seed(51)
    configs = mc.initial_guess(mol, nconf)
    out = {"configs0": configs.configs.copy()}
    np.random.seed(52)
    df, configs, weights = dmc.rundmc(wf, configs, tstep=0.02, nblocks=3, nsteps_per_block=2, vmc_warmup=2,
                                      accumulators={"energy": EnergyAccumulator(mol)})
    out["configs"], out["weights"] = configs.configs.copy(), np.asarray(weights)
    for k, v in df.items():
        out["df_" + k] = np.asarray(v)
    return out


def obdm_inputs(mol, mf):
    """Orbitals of the density matrix: the first 6 spin-up MOs of the synthetic mean field."""
    return np.ascontiguousarray(np.asarray(mf.mo_coeff[0])[:, :6])


def generate_obdm(name="h2o"):
    """OBDMAccumulator of the reference (obdm.py:25-214; in-tree numba orbital evaluator) on the walkers
    `configs1` of the main golden file -> tests/golden/obdm_<name>.npz."""
    import pyqmc.configurations.coord as coord
    import pyqmc.wf.orbitals
    from pyqmc.observables.obdm import OBDMAccumulator

    mol, wf = build_reference(name)
    _, mf, _ = helpers.make_system(name)
    data = dict(np.load(os.path.join(HERE, f"{name}.npz")))
    configs = coord.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    c = obdm_inputs(mol, mf)
    out = {}
    for tag, kw in (("all", {}), ("up", {"spin": 0})):
        acc = OBDMAccumulator(mol, c, nsweeps=3, tstep=0.5, warmup=25, **kw)
        acc.orbitals = pyqmc.wf.orbitals.MoleculeOrbitalEvaluator(mol, [c, c], evaluate_orbitals_with="numba")
        np.random.seed(61)
        first = acc(configs, wf)
        second = acc.avg(configs, wf)  # continues the auxiliary walk
        out[f"{tag}_value"], out[f"{tag}_norm"] = first["value"], first["norm"]
        out[f"{tag}_avg_value"], out[f"{tag}_avg_norm"] = second["value"], second["norm"]
    return out


def generate_tbdm(name="h2o"):
    """TBDMAccumulator of the reference (tbdm.py:26-282; in-tree numba orbital evaluator) on the walkers `configs1`
    of the main golden file, sectors up-down (full index list) and up-up (a chosen list) -> tests/golden/tbdm_<name>.npz."""
    import pyqmc.configurations.coord as coord
    import pyqmc.wf.orbitals
    from pyqmc.observables.tbdm import TBDMAccumulator

    mol, wf = build_reference(name)
    mol.cart = False
    _, mf, _ = helpers.make_system(name)
    data = dict(np.load(os.path.join(HERE, f"{name}.npz")))
    configs = coord.OpenConfigs(data["configs1"].copy())
    wf.recompute(configs)
    cu = np.ascontiguousarray(np.asarray(mf.mo_coeff[0])[:, :3])
    cd = np.ascontiguousarray(np.asarray(mf.mo_coeff[1])[:, 1:5])
    out = {"ijkl_uu": np.array([[0, 0, 0, 0], [0, 1, 1, 0], [2, 1, 0, 2], [1, 1, 2, 2], [2, 0, 1, 1]])}
    for tag, spin, ijkl in (("ud", (0, 1), None), ("uu", (0, 0), out["ijkl_uu"])):
        orb = [cu, cd] if tag == "ud" else [cu, cu]
        acc = TBDMAccumulator(mol, orb, spin, nsweeps=2, tstep=0.5, warmup=7, ijkl=ijkl)
        acc.orbitals = pyqmc.wf.orbitals.MoleculeOrbitalEvaluator(mol, orb, evaluate_orbitals_with="numba")
        np.random.seed(71)
        first = acc(configs, wf)
        second = acc.avg(configs, wf)  # continues both auxiliary walks
        for k in ("value", "norm_a", "norm_b"):
            out[f"{tag}_{k}"], out[f"{tag}_avg_{k}"] = first[k], second[k]
    return out


def main():
    warnings.filterwarnings("ignore")
    refload.load()
    if len(sys.argv) > 1 and sys.argv[1] == "tbdm":
        path = os.path.join(HERE, "tbdm_h2o.npz")
        np.savez_compressed(path, **generate_tbdm("h2o"))
        print("wrote", path)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "obdm":
        path = os.path.join(HERE, "obdm_h2o.npz")
        np.savez_compressed(path, **generate_obdm("h2o"))
        print("wrote", path)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "rundmc":
        for name in RUNDMC_SYSTEMS:
            path = os.path.join(HERE, f"rundmc_{name}.npz")
            np.savez_compressed(path, **generate_rundmc(name))
            print("wrote", path)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sr":
        for name in SR_SYSTEMS:
            path = os.path.join(HERE, f"sr_{name}.npz")
            np.savez_compressed(path, **generate_sr(name))
            print("wrote", path)
        return
    names = sys.argv[1:] if len(sys.argv) > 1 else SYSTEMS + PBC_SYSTEMS
    for name in names:
        data = generate(name)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"wrote {path}: {len(data)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
