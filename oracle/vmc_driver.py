"""TEST INFRASTRUCTURE (oracle) -- restatement of the reference VMC driver loop.

Follows ``pyqmc/method/mc.py``: ``initial_guess`` 25-73, ``limdrift`` 76-89, ``vmc_worker``
102-153 (drift-diffusion proposal, reverse-move probability, Metropolis test, masked
internal update, accumulator averaging).  RNG: global legacy ``np.random`` in the
reference's order -- per electron ``normal(N,3)`` then ``rand(N)`` (mc.py:119,132).

``record`` (optional list) receives per-(step, electron) dictionaries with the accept mask and
the intermediate quantities, which is what the parity tests compare.
"""
import numpy as np

from .walkers import Walkers


def initial_guess(mol, nconfig, r=1.0):
    ne = int(np.sum(mol.nelec))
    epos = np.zeros((nconfig, ne, 3))
    wts = mol.atom_charges()
    wts = wts / np.sum(wts)
    coords = mol.atom_coords()
    for s in (0, 1):
        neach = np.array(np.floor(mol.nelec[s] * wts), dtype=int)
        nassigned = int(np.sum(neach))
        left = int(mol.nelec[s] - nassigned)
        lo = s * mol.nelec[0]
        epos[:, lo : lo + nassigned, :] = np.repeat(coords, neach, axis=0)
        if left > 0:
            inds = np.argpartition(np.random.random((nconfig, len(wts))), left, axis=1)[:, :left]
            epos[:, lo + nassigned : lo + mol.nelec[s], :] = coords[inds]
    epos += r * np.random.randn(*epos.shape)
    if hasattr(mol, "a"):  # mc.py:69-70
        from .pbc import PeriodicWalkers

        return PeriodicWalkers(epos, mol.lattice_vectors())
    return Walkers(epos)


def limdrift(g, cutoff=1.0):
    tot = np.linalg.norm(g, axis=1)
    big = tot > cutoff
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(big[:, None], cutoff * g / tot[:, None], g)


def vmc_step_moves(wf, configs, tstep, record=None):
    """One sweep over all electrons; returns the mean acceptance (mc.py:115-137)."""
    N, ne, _ = configs.configs.shape
    acc = 0.0
    for e in range(ne):
        g, _, _ = wf.gradient_value(e, configs.electron(e))
        grad = limdrift(np.real(g.T))
        gauss = np.random.normal(scale=np.sqrt(tstep), size=(N, 3))
        new = configs.make_irreducible(e, configs.configs[:, e, :] + gauss + grad * tstep)
        g, val, saved = wf.gradient_value(e, new)
        new_grad = limdrift(np.real(g.T))
        forward = np.sum(gauss**2, axis=1)
        backward = np.sum((gauss + tstep * (grad + new_grad)) ** 2, axis=1)
        t_prob = np.exp(1 / (2 * tstep) * (forward - backward))
        ratio = np.abs(val) ** 2 * t_prob
        u = np.random.rand(N)
        accept = ratio > u
        if record is not None:
            record.append({"e": e, "accept": accept.copy(), "ratio": ratio.copy(), "u": u,
                           "newpos": new.configs.copy()})
        configs.move(e, new, accept)
        wf.updateinternals(e, new, configs, mask=accept, saved_values=saved)
        acc += np.mean(accept) / ne
    return acc


def vmc_worker(wf, configs, tstep, nsteps, accumulators, record=None):
    block_avg = {}
    wf.recompute(configs)
    for _ in range(nsteps):
        acc = vmc_step_moves(wf, configs, tstep, record)
        for k, accumulator in accumulators.items():
            dat = accumulator.avg(configs, wf)
            for m, res in dat.items():
                block_avg[k + m] = block_avg.get(k + m, 0.0) + res / nsteps
        block_avg["acceptance"] = acc
    return block_avg, configs


def vmc(wf, configs, tstep=0.5, nblocks=10, nsteps_per_block=10, accumulators=None, record=None):
    accumulators = {} if accumulators is None else accumulators
    rows = []
    for block in range(nblocks):
        avg, configs = vmc_worker(wf, configs, tstep, nsteps_per_block, accumulators, record)
        avg["block"] = block
        avg["nconfig"] = nsteps_per_block * configs.configs.shape[0]
        rows.append(avg)
    out = {k: np.asarray([r[k] for r in rows]) for k in rows[0]} if rows else {}
    return out, configs
