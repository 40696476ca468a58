"""TEST INFRASTRUCTURE (oracle) -- restatement of the reference DMC propagation loop.

Follows ``pyqmc/method/dmc.py``: Umrigar drift limiter ``limdrift`` 22-35, ``get_V2`` 38-46,
``propose_drift_diffusion`` 49-70 (fixed-node rejection for real wave functions),
``propose_tmoves`` 73-120, ``dmc_propagate`` 123-221, ``compute_S`` 224-235, ``branch`` 342-376.
RNG: the global legacy ``np.random`` stream in the reference's order (per T-move electron: the
ECP draws of ``nonlocal_tmoves``, one scalar ``rand()`` per walker in ``select_walker``, then
``rand(N)``; per diffusion electron ``normal(N,3)`` then ``rand(N)``).

Works with any wave function / accumulator pair that follows the protocol (the oracle objects or
the pyqmc_b200 device objects), which is how the parity tests use it.
"""
import numpy as np


def limdrift(g, tau, acyrus=0.5):
    v2 = np.sum(g**2, axis=1)
    big = v2 > 1e-8
    taueff = np.ones(v2.shape) * tau
    taueff[big] = (np.sqrt(1 + 2 * tau * acyrus * v2[big]) - 1) / (acyrus * v2[big])
    return g * taueff[:, np.newaxis]


def get_v2(configs, wf, acc_out):
    if "grad2" in acc_out:
        return acc_out["grad2"]
    N, ne = configs.configs.shape[:2]
    v2 = np.zeros(N)
    for e in range(ne):
        v2 += np.sum(np.abs(wf.gradient(e, configs.electron(e))).T ** 2, axis=1)
    return v2


def propose_drift_diffusion(wf, configs, tstep, e):
    N = configs.configs.shape[0]
    grad = limdrift(np.real(wf.gradient(e, configs.electron(e)).T), tstep)
    gauss = np.random.normal(scale=np.sqrt(tstep), size=(N, 3))
    newepos = configs.make_irreducible(e, configs.configs[:, e, :] + gauss + grad)
    g, ratio_wf, saved = wf.gradient_value(e, newepos)
    new_grad = limdrift(np.real(g.T), tstep)
    forward = np.sum(gauss**2, axis=1)
    backward = np.sum((gauss + grad + new_grad) ** 2, axis=1)
    t_prob = np.exp(1 / (2 * tstep) * (forward - backward))
    ratio = np.abs(ratio_wf) ** 2 * t_prob
    ratio = ratio * np.sign(ratio_wf)  # fixed node (real wave function)
    accept = ratio > np.random.rand(N)
    r2 = np.sum((gauss + grad) ** 2, axis=1)
    return newepos, accept, r2, saved


def propose_tmoves(wf, configs, energy_accumulator, tstep, e):
    moves = energy_accumulator.nonlocal_tmoves(configs, wf, e, tstep)
    amp = moves["ratio"] * moves["weight"]
    fwd = np.zeros_like(amp)
    fwd[amp > 0] = amp[amp > 0]
    norm = 1.0 + np.sum(fwd, axis=1)
    cdf = np.cumsum(fwd / norm[:, None], axis=1)
    selected = np.array([np.searchsorted(row, np.random.rand()) for row in cdf], dtype=int).reshape(len(cdf))
    chosen = selected < amp.shape[1]
    cand = moves["configs"].configs if hasattr(moves["configs"], "configs") else moves["configs"]
    newpos = np.zeros((len(norm), 3))
    back = amp.copy()
    for w, mv in enumerate(selected):
        if chosen[w]:
            newpos[w] = cand[w, mv]
            rev = 1.0 / moves["ratio"][w, mv]
            back[w, :] *= rev
            back[w, mv] = rev * moves["weight"][w, mv]
        else:
            newpos[w] = configs.configs[w, e]
    back[back < 0] = 0.0
    acceptance = norm / (1.0 + np.sum(back, axis=1))
    acceptance[~chosen] = 0.0
    return configs.make_irreducible(e, newpos), chosen, acceptance, np.sum(amp)


def compute_s(e_trial, e_est, branchcut, v2, tau, eloc, nelec):
    e_cut = e_est - eloc
    big = np.abs(e_cut) > branchcut
    e_cut[big] = branchcut * np.sign(e_cut[big])
    return e_trial - e_est + e_cut / np.sqrt(1 + (v2 * tau / nelec) ** 2)


def dmc_propagate(wf, configs, weights, tstep, branchcut_start, e_trial, e_est, nsteps=5, accumulators=None,
                  ekey=("energy", "total"), record=None):
    N, ne = configs.configs.shape[:2]
    wf.recompute(configs)
    energy = accumulators[ekey[0]]
    dat = energy(configs, wf)
    eloc = np.real(dat[ekey[1]])
    v2 = get_v2(configs, wf, dat)
    rows = []
    for _ in range(nsteps):
        r2_acc, r2_prop = np.zeros(N), np.zeros(N)
        p_acc, t_acc = np.zeros(N), np.zeros(N)
        if energy.has_nonlocal_moves():
            for e in range(ne):
                newepos, chosen, prob, _ = propose_tmoves(wf, configs, energy, tstep, e)
                accept = chosen & (prob > np.random.rand(N))
                if record is not None:
                    record.append(("t", e, accept.copy()))
                configs.move(e, newepos, accept)
                wf.updateinternals(e, newepos, configs, mask=accept)
                t_acc += accept / ne
        for e in range(ne):
            newepos, accept, r2, saved = propose_drift_diffusion(wf, configs, tstep, e)
            if record is not None:
                record.append(("d", e, accept.copy()))
            configs.move(e, newepos, accept)
            wf.updateinternals(e, newepos, configs, mask=accept, saved_values=saved)
            r2_prop += r2
            r2_acc[accept] += r2[accept]
            p_acc += accept / ne
        eloc_old, v2_old = eloc.copy(), v2.copy()
        dat = energy(configs, wf)
        eloc = np.real(dat[ekey[1]])
        tdamp = r2_acc / r2_prop
        v2 = get_v2(configs, wf, dat)
        s_new = compute_s(e_trial, e_est, branchcut_start, v2, tstep, eloc, ne)
        s_old = compute_s(e_trial, e_est, branchcut_start, v2_old, tstep, eloc_old, ne)
        weights *= np.exp(tstep * tdamp * (0.5 * s_new + 0.5 * s_old))
        wavg = np.mean(weights)
        avg = {}
        for k, acc in accumulators.items():
            d = acc(configs, wf) if k != ekey[0] else dat
            for m, res in d.items():
                avg[k + m] = np.einsum("...i,i...->...", weights, res) / (N * wavg)
        avg["weight"] = wavg
        avg["acceptance"] = np.mean(p_acc)
        avg["tmove_acceptance"] = np.mean(t_acc)
        rows.append(avg)
    wts = np.asarray([r["weight"] for r in rows])
    rel = wts / np.mean(wts)
    out = {k: np.mean([r[k] * w for r, w in zip(rows, rel)], axis=0) for k in rows[0]}
    out["weight"] = np.mean(wts)
    return out, configs, weights


def branch(configs, weights):
    N = configs.configs.shape[0]
    prob = np.cumsum(weights)
    wtot = prob[-1]
    base = np.random.rand() * wtot
    newinds = np.searchsorted(prob, (base + np.linspace(0, wtot, N, endpoint=False)) % wtot)
    configs.configs = configs.configs[newinds]
    weights.fill(wtot / N)
    return configs, weights, newinds
