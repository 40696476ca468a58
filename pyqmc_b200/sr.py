"""Parameter-gradient accumulators for wave-function optimisation with the reference's interface.

``LinearTransform`` (``pyqmc/observables/accumulators.py:113-207``, real parameters) and
``StochasticReconfiguration`` (``pyqmc/observables/stochastic_reconfiguration.py:48-178``; alias
``PGradTransform``): ``__call__`` returns the per-walker ``dpH, dppsi, dpidpj`` on top of the energy
dictionary, ``avg`` their (weighted) means, ``delta_p`` the SR step.  For device-resident wave
functions ``avg`` runs entirely on the device (``qmcb_sr_avg``): local energy, parameter gradients,
nodal regularisation, the weighted column sums and the P x P product ``dp^T (w f dp)`` -- the one
genuinely dense GEMM of this path -- so the (N, P) gradient matrix never crosses PCIe.
"""
import numpy as np

from . import _lib
from .accumulators import KEYS, EnergyAccumulator, _device_context

_SOURCES = {"det_coeff": 0, "mo_coeff_alpha": 1, "mo_coeff_beta": 2, "acoeff": 3, "bcoeff": 4, "ccoeff": 5}


class LinearTransform:
    """Linearises a dictionary of (real) wave-function parameters; ``to_opt[k]`` are boolean arrays."""

    def __init__(self, parameters, to_opt=None):
        parameters = {k: np.asarray(v) for k, v in parameters.items()}
        if to_opt is None:
            to_opt = {k: np.ones(p.shape, dtype=bool) for k, p in parameters.items()}
        self.to_opt = {k: o for k, o in to_opt.items() if np.any(o)}
        self.shapes = {k: parameters[k].shape for k in self.to_opt}
        self.slices = {k: int(np.prod(s)) for k, s in self.shapes.items()}
        self.dtypes = {k: parameters[k].dtype for k in self.to_opt}
        for k, d in self.dtypes.items():
            if d == complex:
                raise NotImplementedError("complex parameters are not supported by the B200 backend")
        self.nparams = int(np.sum([v.sum() for v in self.to_opt.values()])) if self.to_opt else 0

    def serialize_parameters(self, parameters):
        if len(self.to_opt) == 0:
            return np.zeros((0))
        return np.concatenate([np.asarray(parameters[k])[opt] for k, opt in self.to_opt.items()]).real

    def serialize_gradients(self, pgrad):
        grads = [np.asarray(pgrad[k]).reshape(pgrad[k].shape[0], -1)[:, opt.ravel()] for k, opt in self.to_opt.items()]
        if len(grads) == 0:
            return np.zeros((0))
        return np.concatenate(grads, axis=1)

    def deserialize(self, wf, parameters):
        n, d = 0, {}
        for k, opt in self.to_opt.items():
            opt_ = opt.flatten()
            n_p = int(np.sum(opt_))
            flat = np.array(wf.parameters[k], dtype=self.dtypes[k]).reshape(-1)
            flat[opt_] = np.real(parameters[n : n + n_p])
            d[k] = flat.reshape(self.shapes[k])
            n += n_p
        return d

    def device_layout(self):
        """(source id, flat offset) of every serialised parameter, in serialisation order."""
        src, off = [], []
        for k, opt in self.to_opt.items():
            name = next((n for n in _SOURCES if k.endswith(n)), None)
            if name is None:
                raise KeyError(f"no device gradient for parameter {k}")
            idx = np.nonzero(opt.ravel())[0]
            src.extend([_SOURCES[name]] * len(idx))
            off.extend(idx)
        return np.asarray(src, dtype=np.int32), np.asarray(off, dtype=np.int64)


def nodal_regularization(grad2, nodal_cutoff=1e-3):
    """stochastic_reconfiguration.py:20-46."""
    r = 1.0 / grad2
    mask = r < nodal_cutoff**2
    c = 7.0 / (nodal_cutoff**6)
    b = -15.0 / (nodal_cutoff**4)
    a = 9.0 / (nodal_cutoff**2)
    f = a * r + b * r**2 + c * r**3
    f[np.logical_not(mask)] = 1.0
    return mask, f


class StochasticReconfiguration:
    def __init__(self, enacc, transform, nodal_cutoff=1e-3, eps=1e-1, inverse_strategy="pseudo_inverse", verbose=False):
        self.enacc = enacc
        self.transform = transform
        self.nodal_cutoff = nodal_cutoff
        self.eps = eps
        self.inverse_strategy = inverse_strategy
        self.verbose = verbose

    def __call__(self, configs, wf):
        pgrad = wf.pgradient()
        d = self.enacc(configs, wf)
        energy = d["total"]
        dp = self.transform.serialize_gradients(pgrad)
        node_cut, f = nodal_regularization(d["grad2"], self.nodal_cutoff)
        dp_regularized = dp * f[:, np.newaxis]
        d["dpH"] = np.einsum("i,ij->ij", energy, dp_regularized)
        d["dppsi"] = dp_regularized
        d["dpidpj"] = np.einsum("ij,ik->ijk", dp, dp_regularized)
        return d

    def _device(self, wf):
        if not isinstance(self.enacc, EnergyAccumulator):
            return None
        try:
            ctx = _device_context(wf)
        except TypeError:
            return None
        if ctx is None:
            return None
        return ctx

    def avg(self, configs, wf, weights=None):
        nconf = configs.configs.shape[0]
        if weights is None:
            weights = np.ones(nconf)
        weights = weights / np.sum(weights)
        ctx = self._device(wf)
        if ctx is not None and self.transform.nparams > 0:
            return self._avg_device(ctx, configs, wf, weights)
        pgrad = wf.pgradient()
        den = self.enacc(configs, wf)
        energy = den["total"]
        dp = self.transform.serialize_gradients(pgrad)
        node_cut, f = nodal_regularization(den["grad2"])  # the reference's avg uses the default cutoff (line 102)
        dp_regularized = dp * f[:, np.newaxis]
        d = {k: np.average(it, weights=weights, axis=0) for k, it in den.items()}
        if self.transform.nparams > 0:
            d["dpH"] = np.einsum("i,ij->j", energy, weights[:, np.newaxis] * dp_regularized)
            d["dppsi"] = np.average(dp_regularized, weights=weights, axis=0)
            d["dpidpj"] = np.einsum("ij,ik->jk", dp, weights[:, np.newaxis] * dp_regularized, optimize=True)
        return d

    def _avg_device(self, ctx, configs, wf, weights):
        self.enacc._attach(wf)
        nconf, nelec = configs.configs.shape[:2]
        u, rot = self.enacc.draw_ecp_variates(nconf, nelec)
        src, off = self.transform.device_layout()
        P = len(src)
        en, dpH, dppsi, dpidpj = np.empty(6), np.empty(P), np.empty(P), np.empty((P, P))
        w = np.ascontiguousarray(weights, dtype=np.float64)
        _lib.check(ctx.lib.qmcb_sr_avg(ctx.h, P, _lib.iptr(src), off.ctypes.data_as(_lib.c_i64_p), _lib.dptr(w),
                                       _lib.dptr(u), _lib.dptr(rot), 1e-3, _lib.dptr(en), _lib.dptr(dpH),
                                       _lib.dptr(dppsi), _lib.dptr(dpidpj)))
        d = {k: en[i] for i, k in enumerate(KEYS)}
        d["dpH"], d["dppsi"], d["dpidpj"] = dpH, dppsi, dpidpj
        return d

    def keys(self):
        return self.enacc.keys().union(["dpH", "dppsi", "dpidpj"])

    def shapes(self):
        nparms = int(np.sum([np.sum(opt) for opt in self.transform.to_opt.values()]))
        d = {"dpH": (nparms,), "dppsi": (nparms,), "dpidpj": (nparms, nparms)}
        d.update(self.enacc.shapes())
        return d

    def update_state(self, hdf_file):
        pass

    def delta_p(self, steps, data, verbose=False):
        """SR step (stochastic_reconfiguration.py:138-178)."""
        pgrad = 2 * np.real(data["dpH"] - data["total"] * data["dppsi"])
        Sij = np.real(data["dpidpj"] - np.einsum("i,j->ij", data["dppsi"], data["dppsi"]))
        if self.inverse_strategy == "pseudo_inverse":
            invSij = np.linalg.pinv(Sij, rcond=self.eps)
        elif self.inverse_strategy == "regularized_inverse":
            invSij = np.linalg.inv(Sij + self.eps * np.eye(Sij.shape[0]))
        else:
            raise ValueError("Invalid inverse strategy. Valid options are pseudo_inverse and regularized_inverse.")
        v = np.einsum("ij,j->i", invSij, pgrad)
        dp = [-step * v for step in steps]
        report = {"pgrad": np.linalg.norm(pgrad), "SRdot": np.dot(pgrad, v) / (np.linalg.norm(v) * np.linalg.norm(pgrad))}
        return dp, report


PGradTransform = StochasticReconfiguration


def gradient_generator(mol, wf, to_opt=None, nodal_cutoff=1e-3, eps=1e-3, inverse_strategy="regularized_inverse", **ewald_kwargs):
    """accumulators.py:27-42."""
    return StochasticReconfiguration(EnergyAccumulator(mol, **ewald_kwargs), LinearTransform(wf.parameters, to_opt),
                                     nodal_cutoff=nodal_cutoff, eps=eps, inverse_strategy=inverse_strategy)
