"""Parameter-gradient (stochastic reconfiguration) accumulator on the device.

Interface of the reference's ``StochasticReconfiguration`` (``pyqmc/observables/stochastic_reconfiguration.py
:48-178``; alias ``PGradTransform``) and of its parameter serialiser ``LinearTransform``
(``pyqmc/observables/accumulators.py:113-207``, real parameters): ``__call__`` returns the per-walker
``dpH, dppsi, dpidpj`` on top of the energy dictionary, ``avg`` their weighted means, ``delta_p`` the SR
step, ``transform.serialize_parameters / deserialize`` map between the parameter dictionary and the flat
optimisation vector -- which is all ``linemin`` (``linemin.py:102-270``) needs.

``avg`` -- the call inside the optimisation loop -- runs entirely on the device (``qmcb_sr_avg``): local
energy, parameter gradients, nodal regularisation, the weighted column sums and the P x P product
``dp^T (w f dp)`` (the one genuinely dense GEMM of this path), so the (N, P) gradient matrix never crosses
PCIe.  The serialiser therefore also exports the flat (source array, offset) list the device gathers from.
"""
import numpy as np

from . import _lib
from .accumulators import KEYS, EnergyAccumulator, _device_context

# device-side gradient arrays a serialised parameter can come from (csrc/qmcb200.cu: qmcb_sr_avg)
_SOURCES = {"det_coeff": 0, "mo_coeff_alpha": 1, "mo_coeff_beta": 2, "acoeff": 3, "bcoeff": 4, "ccoeff": 5}
DEFAULT_NODAL_CUTOFF = 1e-3


class ParameterMap:
    """Flat view of the optimisable entries of a parameter dictionary.

    ``to_opt[name]`` is a boolean array shaped like ``parameters[name]``; names without a True entry are
    dropped.  The flat vector lists, name after name in ``to_opt`` order, the selected entries in C order
    (the reference's convention, so optimisation vectors are interchangeable)."""

    def __init__(self, parameters, to_opt=None):
        if to_opt is None:
            to_opt = {k: np.ones(np.shape(v), dtype=bool) for k, v in parameters.items()}
        self.to_opt, self.shapes, self.dtypes, self.slices, self._picked = {}, {}, {}, {}, {}
        for name, flags in to_opt.items():
            flags = np.asarray(flags, dtype=bool)
            if not flags.any():
                continue
            value = np.asarray(parameters[name])
            self.to_opt[name] = flags
            self.shapes[name], self.dtypes[name] = value.shape, value.dtype
            self.slices[name] = value.size
            self._picked[name] = np.flatnonzero(flags)
        self.nparams = sum(len(i) for i in self._picked.values())
        # complex parameters contribute their imaginary parts as extra real variables appended after all the real
        # parts (the reference's layout, accumulators.py:121-133, 142, 155)
        self.complex = {k: np.issubdtype(d, np.complexfloating) for k, d in self.dtypes.items()}
        flags = [np.full(len(i), self.complex[k]) for k, i in self._picked.items()]
        self._imag = np.concatenate(flags) if any(self.complex.values()) else np.zeros(0, dtype=bool)

    def serialize_parameters(self, parameters):
        parts = [np.ravel(np.asarray(parameters[k]))[i] for k, i in self._picked.items()]
        if not parts:
            return np.empty(0)
        flat = np.concatenate(parts)
        return np.concatenate((flat.real, flat[self._imag].imag)) if len(self._imag) else np.real(flat)

    def serialize_gradients(self, pgrad):
        """(N, P) matrix of d ln Psi / d p from a ``pgradient()`` dictionary of (N, *shape) arrays; the derivative
        with respect to an imaginary part is i times the one with respect to the real part."""
        parts = [np.asarray(pgrad[k]).reshape(len(pgrad[k]), -1)[:, i] for k, i in self._picked.items()]
        if not parts:
            return np.empty(0)
        dp = np.concatenate(parts, axis=1)
        return np.concatenate((dp, dp[:, self._imag] * 1j), axis=1) if len(self._imag) else dp

    def deserialize(self, wf, vector):
        """Parameter dictionary with the optimisable entries replaced by ``vector`` (others as in ``wf``)."""
        out, start, istart = {}, 0, self.nparams
        for name, picked in self._picked.items():
            full = np.array(wf.parameters[name], dtype=self.dtypes[name]).ravel()
            full[picked] = np.real(vector[start:start + len(picked)])
            if self.complex[name]:
                full[picked] += 1j * np.asarray(vector[istart:istart + len(picked)])
                istart += len(picked)
            out[name] = full.reshape(self.shapes[name])
            start += len(picked)
        return out

    def device_layout(self):
        """(source id, flat offset) of every serialised parameter, in serialisation order."""
        source, offset = [], []
        for name, picked in self._picked.items():
            kind = next((s for s in _SOURCES if name.endswith(s)), None)
            if kind is None:
                raise KeyError(f"no device gradient for parameter {name}")
            source.append(np.full(len(picked), _SOURCES[kind], dtype=np.int32))
            offset.append(picked.astype(np.int64))
        if not source:
            return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64)
        return np.concatenate(source), np.concatenate(offset)


LinearTransform = ParameterMap


def nodal_regularization(grad2, nodal_cutoff=DEFAULT_NODAL_CUTOFF):
    """Damping of the parameter gradients near the node (stochastic_reconfiguration.py:20-46; Pathak &
    Wagner, AIP Advances 10, 085213): with ``u = 1 / (|grad Psi / Psi|^2 cutoff^2)``, walkers with u < 1 get
    the factor ``9u - 15u^2 + 7u^3`` (which is 1 with zero slope at u = 1), the others 1."""
    r = 1.0 / grad2
    near = r < nodal_cutoff**2
    a, b, c = 9.0 / nodal_cutoff**2, -15.0 / nodal_cutoff**4, 7.0 / nodal_cutoff**6
    return near, np.where(near, a * r + b * r**2 + c * r**3, 1.0)


class StochasticReconfiguration:
    """Energy + parameter-derivative accumulator; see the module docstring."""

    def __init__(self, enacc, transform, nodal_cutoff=DEFAULT_NODAL_CUTOFF, eps=1e-1,
                 inverse_strategy="pseudo_inverse", verbose=False):
        self.enacc, self.transform = enacc, transform
        self.nodal_cutoff, self.eps = nodal_cutoff, eps
        self.inverse_strategy, self.verbose = inverse_strategy, verbose

    def _walker_terms(self, configs, wf, cutoff):
        out = self.enacc(configs, wf)
        dp = self.transform.serialize_gradients(wf.pgradient())
        _, damp = nodal_regularization(out["grad2"], cutoff)
        return out, dp, dp * damp[:, None]

    def __call__(self, configs, wf):
        out, dp, dp_reg = self._walker_terms(configs, wf, self.nodal_cutoff)
        out["dpH"] = out["total"][:, None] * dp_reg
        out["dppsi"] = dp_reg
        out["dpidpj"] = dp[:, :, None] * dp_reg[:, None, :]
        return out

    def avg(self, configs, wf, weights=None):
        n = len(configs.configs)
        w = np.ones(n) if weights is None else np.asarray(weights, dtype=float)
        w = w / np.sum(w)
        ctx = self._device(wf)
        if ctx is not None and self.transform.nparams > 0:
            return self._avg_device(ctx, configs, wf, w)
        # the reference's avg regularises with the DEFAULT cutoff whatever nodal_cutoff is (line 102)
        out, dp, dp_reg = self._walker_terms(configs, wf, DEFAULT_NODAL_CUTOFF)
        mean = {k: np.tensordot(w, v, axes=(0, 0)) for k, v in out.items()}
        if self.transform.nparams > 0:
            mean["dpH"] = (w * out["total"]) @ dp_reg
            mean["dppsi"] = w @ dp_reg
            mean["dpidpj"] = dp.T @ (w[:, None] * dp_reg)
        return mean

    def _device(self, wf):
        if not isinstance(self.enacc, EnergyAccumulator):
            return None
        if wf.dtype == complex:  # qmcb_sr_avg reduces real gradients: complex ones are reduced below from pgradient()
            return None
        try:
            return _device_context(wf)
        except TypeError:
            return None

    def _avg_device(self, ctx, configs, wf, weights):
        self.enacc._attach(wf)
        nconf, nelec = configs.configs.shape[:2]
        u, rot = self.enacc.draw_ecp_variates(nconf, nelec)
        src, off = self.transform.device_layout()
        P = len(src)
        en, dpH, dppsi, dpidpj = np.empty(6), np.empty(P), np.empty(P), np.empty((P, P))
        w = np.ascontiguousarray(weights, dtype=np.float64)
        _lib.check(ctx.lib.qmcb_sr_avg(ctx.h, P, _lib.iptr(src), off.ctypes.data_as(_lib.c_i64_p), _lib.dptr(w),
                                       _lib.dptr(u), _lib.dptr(rot), DEFAULT_NODAL_CUTOFF, _lib.dptr(en), _lib.dptr(dpH),
                                       _lib.dptr(dppsi), _lib.dptr(dpidpj)))
        d = {k: en[i] for i, k in enumerate(KEYS)}
        d["dpH"], d["dppsi"], d["dpidpj"] = dpH, dppsi, dpidpj
        return d

    def keys(self):
        return self.enacc.keys().union(["dpH", "dppsi", "dpidpj"])

    def shapes(self):
        P = self.transform.nparams
        return {"dpH": (P,), "dppsi": (P,), "dpidpj": (P, P), **self.enacc.shapes()}

    def update_state(self, hdf_file):
        """Nothing to carry between optimisation iterations (interface of line 129)."""

    def delta_p(self, steps, data, verbose=False):
        """Parameter changes ``-step S^-1 g`` for every step size, with ``g = 2 Re(<E dp> - <E><dp>)`` and the
        overlap matrix ``S = <dp_i dp_j> - <dp_i><dp_j>`` inverted by pseudo-inverse (cut at ``eps``) or after
        adding ``eps`` to the diagonal (stochastic_reconfiguration.py:138-178)."""
        force = 2 * np.real(data["dpH"] - data["total"] * data["dppsi"])
        overlap = np.real(data["dpidpj"] - np.outer(data["dppsi"], data["dppsi"]))
        if self.inverse_strategy == "pseudo_inverse":
            direction = np.linalg.pinv(overlap, rcond=self.eps) @ force
        elif self.inverse_strategy == "regularized_inverse":
            direction = np.linalg.solve(overlap + self.eps * np.eye(len(force)), force)
        else:
            raise ValueError("Invalid inverse strategy. Valid options are pseudo_inverse and regularized_inverse.")
        cosine = force @ direction / (np.linalg.norm(direction) * np.linalg.norm(force))
        return [-s * direction for s in steps], {"pgrad": np.linalg.norm(force), "SRdot": cosine}


PGradTransform = StochasticReconfiguration


def gradient_generator(mol, wf, to_opt=None, nodal_cutoff=DEFAULT_NODAL_CUTOFF, eps=1e-3,
                       inverse_strategy="regularized_inverse", **ewald_kwargs):
    """The accumulator ``recipes.OPTIMIZE`` builds (accumulators.py:27-42), on device objects."""
    return StochasticReconfiguration(EnergyAccumulator(mol, **ewald_kwargs), ParameterMap(wf.parameters, to_opt),
                                     nodal_cutoff=nodal_cutoff, eps=eps, inverse_strategy=inverse_strategy)
