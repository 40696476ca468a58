"""TEST INFRASTRUCTURE (oracle) -- product of wave-function factors.

Restates ``MultiplyWF`` (``pyqmc/wf/multiplywf.py:71-132``): log-values and log-gradients
add, ratios multiply, and the Laplacian of the product picks up the cross terms
2 sum_{i<j} grad_i . grad_j (multiplywf.py:121-129).
"""
import numpy as np


class ProductOracle:
    def __init__(self, *factors):
        self.wf_factors = list(factors)
        self.dtype = complex if any(f.dtype == complex for f in factors) else float  # multiplywf.py:79-80

    @property
    def parameters(self):
        out = {}
        for i, f in enumerate(self.wf_factors):
            for k, v in f.parameters.items():
                out[f"wf{i + 1}{k}"] = v
        return out

    def recompute(self, configs):
        sign, val = 1.0, 0.0
        for f in self.wf_factors:
            s, v = f.recompute(configs)
            sign, val = sign * s, val + v
        return sign, val

    def value(self):
        sign, val = 1.0, 0.0
        for f in self.wf_factors:
            s, v = f.value()
            sign, val = sign * s, val + v
        return sign, val

    def updateinternals(self, e, epos, configs, mask=None, saved_values=None):
        saved = [None] * len(self.wf_factors) if saved_values is None else saved_values
        for f, sv in zip(self.wf_factors, saved):
            f.updateinternals(e, epos, configs, mask=mask, saved_values=sv)

    def gradient(self, e, epos):
        return sum(f.gradient(e, epos) for f in self.wf_factors)

    def gradient_value(self, e, epos):
        g, v, saved = 0.0, 1.0, []
        for f in self.wf_factors:
            gi, vi, si = f.gradient_value(e, epos)
            g, v = g + gi, v * vi
            saved.append(si)
        return g, v, tuple(saved)

    def gradient_laplacian(self, e, epos):
        parts = [f.gradient_laplacian(e, epos) for f in self.wf_factors]
        grads = [p[0] for p in parts]
        lap = sum(p[1] for p in parts)
        for i in range(len(grads)):
            for j in range(i + 1, len(grads)):
                lap = lap + 2.0 * np.sum(grads[i] * grads[j], axis=0)
        return sum(grads), lap

    def testvalue(self, e, epos, mask=None):
        v, saved = 1.0, []
        for f in self.wf_factors:
            vi, si = f.testvalue(e, epos, mask=mask)
            v = v * vi
            saved.append(si)
        return v, tuple(saved)

    def testvalue_many(self, e, epos, mask=None):
        v = 1.0
        for f in self.wf_factors:
            v = v * f.testvalue_many(e, epos, mask=mask)
        return v

    def pgradient(self):
        out = {}
        for i, f in enumerate(self.wf_factors):
            for k, v in f.pgradient().items():
                out[f"wf{i + 1}{k}"] = v
        return out
