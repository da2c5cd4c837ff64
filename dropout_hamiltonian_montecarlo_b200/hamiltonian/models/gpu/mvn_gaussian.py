"""``mvn_gaussian(hyper)`` -- drop-in for reference ``hamiltonian/models/cpu/mvn_gaussian.py``
(BASELINE config 1): one GPU thread per chain evaluates ``(x-mu) Sigma^-1`` and the energy."""
import numpy as np

from ... import _base
from ....runtime import MvnHandle


class mvn_gaussian(_base.ChainModel):
    var_names = ("x",)

    def __init__(self, _hyper, *, device=None):
        super().__init__(device)
        self.hyper = _hyper
        self.precision = "fp32"
        self._handle = None

    def handle_for(self, **args):
        if self._handle is None:
            self._handle = MvnHandle(self.ctx, self.hyper["mu"], self.hyper["cov"])
        return self._handle

    def var_shapes(self, handle):
        return {"x": (handle.P,)}

    def grad(self, par, **args):
        """mvn_gaussian.py:14-20."""
        h = self.handle_for()
        q, squeeze, like = self.flatten(par, self.var_shapes(h))
        g, _ = h.grad(h.pack(q), 0, 0, 0, True)
        return self.unflatten(g[:, : h.P], self.var_shapes(h), squeeze, like)

    def negative_log_posterior(self, par, **args):
        """mvn_gaussian.py:22-31."""
        h = self.handle_for()
        q, squeeze, _ = self.flatten(par, self.var_shapes(h))
        out = h.nlp(h.pack(q), 0, 0, 0).cpu().numpy()
        return float(out[0]) if squeeze else out
