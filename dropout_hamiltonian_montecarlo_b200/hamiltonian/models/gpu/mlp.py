"""``mlp(hyper, n_in, n_mid_units, n_out)`` -- drop-in for reference ``hamiltonian/models/gpu/mlp.py``
(Chainer): 3 ``Linear`` layers with dropout 0.1 on the two hidden pre-activations and on the last hidden
output (``mlp.py:19-31``), mean softmax cross-entropy with integer labels (``:52,57``),
``grad = d loss/d theta + alpha/2 theta`` (``:63``), ``log_likelihood`` = the loss (``:66-78``),
``NLP = loss + log_prior`` (``:80-82``).  ``par`` keys are Chainer's: '/l1/W' (mid,in), '/l1/b', '/l2/W',
'/l2/b', '/l3/W' (out,mid), '/l3/b'; a leading chain axis evaluates many chains at once.
Dropout masks come from the in-kernel Philox generator (fresh per evaluation, like Chainer) or are injected
with ``set_masks`` for parity tests.  As in the reference, dropout stays ON in ``predict``.
"""
import numpy as np
import torch

from ... import _base
from ...._lib import PREC
from ....runtime import MlpHandle

KEYS = ("/l1/W", "/l1/b", "/l2/W", "/l2/b", "/l3/W", "/l3/b")


class mlp(_base.ChainModel):
    var_names = KEYS

    def __init__(self, _hyper, n_in, n_mid_units, n_out, *, dropout=0.1, seed=0, chain_id0=0, device=None,
                 precision="bf16x3"):
        super().__init__(device)
        self.hyper = _hyper
        self.n_in, self.n_mid, self.n_out = int(n_in), int(n_mid_units), int(n_out)
        self.dropout = dropout
        self.seed = seed
        self.chain_id0 = chain_id0
        self.precision = precision  # "bf16x3" / "bf16": large GEMMs on tcgen05; "fp32": CUDA cores
        self._bound = None

    def var_shapes(self, handle=None):
        i, m, o = self.n_in, self.n_mid, self.n_out
        return {"/l1/W": (m, i), "/l1/b": (m,), "/l2/W": (m, m), "/l2/b": (m,), "/l3/W": (o, m), "/l3/b": (o,)}

    def bind(self, X, y):
        key = (_base.array_key(X), _base.array_key(y))
        if self._bound is not None and self._bound[0] == key:
            return self._bound[1]
        ctx = self.ctx
        Xt = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.asarray(X))
        Xt = Xt.to(torch.float32).contiguous()
        yt = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))
        yt = yt.to(torch.int32).contiguous()
        if Xt.is_cuda:
            yt = yt.to(ctx.device)
        h = MlpHandle(ctx, Xt.shape[0], self.n_in, self.n_mid, self.n_out, float(self.hyper["alpha"]), self.dropout,
                      self.seed, self.chain_id0)
        h.bind(Xt, yt)
        if self._bound is not None:
            self._bound[1].close()
        self._bound = (key, h, (X, y))
        return h

    def handle_for(self, **args):
        return self.bind(args["X_train"], args["y_train"])

    def set_masks(self, masks):
        """masks: [3, C, B, n_mid] keep flags (array / tensor) or None; applies to the bound data."""
        h = self._bound[1]
        if masks is None:
            h.set_masks(None)
        else:
            h.set_masks(torch.as_tensor(np.asarray(masks)).to(torch.uint8).contiguous().to(self.ctx.device))

    def grad(self, par, **args):
        """mlp.py:47-64."""
        h = self.handle_for(**args)
        q, squeeze, like = self.flatten(par, self.var_shapes())
        g, _ = h.grad(h.pack(q), 0, h.N, PREC[self.precision], True)
        return self.unflatten(g[:, : h.P], self.var_shapes(), squeeze, like)

    def log_likelihood(self, par, **args):
        """mlp.py:66-78 -- returns the (positive) mean cross-entropy, as the reference does."""
        h = self.handle_for(**args)
        q, squeeze, _ = self.flatten(par, self.var_shapes())
        _, ll = h.grad(h.pack(q), 0, h.N, PREC[self.precision], False)
        out = ll.cpu().numpy()
        return float(out[0]) if squeeze else out

    def log_prior(self, par, **args):
        """mlp.py:40-45."""
        return -sum(0.5 * self.hyper["alpha"] * float(np.sum(np.square(_base.to_numpy(par[v])))) / _base.to_numpy(par[v]).size
                    for v in par)

    def negative_log_posterior(self, par, **args):
        """mlp.py:80-82."""
        h = self.handle_for(**args)
        q, squeeze, _ = self.flatten(par, self.var_shapes())
        out = h.nlp(h.pack(q), 0, h.N, PREC[self.precision]).cpu().numpy()
        return float(out[0]) if squeeze else out

    def predict(self, par, X_test, prob=False):
        """mlp.py:84-95 -- forward pass with dropout ON (the reference never passes ``train=False``), then softmax
        (``prob=True``) or argmax.  Masks are drawn by the in-kernel Philox generator (injected training masks are
        shaped for the training batch and are not used here); ``dropout=0`` makes the pass deterministic."""
        Xd = X_test if (isinstance(X_test, torch.Tensor) and X_test.is_cuda) else torch.as_tensor(
            np.ascontiguousarray(np.asarray(X_test, dtype=np.float32))).to(self.ctx.device)
        Xd = Xd.to(torch.float32).contiguous()
        h = self._bound[1] if self._bound is not None else None
        own = h is None
        if own:  # no training data bound: a handle just for the forward pass
            h = MlpHandle(self.ctx, Xd.shape[0], self.n_in, self.n_mid, self.n_out, float(self.hyper["alpha"]), self.dropout,
                          self.seed, self.chain_id0)
        q, squeeze, _ = self.flatten(par, self.var_shapes())
        probs, labels = h.predict(h.pack(q), Xd, PREC[self.precision], want_probs=prob, want_labels=not prob)
        out = (probs if prob else labels).cpu().numpy()
        if own:
            h.close()
        return out[0] if squeeze else out
