"""``logistic(hyper)`` -- drop-in for reference ``hamiltonian/models/cpu/logistic.py``,
evaluated for a batch of chains by libbhmc.so.

``par`` is the reference's ``{'weights': [D,1], 'bias': [1]}`` (a leading chain axis batches chains);
``y_train`` is the 0/1 label vector.  The library evaluates the model on the softmax kernels as a two-class
softmax whose class-0 parameters are pinned to zero (``bhmc_logistic_create``), so all three precisions, every
sampler and both multi-GPU shardings apply unchanged.  Unlike the CPU softmax, the logistic ``log_prior``
carries the quadratic term (logistic.py:15-21), so the Metropolis energy is a proper posterior energy.
"""
import numpy as np
import torch

from .softmax import softmax
from ....runtime import LogisticHandle, SoftmaxHandle
from ...._lib import PREC


class logistic(softmax):
    var_names = ("weights", "bias")

    def __init__(self, _hyper, **kw):
        kw.pop("prior", None)
        super().__init__(_hyper, prior="cpu", **kw)

    @staticmethod
    def _labels_of(y, n_classes=None):
        if isinstance(y, torch.Tensor):
            y = y.detach().cpu().numpy()
        y = np.asarray(y).reshape(-1)
        if not np.all((y == 0) | (y == 1)):
            raise ValueError("logistic: y_train must be a 0/1 vector")
        return y.astype(np.int32), 2

    def _new_handle(self, ctx, n, d, k):
        return LogisticHandle(ctx, n, d, self._alpha_local())

    def var_shapes(self, handle):
        return {"weights": (handle.D, 1), "bias": (1,)}

    def log_prior(self, par, **args):
        """logistic.py:15-21: sum_v dim_v/2 log(alpha/2pi) - alpha/2 |theta_v|^2 (per chain when batched)."""
        a = float(self.hyper["alpha"])
        k = 0.0
        for v in par:
            x = np.asarray(par[v].detach().cpu() if isinstance(par[v], torch.Tensor) else par[v], dtype=np.float64)
            base = 1 if v == "bias" else x.shape[-2] * x.shape[-1]
            lead = x.reshape(-1, base) if x.size != base else x.reshape(1, base)
            k = k + base * 0.5 * np.log(a / (2 * np.pi)) - 0.5 * a * np.sum(np.square(lead), axis=1)
        k = np.asarray(k)
        return float(k[0]) if k.size == 1 else k

    def sigmoid(self, y_linear):
        """logistic.py:53-55 (host arrays)."""
        return 1.0 / (1.0 + np.exp(-np.asarray(y_linear)))

    # ---- prediction (logistic.py:42-49,75-87) -----------------------------------------------------
    def _two_class(self, par):
        w = np.asarray(par["weights"].detach().cpu() if isinstance(par["weights"], torch.Tensor) else par["weights"],
                       dtype=np.float32)
        b = np.asarray(par["bias"].detach().cpu() if isinstance(par["bias"], torch.Tensor) else par["bias"],
                       dtype=np.float32)
        squeeze = w.ndim == 2
        w = w.reshape((-1,) + w.shape[-2:])
        b = b.reshape(w.shape[0], 1)
        w2 = np.concatenate([np.zeros_like(w), w], axis=2)          # [C, D, 2]
        b2 = np.concatenate([np.zeros_like(b), b], axis=1)          # [C, 2]
        return w2, b2, squeeze

    def net(self, par, **args):
        """logistic.py:42-49 -> sigmoid(clip(X w + b)), shape [N, 1]."""
        X = args["X_train"]
        w2, b2, squeeze = self._two_class(par)
        Xd = X if (isinstance(X, torch.Tensor) and X.is_cuda) else torch.as_tensor(
            np.ascontiguousarray(np.asarray(X, dtype=np.float32))).to(self.ctx.device)
        Xd = Xd.to(torch.float32).contiguous()
        d = w2.shape[1]
        h = SoftmaxHandle(self.ctx, 1, d, 2, float(self.hyper["alpha"]))
        q = np.concatenate([w2.reshape(w2.shape[0], -1), b2], axis=1)
        probs, _ = h.predict(h.pack(q), Xd, want_probs=True, want_labels=False)
        out = probs[:, :, 1:2].cpu().numpy().astype(np.float64)
        h.close()
        return out[0] if squeeze else out

    def predict(self, par, X, prob=False, batchsize=32):
        """logistic.py:75-87: rows are consumed in whole batches of ``batchsize`` (the remainder is dropped,
        as in the reference) and the result is flattened."""
        n = (X.shape[0] // batchsize) * batchsize if batchsize else X.shape[0]
        yhat = self.net(par, X_train=X[:n])
        out = yhat if prob else (yhat > 0.5).astype(int)
        return out.reshape(out.shape[0], -1) if out.ndim == 3 else out.flatten()

    def predict_stochastic(self, par, X, prob=False, p=0.5, batchsize=32):
        Xn = np.asarray(X.detach().cpu() if isinstance(X, torch.Tensor) else X)
        Z = np.random.binomial(1, p, size=Xn.shape)
        return self.predict(par, np.multiply(Xn, Z), prob, batchsize)
