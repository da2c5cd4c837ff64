"""``sgd`` -- drop-in for reference ``hamiltonian/inference/cpu/sgd.py`` (heavy-ball SGD used by the
notebooks to find a MAP start): per minibatch ``m = gamma m - eps grad; theta += m`` (sgd.py:38-41),
``loss[i] = NLP(theta, last batch)`` per epoch (:42).  ``fit_dropout`` (:47-70) multiplies every minibatch by a
fresh Bernoulli(p) input mask (no rescaling) and reports ``-log_likelihood`` of the unmasked last batch."""
import numpy as np
import torch

from .hmc import _ChainSampler


class sgd(_ChainSampler):
    kind = "sgd"

    def __init__(self, model, start_p, step_size=0.1, **kw):
        super().__init__(model, start_p, step_size=step_size, verbose=kw.pop("verbose", False), **kw)

    def iterate_minibatches(self, X, y, batchsize):
        """sgd.py:19-23 -- sequential windows, remainder dropped."""
        assert X.shape[0] == y.shape[0]
        for start_idx in range(0, X.shape[0] - batchsize + 1, batchsize):
            excerpt = slice(start_idx, start_idx + batchsize)
            yield X[excerpt], y[excerpt]

    def fit(self, epochs=1, batch_size=1, gamma=0.9, **args):
        """sgd.py:25-45 -> (par, loss_val)."""
        verbose = args.pop("verbose", None)
        epochs, batch_size = int(epochs), int(batch_size)
        h, shapes, squeeze, like, q0, s = self._setup(**args)
        s.set_q(q0)
        out = s.sg_run(epochs, 0, batch_size, self.step_size, n_rows=h.N, gamma=gamma, keep_samples=False)
        par = self.model.unflatten(s.get(0), shapes, squeeze, like)
        loss = out["logp"].cpu().numpy()
        if squeeze:
            loss = loss[:, 0]
        if verbose:
            for i in range(0, epochs, max(1, epochs // 10)):
                print("loss: {0:.4f}".format(float(np.mean(loss[i]))))
        self.last_run = dict(n_grad_evals=out["n_grad_evals"], n_chains=s.C)
        return par, loss

    def fit_dropout(self, epochs=1, batch_size=1, gamma=0.9, p=0.5, **args):
        """sgd.py:47-70 -> (par, loss_val).  ``p`` is the KEEP probability (``np.random.binomial(1, p)``, :60).
        Masks come from the in-kernel Philox generator; ``rng="numpy"`` instead draws them on the host from the
        global ``np.random.binomial`` in the reference's order (one [batch, D] draw per minibatch) and injects them."""
        verbose = args.pop("verbose", None)
        rng = args.pop("rng", None)
        epochs, batch_size = int(epochs), int(batch_size)
        h, shapes, squeeze, like, q0, s = self._setup(**args)
        s.set_q(q0)
        masks = None
        if rng is not None:
            nb = (h.N - batch_size) // batch_size + 1
            masks = torch.as_tensor(np.stack([np.random.binomial(1, p, size=(batch_size, h.D)).astype(np.uint8)
                                              for _ in range(epochs * nb)]))
        out = s.sg_run(epochs, 0, batch_size, self.step_size, n_rows=h.N, gamma=gamma, keep_samples=False,
                       dropout_keep=p, masks=masks, step0=self._steps_done)
        self._steps_done += epochs * ((h.N - batch_size) // batch_size + 1)
        par = self.model.unflatten(s.get(0), shapes, squeeze, like)
        loss = out["logp"].cpu().numpy()
        if squeeze:
            loss = loss[:, 0]
        if verbose:
            for i in range(0, epochs, max(1, epochs // 10)):
                print("loss: {0:.4f}".format(float(np.mean(loss[i]))))
        self.last_run = dict(n_grad_evals=out["n_grad_evals"], n_chains=s.C)
        return par, loss
