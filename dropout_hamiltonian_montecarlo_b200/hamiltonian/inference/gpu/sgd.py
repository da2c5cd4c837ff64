"""``sgd`` -- drop-in for reference ``hamiltonian/inference/cpu/sgd.py`` (heavy-ball SGD used by the
notebooks to find a MAP start): per minibatch ``m = gamma m - eps grad; theta += m`` (sgd.py:38-41),
``loss[i] = NLP(theta, last batch)`` per epoch (:42)."""
import numpy as np

from .hmc import _ChainSampler


class sgd(_ChainSampler):
    kind = "sgd"

    def __init__(self, model, start_p, step_size=0.1, **kw):
        super().__init__(model, start_p, step_size=step_size, verbose=kw.pop("verbose", False), **kw)

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
