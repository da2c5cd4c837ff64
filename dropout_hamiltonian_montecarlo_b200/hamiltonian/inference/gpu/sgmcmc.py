"""``sgmcmc`` -- drop-in for reference ``hamiltonian/inference/cpu/sgmcmc.py``: the minibatch
driver shared by SGLD (and, in the reference's intent, SGHMC).

Kept semantics (all lines in ``inference/cpu/sgmcmc.py``): sequential unshuffled windows with
the remainder dropped (:34-38); ``burnin`` epochs at the constant initial step size (:55-63);
during sampling the step size is re-assigned *after* batch j to ``eps0/(1+j*eps0)`` (:72-73,
:88-89); one stored sample and ``NLP(q, last batch)`` per epoch (:79-81).  The full data matrix
is bound once; each minibatch is a row window of it, so no data moves per step.
"""
import numpy as np
import torch

from .hmc import _ChainSampler


class sgmcmc(_ChainSampler):
    kind = "sgld"

    def lr_schedule(self, initial_step_size, step, decay_factor, num_batches):
        """sgmcmc.py:88-89."""
        return initial_step_size * (1.0 / (1.0 + step * decay_factor * num_batches))

    def iterate_minibatches(self, X, y, batchsize):
        """sgmcmc.py:34-38."""
        assert X.shape[0] == y.shape[0]
        for start_idx in range(0, X.shape[0] - batchsize + 1, batchsize):
            excerpt = slice(start_idx, start_idx + batchsize)
            yield X[excerpt], y[excerpt]

    def _noise_tape(self, rng, n_batches, C, h, shapes):
        """sgld.py:41-46 consumption order: per step (per chain) rng.normal per variable in start_p order."""
        names = list(self.model.var_names)
        layout = dict(zip(names, zip(h.var_off, h.var_len)))
        z = np.empty((n_batches, C, h.P), dtype=np.float32)
        for t in range(n_batches):
            for c in range(C):
                for v in self._names():
                    o, l = layout[v]
                    z[t, c, o:o + l] = np.asarray(rng.normal(0, 1, size=shapes[v])).reshape(-1)
        return z

    def sample(self, epochs=1, burnin=1, batch_size=1, rng=None, **args):
        """sgmcmc.py:40-86 -> (posterior, logp_samples)."""
        epochs, burnin, batch_size = int(epochs), int(burnin), int(batch_size)
        h, shapes, squeeze, like, q0, s = self._setup(**args)
        s.set_q(q0)
        nb = (h.N - batch_size) // batch_size + 1
        if rng is None:
            out = s.sg_run(epochs, burnin, batch_size, self.step_size, n_rows=h.N, step0=self._steps_done)
            self._steps_done += nb * (epochs + burnin)
        else:
            # Injected noise (parity mode): the tape of ONE epoch at a time ([n_batches, C, P] fp32) instead of the whole
            # run's -- MNIST softmax at 100 epochs would be 1.9 GB per chain.  The library call is cut at epoch
            # boundaries; what carries over is the step size of the next epoch's first batch (sgmcmc.py:72-73: eps is
            # re-assigned AFTER batch j, so batch 0 of epoch e >= 1 runs at lr(n_batches - 1)).
            samples, logps, n_grad = [], [], 0
            first = 0.0  # 0 = start at step_size
            for e in range(burnin + epochs):
                z = torch.as_tensor(self._noise_tape(rng, nb, s.C, h, shapes))
                sampling = e >= burnin
                o = s.sg_run(1 if sampling else 0, 0 if sampling else 1, batch_size, self.step_size, n_rows=h.N,
                             step0=self._steps_done, z=z, first_step_size=first)
                self._steps_done += nb
                n_grad += o["n_grad_evals"]
                if sampling:
                    first = o["final_step_size"]
                    samples.append(o["samples"])
                    logps.append(o["logp"])
            out = {"samples": torch.cat(samples) if samples else s.ctx.empty((0, s.C, s.P)),
                   "logp": torch.cat(logps) if logps else s.ctx.empty((0, s.C), torch.float64),
                   "n_grad_evals": n_grad, "final_step_size": first if epochs > 0 else self.step_size}
        posterior = self.model.unflatten(out["samples"], shapes, squeeze, like)
        logp = out["logp"].cpu().numpy()
        if squeeze:
            logp = logp[:, 0]
        self.step_size = out["final_step_size"]  # the reference leaves self.step_size decayed (:73)
        if self.verbose and epochs > 0:
            for i in range(0, epochs, max(1, epochs // 10)):
                print("loss: {0:.4f}".format(float(np.mean(logp[i]))))
        self.last_run = dict(n_grad_evals=out["n_grad_evals"], n_chains=s.C)
        return posterior, logp
