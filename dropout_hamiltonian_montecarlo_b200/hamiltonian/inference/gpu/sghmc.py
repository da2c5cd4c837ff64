"""``sghmc`` -- drop-in for reference ``hamiltonian/inference/cpu/sghmc.py``.

The reference class is not runnable through ``sample`` as shipped (its base lacks
``draw_momentum``/``accept`` and ``step`` returns a 3-tuple); the arithmetic of
``step`` (sghmc.py:19-39) is the specification: N(0,1) momentum, L = ceil(2 u path/eps), per
iteration and variable ``r ~ N(0, std 2 eps)``, ``q += eps p``, ``g = grad(q)``,
``p = (1-eps) p + eps g + r``, then hmc's Metropolis test without momentum flip.
``sign='reference'`` keeps the literal ``+eps*grad``; ``sign='descent'`` uses ``-eps*grad``
(Chen et al. 2014).  ``sample`` runs ``epochs`` x minibatches of such steps (one stored sample
per epoch, like sgmcmc.sample).
"""
import numpy as np
import torch

from .hmc import _ChainSampler


class sghmc(_ChainSampler):
    kind = "sghmc"

    def _draws(self, rng, h, shapes, C):
        """Reference consumption order for one step of one chain (sghmc.py:21,25,31,36)."""
        assert C == 1, "host-RNG injection follows the reference's single-chain draw order"
        names = list(self.model.var_names)
        layout = dict(zip(names, zip(h.var_off, h.var_len)))
        z = np.zeros((1, 1, h.P), np.float32)
        for v in self._names():
            o, l = layout[v]
            z[0, 0, o:o + l] = np.asarray(rng.normal(0, 1, size=shapes[v])).reshape(-1)
        u1 = np.random.rand()
        L = int(np.ceil(2 * u1 * self.path_length / self.step_size))
        iters = max(L - 1, 0)
        zn = np.zeros((1, max(iters, 1), 1, h.P), np.float32)
        for it in range(iters):
            for v in self._names():
                o, l = layout[v]
                zn[0, it, 0, o:o + l] = np.asarray(rng.normal(0, 1, size=int(np.prod(shapes[v])))).reshape(-1)
        u2 = np.random.rand()
        return z, np.array([[u1]]), np.array([[u2]]), zn

    def lr_schedule(self, initial_step_size, step, decay_factor, num_batches):
        """sgmcmc.py:88-89 (the reference's sghmc inherits it from sgmcmc)."""
        return initial_step_size * (1.0 / (1.0 + step * decay_factor * num_batches))

    def iterate_minibatches(self, X, y, batchsize):
        """sgmcmc.py:34-38 (inherited from sgmcmc in the reference)."""
        assert X.shape[0] == y.shape[0]
        for start_idx in range(0, X.shape[0] - batchsize + 1, batchsize):
            excerpt = slice(start_idx, start_idx + batchsize)
            yield X[excerpt], y[excerpt]

    def step(self, state, momentum, rng, **args):
        """sghmc.py:19-39 -> (q, p, acceptprob)."""
        saved = self.start
        self.start = state
        try:
            h, shapes, squeeze, like, q0, s = self._setup(**args)
        finally:
            self.start = saved
        s.set_q(q0)
        kw = {}
        if rng is not None:
            z, u1, u2, zn = self._draws(rng, h, shapes, s.C)
            kw = dict(z_momentum=torch.as_tensor(z), u_path=u1, u_accept=u2, z_noise=torch.as_tensor(zn))
        row0, nrows = args.get("_row0", 0), args.get("_nrows", 0)
        out = s.hmc_run(1, self.step_size, self.path_length, row0=row0, nrows=nrows, step0=self._steps_done,
                        keep_samples=False, **kw)
        self._steps_done += 1
        self.last_run = dict(n_grad_evals=out["n_grad_evals"], n_chains=s.C)
        q = self.model.unflatten(s.get(0), shapes, squeeze, like)
        p = self.model.unflatten(s.get(1), shapes, squeeze, like)
        a = out["accept_prob"].cpu().numpy()[0]
        return q, p, (float(a[0]) if squeeze else a)

    def sample(self, epochs=1, burnin=1, batch_size=1, rng=None, **args):
        """Minibatch SGHMC with the sgmcmc.sample epoch structure (sgmcmc.py:40-86): every minibatch is one ``step``
        (sghmc.py:19-39) on that row window.  ``rng=None``: in-kernel Philox draws.  With an ``rng`` the draws of every
        step are taken on the host in the reference's consumption order (``_draws``: momentum, path-length uniform,
        per-iteration noise, accept uniform) and injected; like ``step`` that is a single-chain mode."""
        epochs, burnin, batch_size = int(epochs), int(burnin), int(batch_size)
        h, shapes, squeeze, like, q0, s = self._setup(**args)
        s.set_q(q0)
        nb = (h.N - batch_size) // batch_size + 1
        n_grad = 0
        samples, logp = [], []
        for e in range(burnin + epochs):
            for j in range(nb):
                kw = {}
                if rng is not None:
                    z, u1, u2, zn = self._draws(rng, h, shapes, s.C)
                    kw = dict(z_momentum=torch.as_tensor(z), u_path=u1, u_accept=u2, z_noise=torch.as_tensor(zn))
                o = s.hmc_run(1, self.step_size, self.path_length, row0=j * batch_size, nrows=batch_size,
                              step0=self._steps_done, keep_samples=False, keep_stats=(j == nb - 1), **kw)
                self._steps_done += 1
                n_grad += o["n_grad_evals"]
            if e >= burnin:
                samples.append(s.get(0))
                logp.append(o["loss"].cpu().numpy()[0])
        posterior = self.model.unflatten(np.stack(samples) if samples else np.zeros((0, s.C, h.P), np.float32),
                                         shapes, squeeze, like)
        logp = np.array(logp)
        if squeeze and logp.size:
            logp = logp[:, 0]
        self.last_run = dict(n_grad_evals=n_grad, n_chains=s.C)
        return posterior, logp
