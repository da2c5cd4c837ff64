"""``sgld`` -- drop-in for reference ``hamiltonian/inference/cpu/sgld.py``.

``step`` (sgld.py:31-39): ``eta ~ N(0, std = 2 eps)`` per variable (:41-46 -- 2 eps is passed as the
*standard deviation*), ``p = eta - eps/2 * grad``, ``q += p``; the incoming momentum is ignored.
(The CuPy variant ``inference/gpu/sgld.py:11-20`` uses a different, multiplicative-noise update;
the NumPy one is the parity target)  The update, the Philox noise and the
momentum store are one fused kernel (csrc/update.cu:k_sgld).
"""
import torch

from .sgmcmc import sgmcmc


class sgld(sgmcmc):
    kind = "sgld"

    def step2(self, state, momentum, rng, **args):
        """sgld.py:15-29 is dead code in the reference: it reads the undefined names ``n_batch`` and ``norm`` and raises
        NameError when called; ``sample`` only ever calls ``step``.  Kept so that the attribute exists, with the same
        outcome (an exception) and a message that says why."""
        raise NameError("sgld.step2 is not runnable in the reference either (undefined n_batch / norm, sgld.py:27); "
                        "use sgld.step")

    def step(self, state, momentum, rng, **args):
        """One SGLD update on the batch passed as X_train / y_train -> (q, p)."""
        saved = self.start
        self.start = state
        try:
            h, shapes, squeeze, like, q0, s = self._setup(**args)
        finally:
            self.start = saved
        s.set_q(q0)
        z = torch.as_tensor(self._noise_tape(rng, 1, s.C, h, shapes)) if rng is not None else None
        # one epoch of one batch == one step at the current step size
        s.sg_run(0, 1, h.N, self.step_size, n_rows=h.N, step0=self._steps_done, z=z, keep_samples=False)
        self._steps_done += 1
        q = self.model.unflatten(s.get(0), shapes, squeeze, like)
        p = self.model.unflatten(s.get(1), shapes, squeeze, like)
        return q, p

    def draw_momentum(self, rng, epsilon):
        """sgld.py:41-46 -- ``N(0, std = 2 eps)`` per variable: 2 eps is passed as the standard deviation."""
        import numpy as np
        return {v: rng.normal(0, 2.0 * epsilon, size=np.asarray(self.start[v]).shape) for v in self.start}
