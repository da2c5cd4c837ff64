"""``hmc`` -- drop-in for reference ``hamiltonian/inference/cpu/hmc.py`` (``gpu/hmc.py`` is its
CuPy twin), running ``n_chains`` chains batched on one B200 with no host round-trip per
leapfrog step.

Reference semantics kept (SURVEY 8(a'), all cited lines in ``inference/cpu/hmc.py``):
  * ``step`` redraws the momentum (:41), draws L = ceil(2 u path/eps) (:46), runs L-1
    Gauss-Seidel sweeps with the (eps/2, eps) kick pattern (:49-54), flips the momentum
    (:58-59) and does the Metropolis test on mean-NLP + kinetic energy (:60-71).
  * ``sample`` draws and discards one momentum (:93), runs ``burnin`` then ``niter`` steps,
    records q and ``loss[i] = NLP(q)`` per step (:108-116).
Random numbers: with ``rng=None`` (default) all draws come from the in-kernel Philox
generator.  If the caller passes an ``rng`` (``RandomState``-like), draws are taken on the
host from ``rng.normal`` / the global ``np.random.rand`` *in the reference's consumption
order* and injected, so a chain reproduces the reference trajectory for the same seeds.
"""
import numpy as np
import torch

from ...._lib import KIND, PREC
from ....runtime import SamplerHandle


class DualAveragingStepSize:
    """hmc.py:141-176 (host scalar logic; the reference only ever calls update() once)."""

    def __init__(self, initial_step_size, target_accept=0.8, gamma=0.05, t0=10.0, kappa=0.75):
        self.mu = np.log(10 * initial_step_size)
        self.target_accept = target_accept
        self.gamma = gamma
        self.t = t0
        self.kappa = kappa
        self.error_sum = 0
        self.log_averaged_step = 0

    def update(self, p_accept):
        self.error_sum += self.target_accept - p_accept
        log_step = self.mu - self.error_sum / (np.sqrt(self.t) * self.gamma)
        eta = self.t ** -self.kappa
        self.log_averaged_step = eta * log_step + (1 - eta) * self.log_averaged_step
        self.t += 1
        return np.exp(log_step), np.exp(self.log_averaged_step)


class _StepTrace:
    """``sample_positions`` / ``sample_momentums`` of hmc.py:108-111,119: one entry per sampling iteration, entry ``i`` =
    ``[dict]`` -- the position (or the momentum drawn by ``step``, :41-44) at the START of iteration ``i``.  The
    reference builds these lists by deep-copying every state; here the entries are materialised on access from what
    the run already holds (the samples; the injected tape or the counter-based generator), so a long run does not
    pay for a second and third copy of the chain unless the caller actually reads them."""

    def __init__(self, n, make):
        self._n, self._make = int(n), make

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("iteration %d out of range [0, %d)" % (i, self._n))
        return [self._make(i)]

    def __iter__(self):
        return (self[i] for i in range(self._n))


TAG_MOMENTUM = 0x01000000  # csrc/philox.cuh: stream of the N(0,1) momentum draws, stream_lo = absolute step index


class _ChainSampler:
    """Shared plumbing of hmc / sgld / sghmc / sgd: start-point handling and handle caching."""

    kind = "hmc"

    def __init__(self, model, start_p, path_length=1.0, step_size=0.1, verbose=True, *, n_chains=None, seed=0,
                 sweep="reference", path_length_mode="per_chain", integrator="reference", precision=None,
                 reject_nan=False, chain_id0=0, sign="reference"):
        self.start = start_p
        self.step_size = step_size
        self.path_length = path_length
        self.model = model
        self.verbose = verbose
        self.n_chains = n_chains
        self.seed = int(seed)
        self.sweep = sweep
        self.path_length_mode = path_length_mode
        self.integrator = integrator
        self.precision = precision or getattr(model, "precision", "bf16x3")
        self.reject_nan = reject_nan
        self.chain_id0 = int(chain_id0)
        self.sign = sign
        self._sampler = None
        self._steps_done = 0
        self.last_run = {}

    # -- helpers -----------------------------------------------------------------------------------
    def _names(self):
        return list(self.start.keys())

    def _setup(self, **args):
        """Bind data, flatten the start point, (re)create the device sampler.  Returns
        (handle, shapes, squeeze, like)."""
        model = self.model
        h = model.handle_for(**args)
        shapes = model.var_shapes(h)
        names = list(model.var_names)
        if set(self._names()) != set(names):
            raise ValueError("start_p keys %s do not match the model variables %s" % (self._names(), names))
        q0, squeeze, like = model.flatten(self.start, shapes)
        C = self.n_chains or q0.shape[0]
        if q0.shape[0] == 1 and C > 1:
            q0 = np.repeat(q0, C, axis=0)
        elif q0.shape[0] != C:
            raise ValueError("start_p has %d chains but n_chains=%d" % (q0.shape[0], C))
        squeeze = squeeze and C == 1
        # Gauss-Seidel groups follow the *dict order* of start_p (hmc.py:50); 'joint' moves everything at once
        layout = dict(zip(names, zip(h.var_off, h.var_len)))
        if self.sweep == "reference":
            groups = [layout[v] for v in self._names()]
        elif self.sweep == "joint":
            groups = [(0, h.P)]
        else:
            raise ValueError("sweep must be 'reference' or 'joint'")
        key = (id(h), C, tuple(groups), self.precision, self.path_length_mode, self.integrator, self.sign,
               self.reject_nan, self.seed, self.chain_id0)
        if self._sampler is None or self._sampler[0] != key:
            if self._sampler is not None:
                self._sampler[1].close()
            s = SamplerHandle(h.ctx, h, KIND[self.kind], C, seed=self.seed, chain_id0=self.chain_id0,
                              precision=PREC[self.precision], sweep=groups,
                              shared_path=self.path_length_mode == "shared", leapfrog=self.integrator == "leapfrog",
                              sghmc_descent=self.sign == "descent", reject_nan=self.reject_nan)
            if getattr(model, "row_sharded", False):
                from ....parallel import RowShardHook, row_comm_for
                comm = row_comm_for(h.ctx, getattr(model, "group", None))
                if comm is not None:  # grouped NCCL all-reduce enqueued by the C driver after every evaluation
                    comm.attach(s)
                    s._row_comm = comm
                else:                 # not an NCCL group (gloo tests) or BHMC_ROW_COMM=hook
                    s._row_hook = RowShardHook(s, getattr(model, "group", None))  # keeps the callback alive
            self._sampler = (key, s)
        return h, shapes, squeeze, like, q0, self._sampler[1]

    def _host_draws(self, rng, n_steps, C, h, shapes, scale=1.0):
        """Reference consumption order (hmc.py:41,46,61): per step, per chain: rng.normal per
        variable in start_p order, then two global uniforms."""
        names = list(self.model.var_names)
        layout = dict(zip(names, zip(h.var_off, h.var_len)))
        z = np.empty((n_steps, C, h.P), dtype=np.float32)
        u1 = np.empty((n_steps, C))
        u2 = np.empty((n_steps, C))
        for t in range(n_steps):
            for c in range(C):
                for v in self._names():
                    o, l = layout[v]
                    z[t, c, o:o + l] = np.asarray(rng.normal(0, 1, size=shapes[v])).reshape(-1)
                u1[t, c] = np.random.rand()
                u2[t, c] = np.random.rand()
        return z, u1, u2


    def _drawn_momentum(self, s, step):
        """The N(0,1) momentum the device drew for absolute step index ``step`` (csrc/update.cu:k_hmc_begin keys it by
        (seed, global chain id, step, TAG_MOMENTUM)): regenerated from the counter-based generator -> [C, P] numpy."""
        import ctypes as C
        from ...._lib import check
        out = torch.empty((s.C, s.P), dtype=torch.float32, device=s.ctx.device)
        check(s.ctx.L.bhmc_philox_normal(s.ctx.handle, C.c_void_p(out.data_ptr()), s.C, s.P, s.P, self.seed,
                                         self.chain_id0, int(step) & 0xffffffff, TAG_MOMENTUM))
        return out.cpu().numpy()


class hmc(_ChainSampler):
    kind = "hmc"

    # ---- the reference's building blocks, for callers that compose their own loop (host side; ``step`` / ``sample``
    # do not go through them: the device drivers fuse these operations) ------------------------------------------
    def potential_energy(self, p):
        """hmc.py:74-79 -- despite its name the kinetic energy ``sum_v 0.5 |p_v|^2``."""
        return float(sum(0.5 * np.sum(np.square(np.asarray(p[v]))) for v in p))

    def draw_momentum(self, rng):
        """hmc.py:82-87."""
        return {v: rng.normal(0, 1, size=np.asarray(self.start[v]).shape) for v in self.start}

    def accept(self, current_q, proposal_q, current_p, proposal_p, **args):
        """hmc.py:67-71 -- ``min(1, exp(E_cur - E_new))`` with the Python builtin ``min`` (NaN -> 1)."""
        e_new = self.model.negative_log_posterior(proposal_q, **args) + self.potential_energy(proposal_p)
        e_cur = self.model.negative_log_posterior(current_q, **args) + self.potential_energy(current_p)
        return min(1, np.exp(e_cur - e_new))

    def step(self, state, momentum, rng, **args):
        """hmc.py:39-64 -> (q, p, positions, momentums, acceptprob); ``momentum`` is ignored, as in
        the reference."""
        saved = self.start
        self.start = state
        try:
            h, shapes, squeeze, like, q0, s = self._setup(**args)
        finally:
            self.start = saved
        s.set_q(q0)
        kw = {}
        if rng is not None:
            z, u1, u2 = self._host_draws(rng, 1, s.C, h, shapes)
            kw = dict(z_momentum=torch.as_tensor(z), u_path=u1, u_accept=u2)
        step_index = self._steps_done
        out = s.hmc_run(1, self.step_size, self.path_length, step0=step_index, keep_samples=False, **kw)
        self._steps_done += 1
        q = self.model.unflatten(s.get(0), shapes, squeeze, like)
        p = self.model.unflatten(s.get(1), shapes, squeeze, like)
        a = out["accept_prob"].cpu().numpy()[0]
        a = float(a[0]) if squeeze else a
        # hmc.py:44: positions, momentums = [deepcopy(q)], [deepcopy(p)] -- copies of the start point and of the
        # momentum drawn at :41 (not of the caller's objects)
        p0 = z[0] if rng is not None else self._drawn_momentum(s, step_index)
        positions = [self.model.unflatten(q0.copy(), shapes, squeeze, like)]
        momentums = [self.model.unflatten(p0, shapes, squeeze, like)]
        return q, p, positions, momentums, a

    def find_reasonable_epsilon(self, p_accept, **args):
        """hmc.py:122-130: the dual-averaging update of ``DualAveragingStepSize.update`` written against attributes
        (``self.t``, ``self.mu`` ...) that the reference's ``hmc`` never sets, so calling it there raises
        AttributeError.  Here the same recurrence runs on a ``DualAveragingStepSize`` created on first use from the
        current step size -> (noisy step size, averaged step size)."""
        if getattr(self, "_dual_avg", None) is None:
            self._dual_avg = DualAveragingStepSize(self.step_size)
        return self._dual_avg.update(p_accept)

    def backend_mean(self, multi_backend, niter, ncores=None):
        """hmc.py:132-138 over the files written by ``sample(backend=...)``."""
        from ...sink import backend_mean
        return backend_mean(multi_backend, niter, self.start)

    def sample(self, niter=1e4, burnin=1e3, rng=None, **args):
        """hmc.py:90-119 -> (posterior, loss, sample_positions, sample_momentums).

        Extensions (keyword-only, consumed before the data kwargs reach the model):
          * ``backend="path"``: stream the samples to disk (``hamiltonian/sink.py``; the layout of the reference's
            multicore samplers) instead of returning them -- ``posterior`` is then ``{var: filename}``.
          * ``adapt_step_size=True``: wire ``DualAveragingStepSize`` (hmc.py:141-176; instantiated but never used by
            the reference, :100-104) into burn-in: one ``update(mean accept prob over chains)`` per burn-in step, the
            noisy step size drives the next step, sampling runs at the averaged step size.  ``self.step_size`` is
            updated and the trace is kept in ``last_run['step_sizes']``."""
        backend = args.pop("backend", None)
        adapt = args.pop("adapt_step_size", False)
        target_accept = args.pop("target_accept", 0.8)
        niter, burnin = int(niter), int(burnin)
        h, shapes, squeeze, like, q0, s = self._setup(**args)
        s.set_q(q0)
        C = s.C
        if rng is not None:  # hmc.py:93 -- one momentum drawn and discarded
            for _ in range(C):
                for v in self._names():
                    rng.normal(0, 1, size=shapes[v])
        n_grad = 0
        accept_sum = 0.0
        chunk = max(1, min(512, (256 << 20) // max(1, 4 * C * h.P)))  # bound injected-tape / sample memory

        sink = None
        if backend is not None:
            from ...sink import SampleSink
            sink = SampleSink(backend, {v: shapes[v] for v in self.model.var_names}, niter, C, squeeze)
        step_sizes = []

        z_kept = []  # injected momentum tapes of the sampling iterations (rng given): sample_momentums reads them

        def run(n, keep, per_step=None):
            nonlocal n_grad, accept_sum
            outs = []
            done = 0
            while done < n:
                m = 1 if per_step is not None else min(chunk, n - done)
                kw = {}
                if rng is not None:
                    z, u1, u2 = self._host_draws(rng, m, C, h, shapes)
                    kw = dict(z_momentum=torch.as_tensor(z), u_path=u1, u_accept=u2)
                    if keep:
                        z_kept.append(z)
                o = s.hmc_run(m, self.step_size, self.path_length, step0=self._steps_done, keep_samples=keep, **kw)
                self._steps_done += m
                n_grad += o["n_grad_evals"]
                if keep:
                    smp = o["samples"].cpu().numpy()
                    if sink is not None:
                        sink.append(smp)
                        smp = smp[:0]
                    outs.append((smp, o["loss"].cpu().numpy(), o["accept_prob"].cpu().numpy()))
                else:
                    a_sum = float(o["accept_prob"].sum().item())
                    accept_sum += a_sum
                    if per_step is not None:
                        per_step(a_sum / C)
                done += m
            return outs

        if adapt and burnin > 0:
            da = DualAveragingStepSize(self.step_size, target_accept=target_accept)
            avg = [self.step_size]

            def tune(p_accept):
                self.step_size, a = da.update(p_accept)
                avg[0] = a
                step_sizes.append((float(p_accept), float(self.step_size), float(a)))

            run(burnin, False, per_step=tune)
            self.step_size = float(avg[0])
        else:
            run(burnin, False)
        if self.verbose and burnin > 0 and not adapt:
            _, avg = DualAveragingStepSize(self.step_size).update(accept_sum / max(1, burnin * C))
            print("adapted step size : ", avg)
        q_start = s.get(0) if niter > 0 else None  # position at the start of sampling iteration 0
        step_first = self._steps_done
        outs = run(niter, True)
        if outs:
            samples = np.concatenate([o[0] for o in outs], axis=0)
            loss = np.concatenate([o[1] for o in outs], axis=0)
            acc = np.concatenate([o[2] for o in outs], axis=0)
        else:
            samples = np.zeros((0, C, h.P), np.float32)
            loss = np.zeros((0, C))
            acc = np.zeros((0, C))
        if sink is not None:
            posterior = sink.close()
        else:
            posterior = self.model.unflatten(samples, shapes, squeeze, like)
        if squeeze:
            loss = loss[:, 0]
        if self.verbose and niter > 0:
            for i in range(0, niter, max(1, niter // 10)):
                print("loss: {0:.4f}".format(float(np.mean(loss[i]))))
        self.last_run = dict(n_grad_evals=n_grad, accept_prob=acc, n_chains=C, step_sizes=step_sizes,
                             step_size=self.step_size)
        # hmc.py:108-111,119: per-iteration [position], [momentum] at the start of each step (lazy, see _StepTrace).
        # With a disk backend the samples are not held in memory: positions are then read back from the sink's arrays
        # by the caller (posterior[var] names the files) and only the momentums are offered.
        unfl = self.model.unflatten
        positions = None
        if sink is None:
            positions = _StepTrace(niter, lambda i: unfl((q_start if i == 0 else samples[i - 1]).copy(), shapes, squeeze, like))
        if rng is not None:
            ztape = np.concatenate(z_kept, axis=0) if z_kept else np.zeros((0, C, h.P), np.float32)
            momentums = _StepTrace(niter, lambda i: unfl(ztape[i].copy(), shapes, squeeze, like))
        else:
            momentums = _StepTrace(niter, lambda i: unfl(self._drawn_momentum(s, step_first + i), shapes, squeeze, like))
        return posterior, loss, positions, momentums
