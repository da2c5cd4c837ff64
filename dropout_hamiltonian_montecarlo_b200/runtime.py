"""Thin object layer over the C ABI: context, model and sampler handles.

PyTorch is used here only for device memory and streams (tensor.data_ptr() is what crosses
the ABI); every kernel that runs belongs to libbhmc.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BhmcError, HmcRun, SamplerConfig, SgRun, check

_contexts = {}


class Context:
    """One bhmc_ctx per (device, stream)."""

    def __init__(self, device=None, stream=None):
        L = _lib.lib()
        if not torch.cuda.is_available():
            # still ask the library so the error text comes from the product path
            h = C.c_void_p()
            check(L.bhmc_ctx_create(0, None, C.byref(h)))
            raise BhmcError("CUDA device required")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream() if stream is None else stream
            h = C.c_void_p()
            check(L.bhmc_ctx_create(self.device.index, C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.handle = h
        self.L = L

    def sync(self):
        check(self.L.bhmc_ctx_sync(self.handle))

    @property
    def launches(self):
        return int(self.L.bhmc_ctx_launch_count(self.handle))

    def timing(self, enable):
        check(self.L.bhmc_ctx_timing(self.handle, int(enable)))  # 0 off, 1 all kernel groups, 2 GEMM groups only

    def timing_stride(self, stride):
        check(self.L.bhmc_ctx_timing_stride(self.handle, int(stride)))

    def kernel_units(self, group):
        u = C.c_double()
        check(self.L.bhmc_ctx_kernel_units(self.handle, group, C.byref(u)))
        return u.value

    def kernel_time(self, group):
        ms, n = C.c_double(), C.c_int64()
        check(self.L.bhmc_ctx_kernel_time(self.handle, group, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)


def default_context(device=None):
    idx = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else (device or 0)
    if idx not in _contexts:
        _contexts[idx] = Context(idx)
    return _contexts[idx]


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class ModelHandle:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        L = ctx.L
        self.P = int(L.bhmc_model_n_params(handle))
        self.n_vars = int(L.bhmc_model_n_vars(handle))
        off = (C.c_int64 * self.n_vars)()
        ln = (C.c_int64 * self.n_vars)()
        check(L.bhmc_model_var_layout(handle, off, ln))
        self.var_off, self.var_len = list(off), list(ln)
        self.ld = (self.P + 3) // 4 * 4
        self._keep = []

    def close(self):
        if self.handle is not None:
            self.ctx.L.bhmc_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack(self, q):
        """[C, P] float tensor/array -> device [C, ld] fp32 (16-byte aligned rows)."""
        q = torch.as_tensor(q, dtype=torch.float32)
        if q.dim() == 1:
            q = q[None]
        out = self.ctx.zeros((q.shape[0], self.ld))
        out[:, : self.P] = q.to(self.ctx.device)
        return out

    def grad(self, q_dev, row0, nrows, precision, want_grad=True):
        """q_dev: [C, ld] device fp32 -> (g [C, ld] or None, stat [C] float64), on the context stream."""
        Cn = q_dev.shape[0]
        stat = self.ctx.empty((Cn,), torch.float64)
        if want_grad:
            g = self.ctx.empty((Cn, self.ld))
            check(self.ctx.L.bhmc_model_grad(self.handle, _ptr(q_dev), Cn, self.ld, row0, nrows, precision,
                                             _ptr(g), _ptr(stat)))
            return g, stat
        check(self.ctx.L.bhmc_model_loglik(self.handle, _ptr(q_dev), Cn, self.ld, row0, nrows, precision, _ptr(stat)))
        return None, stat

    def nlp(self, q_dev, row0, nrows, precision):
        Cn = q_dev.shape[0]
        out = self.ctx.empty((Cn,), torch.float64)
        check(self.ctx.L.bhmc_model_nlp(self.handle, _ptr(q_dev), Cn, self.ld, row0, nrows, precision, _ptr(out)))
        return out


class SoftmaxHandle(ModelHandle):
    def __init__(self, ctx, n_rows, n_features, n_classes, alpha, prior=0):
        h = C.c_void_p()
        check(ctx.L.bhmc_softmax_create(ctx.handle, n_rows, n_features, n_classes, float(alpha), prior, C.byref(h)))
        super().__init__(ctx, h)
        self.N, self.D, self.K = n_rows, n_features, n_classes

    def bind(self, X_dev, labels_dev, precision_mask=0b111):
        assert X_dev.dtype == torch.float32 and X_dev.is_contiguous() and X_dev.shape == (self.N, self.D)
        assert labels_dev.dtype == torch.int32 and labels_dev.is_contiguous() and labels_dev.shape == (self.N,)
        self._keep = [X_dev, labels_dev]  # the library keeps the pointers bound
        check(self.ctx.L.bhmc_softmax_bind_data(self.handle, _ptr(X_dev), _ptr(labels_dev), precision_mask))

    def bind_host(self, X_host, labels_host, precision_mask=0b111):
        """X_host/labels_host: pinned (or pageable) CPU tensors; the H2D copy happens inside the call."""
        assert X_host.dtype == torch.float32 and X_host.is_contiguous() and tuple(X_host.shape) == (self.N, self.D)
        assert labels_host.dtype == torch.int32 and labels_host.is_contiguous()
        self._keep = [X_host, labels_host]
        check(self.ctx.L.bhmc_softmax_bind_data_host(self.handle, C.c_void_p(X_host.data_ptr()),
                                                     C.c_void_p(labels_host.data_ptr()), precision_mask))

    def operand_info(self):
        """(exact, x_scale) found by the last bind: exact -> bf16x3 runs 2 MMAs per product on bf16(x_scale * X)."""
        ex, sc = C.c_int32(0), C.c_float(1.0)
        check(self.ctx.L.bhmc_softmax_operand_info(self.handle, C.byref(ex), C.byref(sc)))
        return bool(ex.value), float(sc.value)

    def predict(self, q_dev, X_dev, want_probs=True, want_labels=True):
        Cn, n = q_dev.shape[0], X_dev.shape[0]
        probs = self.ctx.empty((Cn, n, self.K)) if want_probs else None
        labels = self.ctx.empty((Cn, n), torch.int32) if want_labels else None
        check(self.ctx.L.bhmc_softmax_predict(self.handle, _ptr(q_dev), Cn, self.ld, _ptr(X_dev), n, _ptr(probs),
                                              _ptr(labels)))
        return probs, labels


class LogisticHandle(SoftmaxHandle):
    """models/cpu/logistic.py on the softmax kernels (two-class softmax, class 0 pinned to zero): P = D + 1."""

    def __init__(self, ctx, n_rows, n_features, alpha):
        h = C.c_void_p()
        check(ctx.L.bhmc_logistic_create(ctx.handle, n_rows, n_features, float(alpha), C.byref(h)))
        ModelHandle.__init__(self, ctx, h)
        self.N, self.D, self.K = n_rows, n_features, 2


class MlpHandle(ModelHandle):
    def __init__(self, ctx, n_rows, n_in, n_mid, n_out, alpha, ratio=0.1, seed=0, chain_id0=0):
        h = C.c_void_p()
        check(ctx.L.bhmc_mlp_create(ctx.handle, n_rows, n_in, n_mid, n_out, float(alpha), float(ratio), seed, chain_id0,
                                    C.byref(h)))
        super().__init__(ctx, h)
        self.N, self.n_in, self.n_mid, self.n_out = n_rows, n_in, n_mid, n_out

    def bind(self, X, labels):
        """X [N, n_in] fp32 and labels [N] int32: CUDA tensors stay bound, CPU tensors are copied by the library."""
        assert X.dtype == torch.float32 and X.is_contiguous() and tuple(X.shape) == (self.N, self.n_in)
        assert labels.dtype == torch.int32 and labels.is_contiguous()
        self._keep = [X, labels]
        check(self.ctx.L.bhmc_mlp_bind_data(self.handle, C.c_void_p(X.data_ptr()), C.c_void_p(labels.data_ptr()),
                                            0 if X.is_cuda else 1))
        if not X.is_cuda:
            self.ctx.sync()

    def predict(self, q_dev, X_dev, precision, want_probs=True, want_labels=True):
        """mlp.predict on caller rows: (probs [C, n, n_out] or None, labels [C, n] or None)."""
        Cn, n = q_dev.shape[0], X_dev.shape[0]
        probs = self.ctx.empty((Cn, n, self.n_out)) if want_probs else None
        labels = self.ctx.empty((Cn, n), torch.int32) if want_labels else None
        check(self.ctx.L.bhmc_mlp_predict(self.handle, _ptr(q_dev), Cn, self.ld, _ptr(X_dev), n, precision, _ptr(probs),
                                          _ptr(labels)))
        return probs, labels

    def set_masks(self, masks):
        """masks: uint8 CUDA tensor [3, C, B, n_mid] of keep flags, or None for Philox dropout."""
        self._masks = masks
        check(self.ctx.L.bhmc_mlp_set_masks(self.handle, _ptr(masks)))


class MvnHandle(ModelHandle):
    def __init__(self, ctx, mu, cov):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        cov = np.asarray(cov, dtype=np.float64)
        cinv = np.ascontiguousarray(np.linalg.inv(cov))
        h = C.c_void_p()
        check(ctx.L.bhmc_mvn_create(ctx.handle, mu.shape[0], mu.ctypes.data_as(C.c_void_p),
                                    cinv.ctypes.data_as(C.c_void_p), float(np.log(np.linalg.det(cov))), C.byref(h)))
        super().__init__(ctx, h)


class SamplerHandle:
    """bhmc_sampler: resident chain state [C, ld] + the step drivers."""

    def __init__(self, ctx, model, kind, n_chains, *, seed=0, chain_id0=0, precision=1, sweep=None,
                 shared_path=False, leapfrog=False, sghmc_descent=False, reject_nan=False):
        self.ctx, self.model = ctx, model
        cfg = SamplerConfig()
        cfg.kind = kind
        cfg.n_chains = n_chains
        cfg.chain_id0 = chain_id0
        cfg.seed = seed
        cfg.precision = precision
        groups = sweep if sweep is not None else list(zip(model.var_off, model.var_len))
        cfg.n_sweep = len(groups)
        for i, (o, l) in enumerate(groups):
            cfg.sweep_off[i], cfg.sweep_len[i] = o, l
        cfg.shared_path = int(shared_path)
        cfg.leapfrog = int(leapfrog)
        cfg.sghmc_descent = int(sghmc_descent)
        cfg.reject_nan = int(reject_nan)
        self.cfg = cfg
        h = C.c_void_p()
        check(ctx.L.bhmc_sampler_create(ctx.handle, model.handle, C.byref(cfg), C.byref(h)))
        self.handle = h
        self.C, self.P = n_chains, model.P
        self.n_sweep = len(groups)

    def close(self):
        if self.handle is not None:
            self.ctx.L.bhmc_sampler_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_q(self, q):
        """q: [C, P] numpy / CPU tensor / CUDA tensor (fp32)."""
        if isinstance(q, torch.Tensor) and q.is_cuda:
            q = q.to(torch.float32).contiguous()
            assert tuple(q.shape) == (self.C, self.P)
            check(self.ctx.L.bhmc_sampler_set_q(self.handle, _ptr(q), 0))
            self.ctx.sync()
        else:
            a = np.ascontiguousarray(np.asarray(q, dtype=np.float32).reshape(self.C, self.P))
            check(self.ctx.L.bhmc_sampler_set_q(self.handle, a.ctypes.data_as(C.c_void_p), 1))
            self.ctx.sync()

    def get(self, which=0):
        out = np.empty((self.C, self.P), dtype=np.float32)
        check(self.ctx.L.bhmc_sampler_get(self.handle, which, out.ctypes.data_as(C.c_void_p), 1))
        return out

    def hmc_run(self, n_steps, step_size, path_length, *, row0=0, nrows=0, step0=0, z_momentum=None, u_path=None,
                u_accept=None, z_noise=None, keep_samples=True, keep_stats=True, schedule="auto"):
        """Runs n_steps transitions.  Injected draws: z_momentum [n,C,P] device fp32; u_path / u_accept
        [n,C] host float64; z_noise [n, iters, C, P] device fp32 (SGHMC).  Returns a dict of device tensors."""
        ctx = self.ctx
        run = HmcRun()
        run.n_steps, run.step_size, run.path_length = n_steps, step_size, path_length
        run.row0, run.nrows, run.step0 = row0, nrows, step0
        run.schedule = {"auto": 0, "lockstep": 1, "streaming": 2}[schedule]
        keep = []
        if z_momentum is not None:
            z_momentum = z_momentum.to(ctx.device, torch.float32).contiguous()
            assert z_momentum.numel() == n_steps * self.C * self.P
            run.z_momentum_dev = z_momentum.data_ptr()
        for name, arr in (("u_path_host", u_path), ("u_accept_host", u_accept)):
            if arr is not None:
                a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(n_steps, self.C))
                keep.append(a)
                setattr(run, name, a.ctypes.data)
        if z_noise is not None:
            z_noise = z_noise.to(ctx.device, torch.float32).contiguous()
            run.z_noise_dev = z_noise.data_ptr()
            run.z_noise_iters = z_noise.shape[1]
        out = {}
        if keep_samples:
            out["samples"] = ctx.empty((n_steps, self.C, self.P))
            run.samples_dev = out["samples"].data_ptr()
        if keep_stats:
            out["loss"] = ctx.empty((n_steps, self.C), torch.float64)
            out["accept_prob"] = ctx.empty((n_steps, self.C), torch.float64)
            out["accepted"] = ctx.empty((n_steps, self.C), torch.int32)
            run.loss_dev = out["loss"].data_ptr()
            run.accept_prob_dev = out["accept_prob"].data_ptr()
            run.accepted_dev = out["accepted"].data_ptr()
        check(ctx.L.bhmc_sampler_hmc_run(self.handle, C.byref(run)))
        out["n_grad_evals"] = int(run.n_grad_evals)
        out["n_grad_launched"] = int(run.n_grad_launched)
        out["n_phases"] = int(run.n_phases)  # gradient launches of the streaming schedule (0 = lockstep ran)
        out["_keep"] = (keep, z_momentum, z_noise)
        return out

    def sg_run(self, epochs, burnin, batch_size, step_size, *, n_rows=0, gamma=0.9, step0=0, z=None,
               keep_samples=True, dropout_keep=0.0, masks=None, first_step_size=0.0, keep_momentum=False):
        ctx = self.ctx
        run = SgRun()
        run.epochs, run.burnin, run.batch_size, run.n_rows = epochs, burnin, batch_size, n_rows
        run.step_size, run.gamma, run.step0 = step_size, gamma, step0
        run.first_step_size, run.keep_momentum = float(first_step_size), int(bool(keep_momentum))
        if z is not None:
            z = z.to(ctx.device, torch.float32).contiguous()
            run.z_dev = z.data_ptr()
        run.dropout_keep = float(dropout_keep)
        if masks is not None:  # [(burnin+epochs)*n_batches, batch_size, D] keep flags (sgd.fit_dropout injection)
            masks = masks.to(ctx.device, torch.uint8).contiguous()
            run.mask_dev = masks.data_ptr()
        out = {}
        if keep_samples:
            out["samples"] = ctx.empty((epochs, self.C, self.P))
            run.samples_dev = out["samples"].data_ptr()
        out["logp"] = ctx.empty((epochs, self.C), torch.float64)
        run.logp_dev = out["logp"].data_ptr()
        check(ctx.L.bhmc_sampler_sg_run(self.handle, C.byref(run)))
        out["n_grad_evals"] = int(run.n_grad_evals)
        out["final_step_size"] = float(run.final_step_size)
        out["_keep"] = (z, masks)
        return out
