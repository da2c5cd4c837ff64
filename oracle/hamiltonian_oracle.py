"""TEST INFRASTRUCTURE ONLY -- fp64 NumPy restatement of the reference's sampler hot path.

This file is the *oracle* (checker) for the CUDA path.  It is never the thing measured or
shipped: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``dropout_hamiltonian_montecarlo_b200``) must not, and fails loudly without its CUDA library.

Every function cites the reference ``file:line`` (relative to ``/root/reference``) whose
arithmetic it restates.  The restatement is deliberately *literal*: the reference's quirks
(SURVEY.md section 8(a')) define parity and are all reproduced here.

Pinning: the reference's own tests pin nothing (it has no tests), so the
oracle is pinned against outputs of the *unmodified reference itself*:
  * ``tests/test_oracle_vs_reference.py`` runs both side by side (same RandomState streams,
    identical consumption order) wherever ``/root/reference`` exists, and
  * ``oracle/make_golden.py`` stored reference outputs under ``tests/golden/*.npz`` which
    ``tests/test_oracle_golden.py`` replays everywhere (including the GPU box).
The MLP (``hamiltonian/models/gpu/mlp.py``) needs Chainer+CuPy, which cannot be installed here.  Its restatement is
pinned to the UNMODIFIED reference file executed under ``oracle/chainer_shim.py`` (stand-ins for chainer / cupy that
implement Chainer's documented primitive semantics with torch.autograd): live in ``tests/test_oracle_vs_reference.py``
and through the fixture ``tests/golden/mlp_model.npz``.  That pins everything the reference file itself decides (layer
order, dropout placement, loss reduction, prior terms, parameter names); the primitives are the shim's, not Chainer's
own code -- "pinned under a shim", stated as such in DESIGN.md.

Random draws are never generated in here: every stochastic function takes a ``Draws``
object that hands out standard normals / uniforms in exactly the order the reference
consumes them (``rng.normal`` in dict order, the *global* ``np.random.rand`` for path
length and accept).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

Par = Dict[str, np.ndarray]

# models/cpu/softmax.py:40-41 -- logits are clipped to [-log(1/tiny - 1), -log(eps)]
CLIP_HI = float(-np.log(np.finfo(float).eps))  # 36.04365338911715
CLIP_LO = float(-np.log(1.0 / np.finfo(float).tiny - 1.0))  # -708.3964185322641


# --------------------------------------------------------------------------------------
# draw sources
# --------------------------------------------------------------------------------------
class Draws:
    """Source of the random numbers a reference sampler consumes.

    ``normal(shape)`` -> standard normals (the reference calls ``rng.normal(0, s, shape)``,
    which NumPy evaluates as ``0 + s * z`` with ``z`` the next standard normals of the
    stream, so handing out ``z`` and scaling by ``s`` here is bit-identical).
    ``uniform()`` -> next value of the *global* ``np.random.rand()`` stream.
    """

    def normal(self, shape) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def uniform(self) -> float:  # pragma: no cover - interface
        raise NotImplementedError


class StreamDraws(Draws):
    """Live streams: ``rng`` plays the ``rng`` argument of ``sample``; ``glob`` plays the
    module-level ``np.random`` state.  Used to run oracle and reference side by side."""

    def __init__(self, rng: np.random.RandomState, glob: np.random.RandomState):
        self.rng, self.glob = rng, glob

    def normal(self, shape):
        return self.rng.normal(0.0, 1.0, size=shape)

    def uniform(self):
        return float(self.glob.rand())


class TapeDraws(Draws):
    """Replays pre-recorded draws (the arrays that are also injected into the CUDA path)."""

    def __init__(self, normals: Sequence[np.ndarray], uniforms: Sequence[float]):
        self._n = [np.asarray(a, dtype=np.float64) for a in normals]
        self._u = [float(u) for u in uniforms]
        self.n_pos = 0
        self.u_pos = 0

    def normal(self, shape):
        a = self._n[self.n_pos]
        self.n_pos += 1
        return a.reshape(shape).copy()

    def uniform(self):
        u = self._u[self.u_pos]
        self.u_pos += 1
        return u


class RecordingDraws(Draws):
    """Wraps another source and keeps everything it handed out (to build a tape)."""

    def __init__(self, inner: Draws):
        self.inner = inner
        self.normals: List[np.ndarray] = []
        self.uniforms: List[float] = []

    def normal(self, shape):
        a = self.inner.normal(shape)
        self.normals.append(np.array(a, copy=True))
        return a

    def uniform(self):
        u = self.inner.uniform()
        self.uniforms.append(u)
        return u


# --------------------------------------------------------------------------------------
# softmax regression model -- hamiltonian/models/cpu/softmax.py
# --------------------------------------------------------------------------------------
def softmax_logits(par: Par, X: np.ndarray) -> np.ndarray:
    """Clipped logits. softmax.py:39-41 (and identically :69-71)."""
    z = X @ par["weights"] + par["bias"]
    z = np.minimum(z, CLIP_HI)
    z = np.maximum(z, CLIP_LO)
    return z


def softmax_probs(z: np.ndarray) -> np.ndarray:
    """Row softmax with max subtraction. softmax.py:32-36."""
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def _logsumexp_rows(z: np.ndarray) -> np.ndarray:
    """scipy.special.logsumexp(z, axis=1) as used at softmax.py:18."""
    m = z.max(axis=1)
    return m + np.log(np.exp(z - m[:, None]).sum(axis=1))


def softmax_grad(par: Par, X: np.ndarray, Y: np.ndarray, alpha: float) -> Par:
    """Gradient of the potential, *summed* over rows. softmax.py:45-61:
    ``-(X^T (Y-P) - alpha W)`` and ``-(sum_n (Y-P) - alpha b)``."""
    diff = Y - softmax_probs(softmax_logits(par, X))
    gw = X.T @ diff - alpha * par["weights"]
    gb = diff.sum(axis=0) - alpha * par["bias"]
    return {"weights": -1.0 * gw, "bias": -1.0 * gb}


def softmax_log_likelihood(par: Par, X: np.ndarray, Y: np.ndarray) -> float:
    """sum_n sum_k Y_nk (z_nk - lse_n). softmax.py:63-72 with cross_entropy :17-20."""
    z = softmax_logits(par, X)
    return float(np.sum(Y * (z - _logsumexp_rows(z)[:, None])))


def softmax_log_prior(par: Par, alpha: float, variant: str = "cpu") -> float:
    """``cpu``: models/cpu/softmax.py:22-30 -- a constant, no quadratic term (quirk 7).
    ``gpu``: models/gpu/softmax.py:29-39 -- ``-0.5 alpha sum_v |theta_v|^2 / dim_v``."""
    k = 0.0
    for v in par:
        dim = np.asarray(par[v]).size
        if variant == "cpu":
            k -= 0.5 * dim * np.log(2 * np.pi) - 0.5 * dim * np.log(alpha)
        elif variant == "gpu":
            k -= 0.5 * alpha * np.sum(np.square(par[v])) / dim
        else:
            raise ValueError(variant)
    return float(k)


def softmax_nlp(par: Par, X, Y, alpha: float, variant: str = "cpu") -> float:
    """negative_log_posterior: ``-(LL + log_prior)/N`` -- a *mean* (quirk 6). softmax.py:74-79."""
    n = np.asarray(X).shape[0]
    return (-1.0 / n) * (softmax_log_likelihood(par, X, Y) + softmax_log_prior(par, alpha, variant))


class SoftmaxOracle:
    """Duck-typed like reference ``softmax(hyper)`` (seam 1 of SURVEY 8(b))."""

    def __init__(self, hyper, prior_variant: str = "cpu"):
        self.hyper = hyper
        self.variant = prior_variant

    def grad(self, par, **args):
        return softmax_grad(par, args["X_train"], args["y_train"], self.hyper["alpha"])

    def log_likelihood(self, par, **args):
        return softmax_log_likelihood(par, args["X_train"], args["y_train"])

    def negative_log_posterior(self, par, **args):
        return softmax_nlp(par, args["X_train"], args["y_train"], self.hyper["alpha"], self.variant)

    def predict(self, par, X, prob=False):
        """softmax.py:82-89."""
        p = softmax_probs(softmax_logits(par, X))
        return p if prob else p.argmax(axis=1)


# --------------------------------------------------------------------------------------
# logistic regression model -- hamiltonian/models/cpu/logistic.py
# --------------------------------------------------------------------------------------
def logistic_net(par: Par, X: np.ndarray) -> np.ndarray:
    """sigmoid of the clipped linear predictor, shape [N,1]. logistic.py:42-53."""
    z = X @ par["weights"] + par["bias"]
    z = np.minimum(z, CLIP_HI)
    z = np.maximum(z, CLIP_LO)
    return 1.0 / (1.0 + np.exp(-z))


def logistic_grad(par: Par, X: np.ndarray, y: np.ndarray, alpha: float) -> Par:
    """logistic.py:24-40: ``-(X^T (y - yhat) - alpha w)`` and ``-(sum (y - yhat) - alpha b)`` (sums over rows)."""
    diff = np.asarray(y).reshape(-1, 1) - logistic_net(par, X)
    gw = X.T @ diff - alpha * par["weights"]
    gb = diff.sum(axis=0) - alpha * par["bias"]
    return {"weights": -1.0 * gw, "bias": -1.0 * gb}


def logistic_log_likelihood(par: Par, X: np.ndarray, y: np.ndarray) -> float:
    """logistic.py:64-72: ``sum y log yhat + (1-y) log(1-yhat)``."""
    yp = np.squeeze(logistic_net(par, X), axis=1)
    y = np.asarray(y)
    return float(np.sum(y * np.log(yp) + (1.0 - y) * np.log(1.0 - yp)))


def logistic_log_prior(par: Par, alpha: float) -> float:
    """logistic.py:15-21: ``sum_v dim_v/2 log(alpha/2pi) - alpha/2 |theta_v|^2`` -- WITH the quadratic term."""
    k = 0.0
    for v in par:
        dim = np.asarray(par[v]).size
        k += dim * 0.5 * np.log(alpha / (2 * np.pi))
        k -= 0.5 * alpha * np.sum(np.square(par[v]))
    return float(k)


def logistic_nlp(par: Par, X, y, alpha: float) -> float:
    """logistic.py:55-60: ``-(LL + log_prior)/N``."""
    n = np.asarray(X).shape[0]
    return (-1.0 / n) * (logistic_log_likelihood(par, X, y) + logistic_log_prior(par, alpha))


class LogisticOracle:
    """Duck-typed like reference ``logistic(hyper)``."""

    def __init__(self, hyper):
        self.hyper = hyper

    def grad(self, par, **args):
        return logistic_grad(par, args["X_train"], args["y_train"], self.hyper["alpha"])

    def log_likelihood(self, par, **args):
        return logistic_log_likelihood(par, args["X_train"], args["y_train"])

    def negative_log_posterior(self, par, **args):
        return logistic_nlp(par, args["X_train"], args["y_train"], self.hyper["alpha"])

    def predict(self, par, X, prob=False, batchsize=32):
        """logistic.py:75-87: whole batches only (remainder dropped), flattened."""
        n = (X.shape[0] // batchsize) * batchsize
        yhat = logistic_net(par, X[:n])
        return yhat.flatten() if prob else (yhat > 0.5).astype(int).flatten()


# --------------------------------------------------------------------------------------
# 2-D Gaussian target -- hamiltonian/models/cpu/mvn_gaussian.py
# --------------------------------------------------------------------------------------
class MvnGaussianOracle:
    def __init__(self, hyper):
        self.hyper = hyper

    def grad(self, par, **args):
        """(x-mu) Sigma^-1. mvn_gaussian.py:14-20."""
        return {"x": np.dot(par["x"] - self.hyper["mu"], np.linalg.inv(self.hyper["cov"]))}

    def negative_log_posterior(self, par, **args):
        """0.5 (d log 2pi + log det Sigma + quad). mvn_gaussian.py:22-31."""
        mu, cov = self.hyper["mu"], self.hyper["cov"]
        d = par["x"] - mu
        return 0.5 * (mu.shape[0] * np.log(2 * np.pi) + np.log(np.linalg.det(cov))
                      + np.dot(np.dot(d.T, np.linalg.inv(cov)), d))

    log_likelihood = None


# --------------------------------------------------------------------------------------
# dropout MLP -- hamiltonian/models/gpu/mlp.py (Chainer; restated, pinned to the reference file under oracle/chainer_shim.py)
# --------------------------------------------------------------------------------------
MLP_KEYS = ("/l1/W", "/l1/b", "/l2/W", "/l2/b", "/l3/W", "/l3/b")
MLP_DROPOUT = 0.1  # mlp.py:29-31 ratio=.1


def mlp_forward(par: Par, X: np.ndarray, masks: Optional[Sequence[np.ndarray]]):
    """MyNetwork.forward, mlp.py:28-31: ``relu(dropout(l1 x))``, ``relu(dropout(l2 h))``,
    ``l3(dropout(h))``; Chainer ``L.Linear`` stores W as (out,in): y = x W^T + b; Chainer
    dropout is inverted dropout: ``x * mask / (1-ratio)``.  ``masks`` = three 0/1 keep-masks
    (shapes [B,n_mid],[B,n_mid],[B,n_mid]) or None for no dropout.  Returns (logits, cache)."""
    scale = 1.0 / (1.0 - MLP_DROPOUT)
    m = [None, None, None] if masks is None else [np.asarray(a, dtype=X.dtype) * scale for a in masks]
    a1 = X @ par["/l1/W"].T + par["/l1/b"]
    d1 = a1 if m[0] is None else a1 * m[0]
    h1 = np.maximum(d1, 0.0)
    a2 = h1 @ par["/l2/W"].T + par["/l2/b"]
    d2 = a2 if m[1] is None else a2 * m[1]
    h2 = np.maximum(d2, 0.0)
    h2d = h2 if m[2] is None else h2 * m[2]
    logits = h2d @ par["/l3/W"].T + par["/l3/b"]
    return logits, (X, m, d1, h1, d2, h2, h2d)


def mlp_loss(par: Par, X, y_int, masks) -> float:
    """F.softmax_cross_entropy: *mean* over the batch. mlp.py:57,66-78 (``log_likelihood``
    returns this positive loss)."""
    logits, _ = mlp_forward(par, X, masks)
    lse = _logsumexp_rows(logits)
    return float(np.mean(lse - logits[np.arange(len(y_int)), np.asarray(y_int, dtype=np.int64)]))


def mlp_log_prior(par: Par, alpha: float) -> float:
    """-0.5 alpha sum_v |theta_v|^2 / dim_v. mlp.py:40-45."""
    return float(-sum(0.5 * alpha * np.sum(np.square(par[v])) / np.asarray(par[v]).size for v in par))


def mlp_nlp(par, X, y_int, masks, alpha) -> float:
    """loss + log_prior (sic). mlp.py:80-82."""
    return mlp_loss(par, X, y_int, masks) + mlp_log_prior(par, alpha)


def mlp_grad(par: Par, X, y_int, masks, alpha: float) -> Par:
    """d(mean CE)/d theta + 0.5 alpha theta. mlp.py:47-64 (back-prop written out by hand)."""
    logits, (X, m, d1, h1, d2, h2, h2d) = mlp_forward(par, X, masks)
    B = X.shape[0]
    p = softmax_probs(logits)
    p[np.arange(B), np.asarray(y_int, dtype=np.int64)] -= 1.0
    dlog = p / B
    g = {}
    g["/l3/W"] = dlog.T @ h2d
    g["/l3/b"] = dlog.sum(axis=0)
    dh2d = dlog @ par["/l3/W"]
    dh2 = dh2d if m[2] is None else dh2d * m[2]
    dd2 = dh2 * (d2 > 0)
    da2 = dd2 if m[1] is None else dd2 * m[1]
    g["/l2/W"] = da2.T @ h1
    g["/l2/b"] = da2.sum(axis=0)
    dh1 = da2 @ par["/l2/W"]
    dd1 = dh1 * (d1 > 0)
    da1 = dd1 if m[0] is None else dd1 * m[0]
    g["/l1/W"] = da1.T @ X
    g["/l1/b"] = da1.sum(axis=0)
    return {k: g[k] + 0.5 * alpha * par[k] for k in par}


class MlpOracle:
    """Duck-typed like reference ``mlp(hyper, n_in, n_mid, n_out)``; dropout masks are supplied
    by ``mask_fn(batch_rows, n_mid) -> 3 masks`` (or None = dropout off)."""

    def __init__(self, hyper, mask_fn: Optional[Callable] = None):
        self.hyper = hyper
        self.mask_fn = mask_fn

    def _masks(self, par, X):
        return None if self.mask_fn is None else self.mask_fn(X.shape[0], par["/l1/b"].shape[0])

    def grad(self, par, **args):
        return mlp_grad(par, args["X_train"], args["y_train"], self._masks(par, args["X_train"]), self.hyper["alpha"])

    def log_likelihood(self, par, **args):
        return mlp_loss(par, args["X_train"], args["y_train"], self._masks(par, args["X_train"]))

    def negative_log_posterior(self, par, **args):
        X = args["X_train"]
        return mlp_nlp(par, X, args["y_train"], self._masks(par, X), self.hyper["alpha"])


# --------------------------------------------------------------------------------------
# HMC -- hamiltonian/inference/cpu/hmc.py
# --------------------------------------------------------------------------------------
def kinetic_energy(p: Par) -> float:
    """0.5 sum p^2 over all variables (identity mass). hmc.py:74-79."""
    return float(sum(0.5 * np.sum(np.square(p[v])) for v in p))


def path_length_steps(u: float, path_length: float, eps: float) -> int:
    """L = ceil(2 u path_length / eps); the leapfrog loop runs max(L-1,0) times. hmc.py:46,49."""
    return int(np.ceil(2 * u * path_length / eps))


def accept_probability(model, q, q_new, p, p_new, **data) -> float:
    """min(1, exp(E_cur - E_new)), E = NLP + kinetic. hmc.py:67-71.  ``min`` is the Python
    builtin, so a NaN energy difference yields 1 (``nan < 1`` is False) -- kept literally."""
    e_new = model.negative_log_posterior(q_new, **data) + kinetic_energy(p_new)
    e_cur = model.negative_log_posterior(q, **data) + kinetic_energy(p)
    with np.errstate(over="ignore", invalid="ignore"):
        return min(1, np.exp(e_cur - e_new))


def hmc_step(model, q: Par, var_order: Sequence[str], eps: float, path_length: float,
             draws: Draws, **data):
    """One HMC transition. hmc.py:39-64.

    Order of operations (quirks 1-5, 8): momentum redrawn per variable in dict order; L from
    the global uniform stream; one gradient at q; then for each of L-1 iterations a
    Gauss-Seidel sweep over variables, each doing ``p_v -= eps/2 g_v; q_v += eps p_v;
    g = grad(q) [all variables]; p_v -= eps g_v``; momentum negated; Metropolis test with a
    second global uniform that is *not drawn* when the acceptance probability is not finite.
    Returns dict(q, p, accept_prob, accepted, L, n_grad)."""
    q = {v: np.array(q[v], dtype=np.float64, copy=True) for v in var_order}
    p = {v: draws.normal(q[v].shape) for v in var_order}
    q_new = {v: q[v].copy() for v in var_order}
    p_new = {v: p[v].copy() for v in var_order}
    L = path_length_steps(draws.uniform(), path_length, eps)
    g = model.grad(q, **data)
    n_grad = 1
    for _ in range(max(L - 1, 0)):
        for v in var_order:
            p_new[v] = p_new[v] - (0.5 * eps) * g[v]
            q_new[v] = q_new[v] + eps * p_new[v]
            g = model.grad(q_new, **data)
            n_grad += 1
            p_new[v] = p_new[v] - eps * g[v]
    for v in var_order:
        p_new[v] = -p_new[v]
    a = accept_probability(model, q, q_new, p, p_new, **data)
    accepted = False
    if np.isfinite(a):
        accepted = bool(draws.uniform() < a)
    out_q, out_p = (q_new, p_new) if accepted else (q, p)
    return dict(q=out_q, p=out_p, accept_prob=float(a), accepted=accepted, L=L, n_grad=n_grad)


def hmc_sample(model, start: Par, eps: float, path_length: float, niter: int, burnin: int,
               draws: Draws, **data):
    """hmc.sample driver, hmc.py:90-119: one discarded momentum draw (:93), ``burnin`` steps,
    then ``niter`` steps each recording q and ``loss[i] = NLP(q)``."""
    order = list(start.keys())
    q = {v: np.asarray(start[v], dtype=np.float64) for v in order}
    for v in order:  # hmc.py:93 -- drawn and thrown away
        draws.normal(q[v].shape)
    accept_probs = []
    n_grad = 0
    for _ in range(int(burnin)):
        r = hmc_step(model, q, order, eps, path_length, draws, **data)
        q = r["q"]
        n_grad += r["n_grad"]
    post = {v: [] for v in order}
    loss = np.zeros(int(niter))
    for i in range(int(niter)):
        r = hmc_step(model, q, order, eps, path_length, draws, **data)
        q = r["q"]
        n_grad += r["n_grad"]
        accept_probs.append(r["accept_prob"])
        loss[i] = model.negative_log_posterior(q, **data)
        for v in order:
            post[v].append(q[v])
    return ({v: np.array(post[v]) for v in order}, loss,
            dict(accept_prob=np.array(accept_probs), n_grad=n_grad))


# --------------------------------------------------------------------------------------
# SGLD / SGHMC -- hamiltonian/inference/cpu/{sgmcmc,sgld,sghmc}.py
# --------------------------------------------------------------------------------------
def sgld_step(model, q: Par, var_order, eps: float, draws: Draws, **data):
    """sgld.py:31-39 with draw_momentum :41-46: ``eta_v ~ N(0, std = 2 eps)`` (quirk 10);
    ``p_v = eta_v - eps/2 g_v``; ``q_v += p_v``; the incoming momentum is ignored."""
    q = {v: np.array(q[v], dtype=np.float64, copy=True) for v in var_order}
    p = {v: (2.0 * eps) * draws.normal(q[v].shape) for v in var_order}
    g = model.grad(q, **data)
    for v in var_order:
        p[v] = p[v] + (-0.5 * eps * g[v])
        q[v] = q[v] + p[v]
    return q, p


def sghmc_step(model, q: Par, var_order, eps: float, path_length: float, draws: Draws,
               sign: str = "reference", **data):
    """sghmc.py:19-39 (runnable only through the hmc mixin): N(0,1) momentum; L as
    in HMC; per iteration per variable ``r ~ N(0, std 2 eps)`` drawn *before* the drift;
    ``q_v += eps p_v; g = grad(q); p_v = (1-eps) p_v + eps g_v + r``; no momentum flip; hmc's
    accept.  ``sign='reference'`` is the literal ``+eps g``; ``'descent'`` uses ``-eps g``."""
    sg = 1.0 if sign == "reference" else -1.0
    q = {v: np.array(q[v], dtype=np.float64, copy=True) for v in var_order}
    p = {v: draws.normal(q[v].shape) for v in var_order}
    q_new = {v: q[v].copy() for v in var_order}
    p_new = {v: p[v].copy() for v in var_order}
    L = path_length_steps(draws.uniform(), path_length, eps)
    g = model.grad(q, **data)
    n_grad = 1
    for _ in range(max(L - 1, 0)):
        for v in var_order:
            r = (2 * eps) * draws.normal(q[v].shape)
            q_new[v] = q_new[v] + eps * p_new[v]
            g = model.grad(q_new, **data)
            n_grad += 1
            p_new[v] = (1 - eps) * p_new[v] + sg * eps * g[v] + r
    a = accept_probability(model, q, q_new, p, p_new, **data)
    accepted = False
    if np.isfinite(a):
        accepted = bool(draws.uniform() < a)
    out_q, out_p = (q_new, p_new) if accepted else (q, p)
    return dict(q=out_q, p=out_p, accept_prob=float(a), accepted=accepted, L=L, n_grad=n_grad)


def minibatch_windows(n_rows: int, batch_size: int) -> List[Tuple[int, int]]:
    """Sequential, unshuffled, trailing remainder dropped. sgmcmc.py:34-38 (quirk 12)."""
    return [(s, s + batch_size) for s in range(0, n_rows - batch_size + 1, batch_size)]


def sgld_step_sizes(eps0: float, n_batches_seen: int) -> float:
    """lr_schedule, sgmcmc.py:88-89 with decay_factor*num_batches == eps0 (:50-51):
    ``eps0 / (1 + j eps0)``."""
    return eps0 * (1.0 / (1.0 + n_batches_seen * eps0))


def sgmcmc_sample(step_fn, model, start: Par, eps0: float, epochs: int, burnin: int,
                  batch_size: int, draws: Draws, X, Y):
    """sgmcmc.sample driver, sgmcmc.py:40-86: burn-in epochs at constant eps0; sampling epochs
    where the step size is re-assigned *after* batch j to ``eps0/(1+j eps0)`` (so batch 0 of
    a later epoch inherits the last value of the previous one, quirk 11); one stored sample
    and ``NLP(q, last batch)`` per epoch.  ``step_fn(model,q,order,eps,draws,**data)->(q,p)``."""
    order = list(start.keys())
    q = {v: np.asarray(start[v], dtype=np.float64) for v in order}
    num_batches = np.ceil(Y.shape[0] / float(batch_size))
    decay = eps0 / num_batches
    eps = eps0
    wins = minibatch_windows(X.shape[0], batch_size)
    n_grad = 0
    for _ in range(int(burnin)):
        for (a, b) in wins:
            q, _p = step_fn(model, q, order, eps, draws, X_train=X[a:b], y_train=Y[a:b])
            n_grad += 1
    logp = np.zeros(int(epochs))
    post = {v: [] for v in order}
    for i in range(int(epochs)):
        for j, (a, b) in enumerate(wins):
            q, _p = step_fn(model, q, order, eps, draws, X_train=X[a:b], y_train=Y[a:b])
            n_grad += 1
            eps = eps0 * (1.0 / (1.0 + j * decay * num_batches))
        a, b = wins[-1]
        logp[i] = model.negative_log_posterior(q, X_train=X[a:b], y_train=Y[a:b])
        for v in order:
            post[v].append(q[v])
    return {v: np.array(post[v]) for v in order}, logp, dict(n_grad=n_grad, final_eps=eps)


# --------------------------------------------------------------------------------------
# SGD -- hamiltonian/inference/cpu/sgd.py
# --------------------------------------------------------------------------------------
def sgd_fit(model, start: Par, eps: float, epochs: int, batch_size: int, gamma: float, X, Y):
    """Heavy-ball SGD, sgd.py:25-45: ``m = gamma m - eps g; theta += m`` per minibatch;
    ``loss[i] = NLP(theta, last batch)`` per epoch."""
    order = list(start.keys())
    par = {v: np.array(start[v], dtype=np.float64, copy=True) for v in order}
    mom = {v: np.zeros_like(par[v]) for v in order}
    loss = np.zeros(int(epochs))
    wins = minibatch_windows(X.shape[0], batch_size)
    for i in range(int(epochs)):
        for (a, b) in wins:
            g = model.grad(par, X_train=X[a:b], y_train=Y[a:b])
            for v in order:
                mom[v] = gamma * mom[v] - eps * g[v]
                par[v] = par[v] + mom[v]
        a, b = wins[-1]
        loss[i] = model.negative_log_posterior(par, X_train=X[a:b], y_train=Y[a:b])
    return par, loss


def sgd_fit_dropout(model, start: Par, eps: float, epochs: int, batch_size: int, gamma: float, X, Y, masks):
    """sgd.py:47-70: like ``fit`` but the gradient sees ``X_batch * Z`` with ``Z ~ Binomial(1, p)`` drawn per
    minibatch from the global ``np.random`` (``masks[k]`` = the k-th drawn Z, injected), and the per-epoch loss is
    ``-log_likelihood(theta, last batch)`` on the UNMASKED batch (:67)."""
    order = list(start.keys())
    par = {v: np.array(start[v], dtype=np.float64, copy=True) for v in order}
    mom = {v: np.zeros_like(par[v]) for v in order}
    loss = np.zeros(int(epochs))
    wins = minibatch_windows(X.shape[0], batch_size)
    k = 0
    for i in range(int(epochs)):
        for (a, b) in wins:
            g = model.grad(par, X_train=X[a:b] * masks[k], y_train=Y[a:b])
            k += 1
            for v in order:
                mom[v] = gamma * mom[v] - eps * g[v]
                par[v] = par[v] + mom[v]
        a, b = wins[-1]
        loss[i] = -1.0 * model.log_likelihood(par, X_train=X[a:b], y_train=Y[a:b])
    return par, loss


# --------------------------------------------------------------------------------------
# helpers shared by tests and the bench (data synthesis, flattening, ESS)
# --------------------------------------------------------------------------------------
def one_hot(y, k) -> np.ndarray:
    """utils.py:4-8."""
    out = np.zeros((len(y), k))
    out[np.arange(len(y)), np.asarray(y, dtype=np.int64)] = 1.0
    return out


def flatten_par(par: Par, order: Sequence[str]) -> np.ndarray:
    return np.concatenate([np.asarray(par[v], dtype=np.float64).ravel() for v in order])


def dual_averaging_update(state: dict, p_accept: float):
    """DualAveragingStepSize.update, hmc.py:153-176 (state: mu,t,gamma,kappa,error_sum,
    log_averaged_step,target_accept)."""
    state["error_sum"] += state["target_accept"] - p_accept
    log_step = state["mu"] - state["error_sum"] / (math.sqrt(state["t"]) * state["gamma"])
    eta = state["t"] ** -state["kappa"]
    state["log_averaged_step"] = eta * log_step + (1 - eta) * state["log_averaged_step"]
    state["t"] += 1
    return math.exp(log_step), math.exp(state["log_averaged_step"])
