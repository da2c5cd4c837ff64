"""TEST INFRASTRUCTURE ONLY -- mints ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` is mounted):

    python -m oracle.make_golden

Each fixture stores the inputs, the exact random draws the reference consumed (recorded by
wrapping the ``rng`` argument and patching the module-level ``np.random.rand`` the
reference calls at ``hamiltonian/inference/cpu/hmc.py:46,61``) and the reference's outputs.
``tests/test_oracle_golden.py`` replays them through ``oracle/hamiltonian_oracle.py``;
``tests/test_gpu_parity.py`` replays them through the CUDA path.  The reference tree does
not exist on the GPU box, hence the committed vectors.
"""
import contextlib
import io
import os

import numpy as np

from oracle.ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class RecordingRng:
    """Stands in for the ``rng`` argument (a RandomState): records the *standard* normals."""

    def __init__(self, seed):
        self._rs = np.random.RandomState(seed)
        self.normals = []

    def normal(self, loc=0.0, scale=1.0, size=None):
        z = self._rs.normal(0.0, 1.0, size=size)
        self.normals.append(np.array(z, copy=True).ravel())
        return loc + scale * z  # == RandomState.normal(loc, scale, size) bit for bit


@contextlib.contextmanager
def patched_global_rand(seed=None, forced=None):
    """Replace ``np.random.rand`` (the global stream used for path length and accept) with a
    recorder.  ``forced`` = iterable of values to hand out first (to pin a given L)."""
    rs = np.random.RandomState(seed)
    forced = list(forced or [])
    rec = []
    orig = np.random.rand

    def rand(*a):
        assert not a
        u = forced.pop(0) if forced else float(rs.rand())
        rec.append(u)
        return u

    np.random.rand = rand
    try:
        yield rec
    finally:
        np.random.rand = orig


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def synth_softmax(n, d, k, seed, scale=1.0):
    rs = np.random.RandomState(seed)
    X = rs.rand(n, d) * scale
    w_true = rs.normal(0, 0.5, (d, k))
    y = np.argmax(X @ w_true + rs.gumbel(size=(n, k)), axis=1)
    return X, y


def pack_normals(lst):
    return np.concatenate(lst) if lst else np.zeros(0)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()

    # ---- 1. softmax grad / log-lik / NLP (models/cpu/softmax.py) -------------------------
    cases = {}
    for name, (n, d, k, alpha, wscale, xscale) in {
        "small": (64, 16, 3, 0.01, 0.3, 1.0),
        "k10": (96, 40, 10, 1.0, 0.2, 1.0),
        "clip": (48, 24, 5, 0.01, 6.0, 3.0),      # logits far beyond the +36.04 clip
        "k38": (40, 72, 38, 0.01, 0.1, 1.0),
    }.items():
        X, y = synth_softmax(n, d, k, seed=100 + len(cases))
        Y = ref.one_hot(y, k)
        rs = np.random.RandomState(7)
        par = {"weights": rs.normal(0, wscale, (d, k)), "bias": rs.normal(0, wscale, k)}
        X = X * xscale
        m = ref.softmax({"alpha": alpha})
        g = m.grad(par, X_train=X, y_train=Y)
        cases[name] = dict(X=X, y=y, W=par["weights"], b=par["bias"], alpha=alpha,
                           gW=g["weights"], gb=g["bias"],
                           ll=m.log_likelihood(par, X_train=X, y_train=Y),
                           nlp=m.negative_log_posterior(par, X_train=X, y_train=Y),
                           probs=m.net(par, X))
    np.savez_compressed(os.path.join(OUT, "softmax_model.npz"),
                        **{f"{c}.{k}": v for c, d_ in cases.items() for k, v in d_.items()})

    # ---- 2. hmc.step on softmax with pinned L (inference/cpu/hmc.py:39-64) ---------------
    n, d, k, alpha = 80, 12, 4, 0.01
    X, y = synth_softmax(n, d, k, seed=11)
    Y = ref.one_hot(y, k)
    eps, path = 2e-3, 2e-2  # eps*N small enough that proposals get accepted sometimes
    steps = {}
    for tag, u_len in {"L0": 0.0, "L1": 0.04, "L2": 0.09, "L5": 0.24, "L20": 0.99}.items():
        rs = np.random.RandomState(3)
        start = {"weights": rs.normal(0, 0.1, (d, k)), "bias": rs.normal(0, 0.1, k)}
        s = ref.hmc(ref.softmax({"alpha": alpha}), start, path_length=path, step_size=eps, verbose=False)
        rng = RecordingRng(5)
        with patched_global_rand(seed=9, forced=[u_len]) as urec:
            q, p, _, _, a = s.step(start, None, rng, X_train=X, y_train=Y)
        steps[tag] = dict(W0=start["weights"], b0=start["bias"], z=pack_normals(rng.normals),
                          u=np.array(urec), qW=q["weights"], qb=q["bias"], pW=p["weights"],
                          pb=p["bias"], accept_prob=a)
    # a non-finite energy case.  NB the reference's ``min(1, np.exp(..))`` (hmc.py:70) is the
    # *Python* builtin: min(1, nan) == 1, so a NaN proposal is ACCEPTED with A == 1 and the
    # ``np.isfinite(acceptprob)`` guard of hmc.py:61 can never fire.  This fixture pins that.
    rs = np.random.RandomState(3)
    start = {"weights": rs.normal(0, 0.1, (d, k)), "bias": rs.normal(0, 0.1, k)}
    # (alpha=1e4, eps=10: the linear part grows ~1e6x per iteration -> inf/nan after ~50)
    s = ref.hmc(ref.softmax({"alpha": 1e4}), start, path_length=300.0, step_size=10.0, verbose=False)
    rng = RecordingRng(5)
    with patched_global_rand(seed=9, forced=[0.9]) as urec, np.errstate(all="ignore"):
        q, p, _, _, a = s.step(start, None, rng, X_train=X, y_train=Y)
    assert a == 1 and len(urec) == 2 and np.isnan(q['weights']).all()
    steps["blowup"] = dict(W0=start["weights"], b0=start["bias"], z=pack_normals(rng.normals),
                           u=np.array(urec), qW=q["weights"], qb=q["bias"], pW=p["weights"],
                           pb=p["bias"], accept_prob=a, eps=10.0, path=300.0, alpha=1e4)
    np.savez_compressed(os.path.join(OUT, "hmc_step_softmax.npz"), X=X, y=y, alpha=alpha, eps=eps,
                        path=path, **{f"{c}.{k_}": v for c, d_ in steps.items() for k_, v in d_.items()})

    # ---- 3. hmc.sample short chain on softmax (hmc.py:90-119) -----------------------------
    rs = np.random.RandomState(4)
    start = {"weights": rs.normal(0, 0.05, (d, k)), "bias": np.zeros(k)}
    s = ref.hmc(ref.softmax({"alpha": alpha}), start, path_length=path, step_size=eps, verbose=False)
    rng = RecordingRng(21)
    with patched_global_rand(seed=22) as urec, quiet():
        post, loss, _, _ = s.sample(niter=12, burnin=4, rng=rng, X_train=X, y_train=Y)
    np.savez_compressed(os.path.join(OUT, "hmc_sample_softmax.npz"), X=X, y=y, alpha=alpha, eps=eps,
                        path=path, W0=start["weights"], b0=start["bias"], z=pack_normals(rng.normals),
                        u=np.array(urec), postW=post["weights"], postb=post["bias"], loss=loss,
                        niter=12, burnin=4)

    # ---- 4. sgld.sample across an epoch boundary (sgmcmc.py:40-86, sgld.py:31-46) ----------
    n2, bs = 130, 32  # 4 batches/epoch, remainder of 2 rows dropped
    X2, y2 = synth_softmax(n2, d, k, seed=12)
    Y2 = ref.one_hot(y2, k)
    start = {"weights": np.zeros((d, k)), "bias": np.zeros(k)}
    s = ref.sgld(ref.softmax({"alpha": alpha}), start, step_size=1e-2, verbose=False)
    rng = RecordingRng(31)
    with quiet():
        post, logp = s.sample(epochs=3, burnin=1, batch_size=bs, rng=rng, X_train=X2, y_train=Y2)
    np.savez_compressed(os.path.join(OUT, "sgld_sample_softmax.npz"), X=X2, y=y2, alpha=alpha, eps0=1e-2,
                        batch_size=bs, epochs=3, burnin=1, z=pack_normals(rng.normals),
                        postW=post["weights"], postb=post["bias"], logp=logp)

    # ---- 5. sghmc.step via the hmc mixin (sghmc.py:19-39), literal sign -----------------
    rs = np.random.RandomState(3)
    start = {"weights": rs.normal(0, 0.1, (d, k)), "bias": rs.normal(0, 0.1, k)}
    s = ref.sghmc_runnable(ref.softmax({"alpha": alpha}), start, path_length=path, step_size=eps, verbose=False)
    rng = RecordingRng(41)
    with patched_global_rand(seed=42, forced=[0.24]) as urec:
        q, p, a = s.step(start, None, rng, X_train=X, y_train=Y)
    np.savez_compressed(os.path.join(OUT, "sghmc_step_softmax.npz"), X=X, y=y, alpha=alpha, eps=eps,
                        path=path, W0=start["weights"], b0=start["bias"], z=pack_normals(rng.normals),
                        u=np.array(urec), qW=q["weights"], qb=q["bias"], pW=p["weights"], pb=p["bias"],
                        accept_prob=a)

    # ---- 6. hmc.sample on the 2-D Gaussian (cfg1; models/cpu/mvn_gaussian.py) ----------
    hyper = {"mu": np.zeros(2), "cov": np.array([[1.0, 0.8], [0.8, 1.0]])}
    s = ref.hmc(ref.mvn_gaussian(hyper), {"x": np.array([0.3, -0.2])}, path_length=1.0, step_size=0.1, verbose=False)
    rng = RecordingRng(51)
    with patched_global_rand(seed=52) as urec, quiet():
        post, loss, _, _ = s.sample(niter=300, burnin=20, rng=rng)
    np.savez_compressed(os.path.join(OUT, "hmc_sample_mvn.npz"), mu=hyper["mu"], cov=hyper["cov"],
                        x0=np.array([0.3, -0.2]), z=pack_normals(rng.normals), u=np.array(urec),
                        post=post["x"], loss=loss, niter=300, burnin=20, eps=0.1, path=1.0)

    # ---- 7. sgd.fit (sgd.py:25-45) -------------------------------------------------------
    start = {"weights": np.zeros((d, k)), "bias": np.zeros(k)}
    s = ref.sgd(ref.softmax({"alpha": alpha}), start, step_size=1e-2)
    with quiet():
        par, loss = s.fit(epochs=3, batch_size=bs, gamma=0.9, X_train=X2, y_train=Y2)
    np.savez_compressed(os.path.join(OUT, "sgd_fit_softmax.npz"), X=X2, y=y2, alpha=alpha, eps=1e-2,
                        batch_size=bs, epochs=3, gamma=0.9, W=par["weights"], b=par["bias"], loss=loss)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def main_next_rows():
    """logistic model (models/cpu/logistic.py), hmc.sample on it, sgd.fit_dropout (sgd.py:47-70)."""
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    # ---- 8. logistic grad / log-lik / NLP / predict ----------------------------------------
    cases = {}
    for name, (n, d, alpha, wscale, xscale) in {"small": (64, 16, 0.01, 0.3, 1.0), "d100": (96, 100, 1.0, 0.1, 1.0),
                                                 "clip": (48, 24, 0.01, 6.0, 3.0)}.items():
        rs = np.random.RandomState(200 + len(cases))
        X = rs.rand(n, d) * xscale
        w_true = rs.normal(0, 1.0, (d, 1))
        y = (rs.rand(n) < 1.0 / (1.0 + np.exp(-(X - X.mean()) @ w_true).ravel())).astype(np.float64)
        par = {"weights": rs.normal(0, wscale, (d, 1)), "bias": rs.normal(0, wscale, 1)}
        m = ref.logistic({"alpha": alpha})
        with np.errstate(all="ignore"):
            g = m.grad(par, X_train=X, y_train=y)
            cases[name] = dict(X=X, y=y, W=par["weights"], b=par["bias"], alpha=alpha, gW=g["weights"], gb=g["bias"],
                               ll=m.log_likelihood(par, X_train=X, y_train=y),
                               nlp=m.negative_log_posterior(par, X_train=X, y_train=y),
                               yhat=m.net(par, X_train=X), pred=m.predict(par, X, batchsize=32),
                               log_prior=m.log_prior(par))
    np.savez_compressed(os.path.join(OUT, "logistic_model.npz"),
                        **{f"{c}.{k}": v for c, d_ in cases.items() for k, v in d_.items()})
    # ---- 9. hmc.sample on logistic: the accept energy carries the quadratic prior term ---------
    n, d, alpha = 120, 10, 1.0
    rs = np.random.RandomState(61)
    X = rs.normal(0, 1, (n, d))
    y = (rs.rand(n) < 1.0 / (1.0 + np.exp(-X @ rs.normal(0, 1, d)))).astype(np.float64)
    start = {"weights": rs.normal(0, 0.1, (d, 1)), "bias": np.zeros(1)}
    eps, path = 5e-3, 5e-2
    s = ref.hmc(ref.logistic({"alpha": alpha}), start, path_length=path, step_size=eps, verbose=False)
    rng = RecordingRng(62)
    with patched_global_rand(seed=63) as urec, quiet():
        post, loss, _, _ = s.sample(niter=16, burnin=4, rng=rng, X_train=X, y_train=y)
    np.savez_compressed(os.path.join(OUT, "hmc_sample_logistic.npz"), X=X, y=y, alpha=alpha, eps=eps, path=path,
                        W0=start["weights"], b0=start["bias"], z=pack_normals(rng.normals), u=np.array(urec),
                        postW=post["weights"], postb=post["bias"], loss=loss, niter=16, burnin=4)
    # ---- 10. sgd.fit_dropout on softmax with the Bernoulli input masks recorded ------------------
    d, k, alpha, n2, bs = 12, 4, 0.01, 130, 32
    X2, y2 = synth_softmax(n2, d, k, seed=12)
    Y2 = ref.one_hot(y2, k)
    start = {"weights": np.zeros((d, k)), "bias": np.zeros(k)}
    s = ref.sgd(ref.softmax({"alpha": alpha}), start, step_size=1e-2)
    masks = []
    orig = np.random.binomial
    rs = np.random.RandomState(71)

    def binomial(nn, pp, size=None):
        z = rs.binomial(nn, pp, size=size)
        masks.append(np.array(z, dtype=np.uint8))
        return z

    np.random.binomial = binomial
    try:
        with quiet():
            par, loss = s.fit_dropout(epochs=3, batch_size=bs, gamma=0.9, p=0.7, X_train=X2, y_train=Y2)
    finally:
        np.random.binomial = orig
    np.savez_compressed(os.path.join(OUT, "sgd_fit_dropout_softmax.npz"), X=X2, y=y2, alpha=alpha, eps=1e-2,
                        batch_size=bs, epochs=3, gamma=0.9, p=0.7, masks=np.stack(masks), W=par["weights"],
                        b=par["bias"], loss=loss)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def main_mlp():
    """the UNMODIFIED ``hamiltonian/models/gpu/mlp.py`` executed under ``oracle/chainer_shim.py`` (fake
    chainer / cupy with Chainer's documented primitive semantics, derivatives by torch.autograd): grad, log_likelihood,
    negative_log_posterior, log_prior and the logits of ``predict`` with the three dropout keep-masks recorded."""
    from oracle import chainer_shim as S
    os.makedirs(OUT, exist_ok=True)
    M = S.load_reference_mlp()
    keys = ("/l1/W", "/l1/b", "/l2/W", "/l2/b", "/l3/W", "/l3/b")
    cases = {}
    for name, (B, n_in, n_mid, n_out, alpha, scale) in {"tiny": (37, 20, 24, 5, 0.05, 0.3), "tc": (80, 64, 64, 10, 0.01, 0.3),
                                                         "ragged": (67, 30, 72, 3, 0.5, 0.1)}.items():
        rs = np.random.RandomState(300 + len(cases))
        par = {"/l1/W": rs.normal(0, scale, (n_mid, n_in)), "/l1/b": rs.normal(0, scale, n_mid),
               "/l2/W": rs.normal(0, scale, (n_mid, n_mid)), "/l2/b": rs.normal(0, scale, n_mid),
               "/l3/W": rs.normal(0, scale, (n_out, n_mid)), "/l3/b": rs.normal(0, scale, n_out)}
        X = rs.rand(B, n_in)
        y = rs.randint(0, n_out, B)
        masks = [(rs.rand(B, n_mid) >= 0.1) for _ in range(3)]
        m = M.mlp({"alpha": alpha}, n_in, n_mid, n_out)
        assert [k for k, _ in m.net.namedparams()] == list(keys)

        def with_masks(fn):
            S.MASKS[:] = [a.copy() for a in masks]
            out = fn()
            assert not S.MASKS  # exactly three dropout calls per forward pass (mlp.py:29-31)
            return out

        g = with_masks(lambda: m.grad(par, X_train=X, y_train=y))
        d = dict(X=X, y=y, alpha=alpha, masks=np.stack(masks).astype(np.uint8),
                 loss=with_masks(lambda: m.log_likelihood(par, X_train=X, y_train=y)),
                 nlp=with_masks(lambda: m.negative_log_posterior(par, X_train=X, y_train=y)),
                 log_prior=m.log_prior(par), pred=with_masks(lambda: m.predict(par, X)))
        for k in keys:
            d["par" + k] = par[k]
            d["grad" + k] = g[k]
        cases[name] = d
    np.savez_compressed(os.path.join(OUT, "mlp_model.npz"), **{f"{c}.{k}": v for c, d_ in cases.items() for k, v in d_.items()})
    print("mlp_model.npz", os.path.getsize(os.path.join(OUT, "mlp_model.npz")))


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "next":
        main_next_rows()
    elif len(sys.argv) > 1 and sys.argv[1] == "mlp":
        main_mlp()
    else:
        main()
