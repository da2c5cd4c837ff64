"""CPU, build container only: run the UNMODIFIED reference (via oracle/ref_shim.py) and the
NumPy oracle side by side on live RandomState streams consumed in the same order.  Skipped
where /root/reference does not exist (the GPU box) -- tests/test_oracle_golden.py covers that."""
import contextlib
import io

import numpy as np
import pytest

from oracle import hamiltonian_oracle as O
from oracle.ref_shim import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def data(n=150, d=20, k=6, seed=0):
    rs = np.random.RandomState(seed)
    X = rs.rand(n, d)
    y = rs.randint(0, k, n)
    return X, y, O.one_hot(y, k)


def test_hmc_sample_softmax_live():
    ref = load_reference()
    X, y, Y = data()
    rs = np.random.RandomState(1)
    start = {"weights": rs.normal(0, .05, (20, 6)), "bias": rs.normal(0, .05, 6)}
    s = ref.hmc(ref.softmax({"alpha": 0.1}), {k: v.copy() for k, v in start.items()},
                path_length=1e-2, step_size=1e-3, verbose=False)
    np.random.seed(123)
    with quiet():
        post, loss, _, _ = s.sample(niter=15, burnin=5, rng=np.random.RandomState(77), X_train=X, y_train=Y)
    draws = O.StreamDraws(np.random.RandomState(77), np.random.RandomState(123))
    opost, oloss, info = O.hmc_sample(O.SoftmaxOracle({"alpha": 0.1}), start, 1e-3, 1e-2, 15, 5, draws,
                                      X_train=X, y_train=Y)
    np.testing.assert_allclose(opost["weights"], post["weights"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(opost["bias"], post["bias"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(oloss, loss, rtol=1e-12)
    assert len(np.unique(loss)) > 3  # the chain really moved


@pytest.mark.parametrize("sampler", ["sgld", "sgd"])
def test_minibatch_drivers_live(sampler):
    ref = load_reference()
    X, y, Y = data(n=210)
    start = {"weights": np.zeros((20, 6)), "bias": np.zeros(6)}
    if sampler == "sgld":
        s = ref.sgld(ref.softmax({"alpha": 0.1}), {k: v.copy() for k, v in start.items()}, step_size=5e-3, verbose=False)
        with quiet():
            post, logp = s.sample(epochs=3, burnin=2, batch_size=50, rng=np.random.RandomState(5), X_train=X, y_train=Y)
        draws = O.StreamDraws(np.random.RandomState(5), np.random.RandomState(0))
        opost, ologp, _ = O.sgmcmc_sample(O.sgld_step, O.SoftmaxOracle({"alpha": 0.1}), start, 5e-3, 3, 2, 50, draws, X, Y)
        np.testing.assert_allclose(opost["weights"], post["weights"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(ologp, logp, rtol=1e-12)
    else:
        s = ref.sgd(ref.softmax({"alpha": 0.1}), {k: v.copy() for k, v in start.items()}, step_size=5e-3)
        with quiet():
            par, loss = s.fit(epochs=4, batch_size=50, gamma=0.9, X_train=X, y_train=Y)
        opar, oloss = O.sgd_fit(O.SoftmaxOracle({"alpha": 0.1}), start, 5e-3, 4, 50, 0.9, X, Y)
        np.testing.assert_allclose(opar["weights"], par["weights"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(oloss, loss, rtol=1e-12)


@pytest.mark.parametrize("sign", ["reference"])
def test_sghmc_step_live(sign):
    ref = load_reference()
    X, y, Y = data()
    rs = np.random.RandomState(2)
    start = {"weights": rs.normal(0, .05, (20, 6)), "bias": rs.normal(0, .05, 6)}
    s = ref.sghmc_runnable(ref.softmax({"alpha": 0.1}), {k: v.copy() for k, v in start.items()},
                           path_length=1e-2, step_size=1e-3, verbose=False)
    np.random.seed(9)
    q, p, a = s.step(start, None, np.random.RandomState(8), X_train=X, y_train=Y)
    draws = O.StreamDraws(np.random.RandomState(8), np.random.RandomState(9))
    r = O.sghmc_step(O.SoftmaxOracle({"alpha": 0.1}), start, ["weights", "bias"], 1e-3, 1e-2, draws, sign=sign,
                     X_train=X, y_train=Y)
    np.testing.assert_allclose(r["accept_prob"], a, rtol=1e-10)
    np.testing.assert_allclose(r["q"]["weights"], q["weights"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(r["p"]["weights"], p["weights"], rtol=1e-10, atol=1e-13)


def test_mvn_live():
    ref = load_reference()
    hyper = {"mu": np.zeros(2), "cov": np.array([[1, .8], [.8, 1.]])}
    s = ref.hmc(ref.mvn_gaussian(hyper), {"x": np.zeros(2)}, path_length=1.0, step_size=0.1, verbose=False)
    np.random.seed(4)
    with quiet():
        post, loss, _, _ = s.sample(niter=200, burnin=10, rng=np.random.RandomState(3))
    draws = O.StreamDraws(np.random.RandomState(3), np.random.RandomState(4))
    opost, oloss, _ = O.hmc_sample(O.MvnGaussianOracle(hyper), {"x": np.zeros(2)}, 0.1, 1.0, 200, 10, draws)
    np.testing.assert_allclose(opost["x"], post["x"], rtol=1e-10, atol=1e-13)


def test_dual_averaging_live():
    ref = load_reference()
    da = ref.DualAveragingStepSize(0.1)
    st = dict(mu=np.log(10 * 0.1), target_accept=0.8, gamma=0.05, t=10.0, kappa=0.75, error_sum=0.0, log_averaged_step=0.0)
    for a in [0.9, 0.3, 0.7, 1.0, 0.0]:
        assert np.allclose(da.update(a), O.dual_averaging_update(st, a), rtol=1e-13)


def test_logistic_hmc_live():
    """models/cpu/logistic.py under hmc.sample: the accept energy carries the quadratic prior term."""
    ref = load_reference()
    rs = np.random.RandomState(11)
    X = rs.normal(0, 1, (90, 8))
    y = (rs.rand(90) < 0.5).astype(np.float64)
    start = {"weights": rs.normal(0, .1, (8, 1)), "bias": np.zeros(1)}
    s = ref.hmc(ref.logistic({"alpha": 0.5}), {k: v.copy() for k, v in start.items()}, path_length=5e-2, step_size=5e-3,
                verbose=False)
    np.random.seed(31)
    with quiet():
        post, loss, _, _ = s.sample(niter=12, burnin=3, rng=np.random.RandomState(32), X_train=X, y_train=y)
    draws = O.StreamDraws(np.random.RandomState(32), np.random.RandomState(31))
    opost, oloss, _ = O.hmc_sample(O.LogisticOracle({"alpha": 0.5}), start, 5e-3, 5e-2, 12, 3, draws, X_train=X, y_train=y)
    np.testing.assert_allclose(opost["weights"], post["weights"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(oloss, loss, rtol=1e-12)


def test_sgd_fit_dropout_live():
    ref = load_reference()
    X, y, Y = data(n=210)
    start = {"weights": np.zeros((20, 6)), "bias": np.zeros(6)}
    s = ref.sgd(ref.softmax({"alpha": 0.1}), {k: v.copy() for k, v in start.items()}, step_size=5e-3)
    np.random.seed(17)
    with quiet():
        par, loss = s.fit_dropout(epochs=3, batch_size=50, gamma=0.9, p=0.6, X_train=X, y_train=Y)
    rs = np.random.RandomState(17)
    masks = [rs.binomial(1, 0.6, size=(50, 20)) for _ in range(3 * 4)]
    opar, oloss = O.sgd_fit_dropout(O.SoftmaxOracle({"alpha": 0.1}), start, 5e-3, 3, 50, 0.9, X, Y, masks)
    np.testing.assert_allclose(opar["weights"], par["weights"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(oloss, loss, rtol=1e-12)


def test_public_api_surface_matches_the_reference():
    """Every public method of the reference's hot-path classes exists on the drop-in classes with the same leading
    positional parameters (module paths hamiltonian.inference.gpu.* / hamiltonian.models.gpu.*)."""
    import importlib
    import inspect
    load_reference()  # applies the compatibility shim and puts the reference tree on sys.path
    pairs = [("hamiltonian.inference.cpu.hmc", "hmc", "inference.gpu.hmc"), ("hamiltonian.inference.cpu.sgld", "sgld", "inference.gpu.sgld"),
             ("hamiltonian.inference.cpu.sghmc", "sghmc", "inference.gpu.sghmc"), ("hamiltonian.inference.cpu.sgd", "sgd", "inference.gpu.sgd"),
             ("hamiltonian.models.cpu.softmax", "softmax", "models.gpu.softmax"), ("hamiltonian.models.cpu.logistic", "logistic", "models.gpu.logistic"),
             ("hamiltonian.models.cpu.mvn_gaussian", "mvn_gaussian", "models.gpu.mvn_gaussian")]
    for ref_mod, cls, ours_mod in pairs:
        R = getattr(importlib.import_module(ref_mod), cls)
        Ours = getattr(importlib.import_module("dropout_hamiltonian_montecarlo_b200.hamiltonian." + ours_mod), cls)
        ref_methods = dict(inspect.getmembers(R, inspect.isfunction))
        our_methods = dict(inspect.getmembers(Ours, inspect.isfunction))
        for name, fn in ref_methods.items():
            assert name in our_methods, "%s.%s is missing" % (cls, name)
            want = [p.name for p in inspect.signature(fn).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
            got = [p.name for p in inspect.signature(our_methods[name]).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
            assert got[:len(want)] == want, "%s.%s%s vs reference %s" % (cls, name, got, want)


def test_mlp_reference_file_under_the_chainer_shim():
    """models/gpu/mlp.py needs Chainer + CuPy (not installable here).  oracle/chainer_shim.py supplies both as thin
    stand-ins (Chainer's documented primitive semantics, derivatives by torch.autograd) so that the UNMODIFIED reference
    file runs: layer order, dropout placement, loss reduction, prior terms, parameter names are the reference's own."""
    from oracle import chainer_shim as S
    M = S.load_reference_mlp()
    rs = np.random.RandomState(5)
    B, n_in, n_mid, n_out, alpha = 29, 17, 23, 6, 0.07
    par = {"/l1/W": rs.normal(0, .3, (n_mid, n_in)), "/l1/b": rs.normal(0, .3, n_mid),
           "/l2/W": rs.normal(0, .3, (n_mid, n_mid)), "/l2/b": rs.normal(0, .3, n_mid),
           "/l3/W": rs.normal(0, .3, (n_out, n_mid)), "/l3/b": rs.normal(0, .3, n_out)}
    X, y = rs.rand(B, n_in), rs.randint(0, n_out, B)
    masks = [(rs.rand(B, n_mid) >= 0.1).astype(np.float64) for _ in range(3)]
    m = M.mlp({"alpha": alpha}, n_in, n_mid, n_out)
    assert tuple(k for k, _ in m.net.namedparams()) == O.MLP_KEYS
    del S.DROPOUT_CALLS[:]
    S.MASKS[:] = [a.copy() for a in masks]
    g = m.grad(par, X_train=X, y_train=y)
    assert S.DROPOUT_CALLS == [(0.1, (B, n_mid))] * 3 and O.MLP_DROPOUT == 0.1  # mlp.py:29-31
    ref = O.mlp_grad(par, X, y, masks, alpha)
    for k in O.MLP_KEYS:
        np.testing.assert_allclose(g[k], ref[k], rtol=1e-10, atol=1e-13)
    S.MASKS[:] = [a.copy() for a in masks]
    np.testing.assert_allclose(m.log_likelihood(par, X_train=X, y_train=y), O.mlp_loss(par, X, y, masks), rtol=1e-12)
    S.MASKS[:] = [a.copy() for a in masks]
    np.testing.assert_allclose(m.negative_log_posterior(par, X_train=X, y_train=y), O.mlp_nlp(par, X, y, masks, alpha), rtol=1e-12)
    np.testing.assert_allclose(m.log_prior(par), O.mlp_log_prior(par, alpha), rtol=1e-12)
    S.MASKS[:] = [a.copy() for a in masks]
    np.testing.assert_array_equal(m.predict(par, X), O.mlp_forward(par, X, masks)[0].argmax(axis=1))
