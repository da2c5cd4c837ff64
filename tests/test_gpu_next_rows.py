"""GPU (B200): the SURVEY 8(f) "next" rows through the C ABI and the reference-facing classes, against golden
vectors minted from the unmodified reference (oracle/make_golden.py next): the logistic model, HMC on it (a
non-degenerate accept path: its energy carries the quadratic prior), sgd.fit_dropout with the reference's recorded
Bernoulli masks, the on-disk sample sink + backend_mean, and DualAveragingStepSize wired into burn-in.
Tolerance as in test_gpu_parity.py: rtol 1e-4 on trajectories, identical accept decisions."""
import os

import numpy as np
import pytest
import torch

from conftest import TapeRng, load_golden, replay_uniforms
from oracle import hamiltonian_oracle as O

pytestmark = pytest.mark.gpu

from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.sgd import sgd  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.logistic import logistic  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.mvn_gaussian import mvn_gaussian  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax  # noqa: E402
from test_gpu_parity import close  # noqa: E402

PRECS = ["fp32", "bf16x3"]


@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("case", ["small", "d100", "clip"])
def test_logistic_model_golden(prec, case):
    g = load_golden("logistic_model.npz")[case]
    m = logistic({"alpha": g["alpha"]}, precision=prec)
    par = {"weights": g["W"], "bias": g["b"]}
    got = m.grad(par, X_train=g["X"], y_train=g["y"])
    rtol = {"fp32": 2e-5, "bf16x3": 1e-4, "bf16": 5e-2}[prec]
    scale = {"fp32": 2e-6, "bf16x3": 2e-5, "bf16": 2e-2}[prec]
    if case == "clip":  # saturated logits: |z| ~ 1e2, the reference's log(1 - yhat) is -inf for some rows
        rtol, scale = rtol * 20, scale * 20
    assert got["weights"].shape == g["gW"].shape and got["bias"].shape == g["gb"].shape
    close(got["weights"], g["gW"], rtol, scale, "grad weights")
    close(got["bias"], g["gb"], rtol, scale, "grad bias")
    if np.isfinite(g["ll"]):
        close(m.log_likelihood(par, X_train=g["X"], y_train=g["y"]), g["ll"], rtol * 2, scale)
        close(m.negative_log_posterior(par, X_train=g["X"], y_train=g["y"]), g["nlp"], rtol * 2, scale)
    close(m.log_prior(par), g["log_prior"], 1e-6)
    if prec == "fp32":
        close(m.net(par, X_train=g["X"]), g["yhat"], 1e-4, 1e-6)
        assert (m.predict(par, g["X"], batchsize=32) == g["pred"]).mean() > 0.99
    # chain batch: two different chains in one launch
    par2 = {"weights": np.stack([g["W"], 0.5 * g["W"]]), "bias": np.stack([g["b"], -g["b"]])}
    got2 = m.grad(par2, X_train=g["X"], y_train=g["y"])
    ref1 = O.logistic_grad({"weights": 0.5 * g["W"], "bias": -g["b"]}, g["X"], g["y"], g["alpha"])
    close(got2["weights"][0], g["gW"], rtol, scale)
    close(got2["weights"][1], ref1["weights"], rtol, scale)


@pytest.mark.parametrize("prec", PRECS)
def test_hmc_sample_logistic_golden(prec):
    g = load_golden("hmc_sample_logistic.npz")
    s = hmc(logistic({"alpha": g["alpha"]}, precision=prec), {"weights": g["W0"], "bias": g["b0"]},
            path_length=g["path"], step_size=g["eps"], verbose=False)
    with replay_uniforms(g["u"]) as ru:
        post, loss, _, _ = s.sample(niter=g["niter"], burnin=g["burnin"], rng=TapeRng(g["z"]), X_train=g["X"], y_train=g["y"])
    assert ru.pos == g["u"].size
    close(post["weights"], g["postW"], 1e-4, 2e-6)
    close(post["bias"], g["postb"], 1e-4, 2e-6)
    close(loss, g["loss"], 1e-5)
    moved_ref = np.any(np.diff(g["postW"], axis=0) != 0, axis=(1, 2))
    moved_got = np.any(np.diff(post["weights"], axis=0) != 0, axis=(1, 2))
    assert (moved_ref == moved_got).all() and moved_ref.any() and not moved_ref.all()


class replay_binomial:
    """np.random.binomial hands out the Bernoulli masks the reference drew."""

    def __init__(self, masks):
        self.masks, self.pos = masks, 0

    def __enter__(self):
        self._orig = np.random.binomial

        def binomial(n, p, size=None):
            m = self.masks[self.pos]
            assert tuple(size) == m.shape
            self.pos += 1
            return m

        np.random.binomial = binomial
        return self

    def __exit__(self, *exc):
        np.random.binomial = self._orig
        return False


@pytest.mark.parametrize("prec", PRECS)
def test_sgd_fit_dropout_golden(prec):
    g = load_golden("sgd_fit_dropout_softmax.npz")
    d, K = g["W"].shape
    s = sgd(softmax({"alpha": g["alpha"]}, precision=prec), {"weights": np.zeros((d, K)), "bias": np.zeros(K)},
            step_size=g["eps"])
    with replay_binomial(g["masks"]) as rb:
        par, loss = s.fit_dropout(epochs=g["epochs"], batch_size=g["batch_size"], gamma=g["gamma"], p=g["p"], rng="numpy",
                                  X_train=g["X"], y_train=O.one_hot(g["y"], K))
    assert rb.pos == g["masks"].shape[0]
    close(par["weights"], g["W"], 1e-4, 2e-6)
    close(par["bias"], g["b"], 1e-4, 2e-6)
    close(loss, g["loss"], 1e-5)


def test_sgd_fit_dropout_philox_masks():
    """In-kernel masks: keep rate p, different masks per minibatch, chains differ only through their start."""
    rs = np.random.RandomState(0)
    n, d, K = 512, 64, 5
    X = rs.rand(n, d)
    y = rs.randint(0, K, n)
    Y = O.one_hot(y, K)
    start = {"weights": np.zeros((d, K)), "bias": np.zeros(K)}
    s = sgd(softmax({"alpha": 0.01}, precision="fp32"), start, step_size=1e-3, seed=5)
    par_keep1, _ = s.fit_dropout(epochs=2, batch_size=128, gamma=0.9, p=1.0, X_train=X, y_train=Y)
    ref, _ = O.sgd_fit_dropout(O.SoftmaxOracle({"alpha": 0.01}), start, 1e-3, 2, 128, 0.9, X, Y, [np.ones((128, d))] * 8)
    close(par_keep1["weights"], ref["weights"], 1e-4, 2e-6)  # p = 1: every input kept
    s2 = sgd(softmax({"alpha": 0.01}, precision="fp32"), start, step_size=1e-3, seed=5)
    par_half, loss = s2.fit_dropout(epochs=2, batch_size=128, gamma=0.9, p=0.5, X_train=X, y_train=Y)
    # E[X*Z] = p X: the first heavy-ball step is -eps * grad at theta = 0, linear in the inputs
    s3 = sgd(softmax({"alpha": 0.01}, precision="fp32"), start, step_size=1e-3, seed=5)
    one, _ = s3.fit_dropout(epochs=1, batch_size=512, gamma=0.0, p=0.5, X_train=X, y_train=Y)
    full, _ = O.sgd_fit_dropout(O.SoftmaxOracle({"alpha": 0.01}), start, 1e-3, 1, 512, 0.0, X, Y, [np.ones((512, d))])
    ratio = float(np.sum(one["weights"] * full["weights"]) / np.sum(full["weights"] ** 2))
    assert 0.45 < ratio < 0.55, ratio
    assert np.isfinite(loss).all() and not np.allclose(par_half["weights"], par_keep1["weights"])


def test_sample_sink_and_backend_mean(tmp_path):
    hyper = {"mu": np.zeros(2), "cov": np.array([[1.0, 0.8], [0.8, 1.0]])}
    kw = dict(path_length=1.0, step_size=0.1, verbose=False, n_chains=8, seed=3)
    a = hmc(mvn_gaussian(hyper), {"x": np.zeros(2)}, **kw)
    post, loss, _, _ = a.sample(niter=50, burnin=5)
    b = hmc(mvn_gaussian(hyper), {"x": np.zeros(2)}, **kw)
    backend = os.path.join(str(tmp_path), "chains0")
    files, loss_b, _, _ = b.sample(niter=50, burnin=5, backend=backend)
    disk = np.load(files["x"])
    assert disk.dtype == np.float32 and disk.shape == (50, 8, 2)
    np.testing.assert_allclose(disk, post["x"], rtol=0, atol=1e-7)  # same seeds -> same chains, fp32 on disk
    np.testing.assert_allclose(loss_b, loss)
    mean = b.backend_mean([backend], 50)
    np.testing.assert_allclose(mean["x"], post["x"].sum(axis=0) / 50, rtol=1e-5, atol=1e-6)


def test_dual_averaging_wired_into_burnin():
    """adapt_step_size=True: the step sizes used in burn-in are exactly DualAveragingStepSize.update() (hmc.py:141-176)
    fed with the observed mean accept probabilities, and sampling runs at the averaged step size."""
    hyper = {"mu": np.zeros(2), "cov": np.array([[1.0, 0.8], [0.8, 1.0]])}
    s = hmc(mvn_gaussian(hyper), {"x": np.zeros(2)}, path_length=1.0, step_size=0.05, verbose=False, n_chains=64, seed=1)
    s.sample(niter=20, burnin=12, adapt_step_size=True, target_accept=0.8)
    trace = s.last_run["step_sizes"]  # (observed accept prob, noisy step used next, averaged step)
    assert len(trace) == 12
    st = dict(mu=np.log(10 * 0.05), target_accept=0.8, gamma=0.05, t=10.0, kappa=0.75, error_sum=0.0, log_averaged_step=0.0)
    for p_acc, noisy, avg in trace:
        assert 0.0 <= p_acc <= 1.0
        e_noisy, e_avg = O.dual_averaging_update(st, p_acc)
        np.testing.assert_allclose([noisy, avg], [e_noisy, e_avg], rtol=1e-12)
    np.testing.assert_allclose(s.last_run["step_size"], trace[-1][2], rtol=1e-12)
    assert len({round(t[0], 6) for t in trace}) > 1  # the acceptance really responded to the step size


def test_hmc_accept_building_block():
    """hmc.accept (hmc.py:67-71) composed from the device NLP and the host kinetic energy == the oracle's."""
    rs = np.random.RandomState(4)
    n, d, K = 200, 12, 4
    X = rs.rand(n, d)
    Y = O.one_hot(rs.randint(0, K, n), K)
    q1 = {"weights": rs.normal(0, .1, (d, K)), "bias": rs.normal(0, .1, K)}
    q2 = {"weights": q1["weights"] + rs.normal(0, .01, (d, K)), "bias": q1["bias"] + rs.normal(0, .01, K)}
    s = hmc(softmax({"alpha": 0.1}, precision="fp32"), q1, path_length=1e-2, step_size=1e-3, verbose=False)
    p1, p2 = s.draw_momentum(np.random.RandomState(1)), s.draw_momentum(np.random.RandomState(2))
    a = s.accept(q1, q2, p1, p2, X_train=X, y_train=Y)
    ref = O.accept_probability(O.SoftmaxOracle({"alpha": 0.1}), q1, q2, p1, p2, X_train=X, y_train=Y)
    close(a, ref, 1e-4, 1e-6)
