"""CPU: the NumPy oracle replays every golden vector minted from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest

from oracle import hamiltonian_oracle as O
from conftest import load_golden

TOL = dict(rtol=1e-10, atol=1e-12)


def split_tape(z, shapes, reps):
    """Cut the flat recorded standard-normal tape into per-call arrays."""
    out, pos = [], 0
    for _ in range(reps):
        for s in shapes:
            n = int(np.prod(s))
            out.append(z[pos:pos + n].reshape(s))
            pos += n
    assert pos <= z.size
    return out, pos


@pytest.mark.parametrize("case", ["small", "k10", "clip", "k38"])
def test_softmax_model(case):
    g = load_golden("softmax_model.npz")[case]
    par = {"weights": g["W"], "bias": g["b"]}
    Y = O.one_hot(g["y"], g["W"].shape[1])
    got = O.softmax_grad(par, g["X"], Y, g["alpha"])
    np.testing.assert_allclose(got["weights"], g["gW"], **TOL)
    np.testing.assert_allclose(got["bias"], g["gb"], **TOL)
    np.testing.assert_allclose(O.softmax_log_likelihood(par, g["X"], Y), g["ll"], **TOL)
    np.testing.assert_allclose(O.softmax_nlp(par, g["X"], Y, g["alpha"]), g["nlp"], **TOL)
    np.testing.assert_allclose(O.softmax_probs(O.softmax_logits(par, g["X"])), g["probs"], **TOL)
    if case == "clip":
        assert (g["X"] @ g["W"] + g["b"]).max() > O.CLIP_HI  # the clip is really exercised


@pytest.mark.parametrize("case", ["L0", "L1", "L2", "L5", "L20", "blowup"])
def test_hmc_step(case):
    G = load_golden("hmc_step_softmax.npz")
    g = G[case]
    eps, path = g.get("eps", G["eps"]), g.get("path", G["path"])
    d, k = g["W0"].shape
    shapes = [(d, k), (k,)]
    normals, used = split_tape(g["z"], shapes, 1)
    assert used == g["z"].size
    draws = O.TapeDraws(normals, list(np.atleast_1d(g["u"])))
    model = O.SoftmaxOracle({"alpha": g.get("alpha", G["alpha"])})
    with np.errstate(all="ignore"):
        r = O.hmc_step(model, {"weights": g["W0"], "bias": g["b0"]}, ["weights", "bias"], eps, path,
                       draws, X_train=G["X"], y_train=O.one_hot(G["y"], k))
    assert draws.u_pos == np.atleast_1d(g["u"]).size  # same number of global uniforms consumed
    np.testing.assert_allclose(r["accept_prob"], g["accept_prob"], **TOL)
    if case == "blowup":
        # literal reference behaviour: builtin min(1, nan) == 1 -> the NaN proposal is accepted
        assert r["accept_prob"] == 1 and r["accepted"] and np.isnan(g["qW"]).all()
    np.testing.assert_allclose(r["q"]["weights"], g["qW"], **TOL)
    np.testing.assert_allclose(r["q"]["bias"], g["qb"], **TOL)
    np.testing.assert_allclose(r["p"]["weights"], g["pW"], **TOL)
    np.testing.assert_allclose(r["p"]["bias"], g["pb"], **TOL)


def test_hmc_step_covers_accept_and_reject():
    G = load_golden("hmc_step_softmax.npz")
    moved = [not np.array_equal(G[c]["qW"], G[c]["W0"]) for c in ["L2", "L5", "L20"]]
    assert any(moved) and not all(moved)


def test_hmc_sample_softmax():
    g = load_golden("hmc_sample_softmax.npz")
    d, k = g["W0"].shape
    shapes = [(d, k), (k,)]
    normals, used = split_tape(g["z"], shapes, 1 + g["niter"] + g["burnin"])
    assert used == g["z"].size
    draws = O.TapeDraws(normals, list(g["u"]))
    post, loss, info = O.hmc_sample(O.SoftmaxOracle({"alpha": g["alpha"]}), {"weights": g["W0"], "bias": g["b0"]},
                                    g["eps"], g["path"], g["niter"], g["burnin"], draws,
                                    X_train=g["X"], y_train=O.one_hot(g["y"], k))
    assert draws.u_pos == g["u"].size
    np.testing.assert_allclose(post["weights"], g["postW"], **TOL)
    np.testing.assert_allclose(post["bias"], g["postb"], **TOL)
    np.testing.assert_allclose(loss, g["loss"], **TOL)


def test_sgld_sample_softmax():
    g = load_golden("sgld_sample_softmax.npz")
    d, k = g["postW"].shape[1:]
    nb = len(O.minibatch_windows(g["X"].shape[0], g["batch_size"]))
    normals, used = split_tape(g["z"], [(d, k), (k,)], nb * (g["epochs"] + g["burnin"]))
    assert used == g["z"].size
    draws = O.TapeDraws(normals, [])
    post, logp, info = O.sgmcmc_sample(O.sgld_step, O.SoftmaxOracle({"alpha": g["alpha"]}),
                                       {"weights": np.zeros((d, k)), "bias": np.zeros(k)}, g["eps0"],
                                       g["epochs"], g["burnin"], g["batch_size"], draws,
                                       g["X"], O.one_hot(g["y"], k))
    np.testing.assert_allclose(post["weights"], g["postW"], **TOL)
    np.testing.assert_allclose(post["bias"], g["postb"], **TOL)
    np.testing.assert_allclose(logp, g["logp"], **TOL)


def test_sghmc_step():
    g = load_golden("sghmc_step_softmax.npz")
    d, k = g["W0"].shape
    L = O.path_length_steps(g["u"][0], g["path"], g["eps"])
    normals, used = split_tape(g["z"], [(d, k), (k,)], 1 + (L - 1))
    assert used == g["z"].size
    draws = O.TapeDraws(normals, list(g["u"]))
    r = O.sghmc_step(O.SoftmaxOracle({"alpha": g["alpha"]}), {"weights": g["W0"], "bias": g["b0"]},
                     ["weights", "bias"], g["eps"], g["path"], draws, sign="reference",
                     X_train=g["X"], y_train=O.one_hot(g["y"], k))
    np.testing.assert_allclose(r["accept_prob"], g["accept_prob"], **TOL)
    np.testing.assert_allclose(r["q"]["weights"], g["qW"], **TOL)
    np.testing.assert_allclose(r["p"]["bias"], g["pb"], **TOL)


def test_hmc_sample_mvn():
    g = load_golden("hmc_sample_mvn.npz")
    normals, used = split_tape(g["z"], [(2,)], 1 + g["niter"] + g["burnin"])
    assert used == g["z"].size
    draws = O.TapeDraws(normals, list(g["u"]))
    post, loss, info = O.hmc_sample(O.MvnGaussianOracle({"mu": g["mu"], "cov": g["cov"]}), {"x": g["x0"]},
                                    g["eps"], g["path"], g["niter"], g["burnin"], draws)
    assert draws.u_pos == g["u"].size
    np.testing.assert_allclose(post["x"], g["post"], **TOL)
    np.testing.assert_allclose(loss, g["loss"], **TOL)


def test_sgd_fit():
    g = load_golden("sgd_fit_softmax.npz")
    d, k = g["W"].shape
    par, loss = O.sgd_fit(O.SoftmaxOracle({"alpha": g["alpha"]}), {"weights": np.zeros((d, k)), "bias": np.zeros(k)},
                          g["eps"], g["epochs"], g["batch_size"], g["gamma"], g["X"], O.one_hot(g["y"], k))
    np.testing.assert_allclose(par["weights"], g["W"], **TOL)
    np.testing.assert_allclose(par["bias"], g["b"], **TOL)
    np.testing.assert_allclose(loss, g["loss"], **TOL)


def test_mlp_grad_matches_autograd():
    """The MLP reference needs Chainer (absent): its restatement is checked against torch.autograd
    of the same forward written with torch ops (parity unpinned)."""
    import torch
    rs = np.random.RandomState(0)
    B, n_in, n_mid, n_out, alpha = 12, 9, 7, 4, 0.01
    X = rs.rand(B, n_in)
    y = rs.randint(0, n_out, B)
    par = {"/l1/W": rs.normal(0, .3, (n_mid, n_in)), "/l1/b": rs.normal(0, .3, n_mid),
           "/l2/W": rs.normal(0, .3, (n_mid, n_mid)), "/l2/b": rs.normal(0, .3, n_mid),
           "/l3/W": rs.normal(0, .3, (n_out, n_mid)), "/l3/b": rs.normal(0, .3, n_out)}
    masks = [(rs.rand(B, n_mid) > 0.1).astype(float) for _ in range(3)]
    g = O.mlp_grad(par, X, y, masks, alpha)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in par.items()}
    tm = [torch.tensor(m) / 0.9 for m in masks]
    h = torch.relu((torch.tensor(X) @ tp["/l1/W"].T + tp["/l1/b"]) * tm[0])
    h = torch.relu((h @ tp["/l2/W"].T + tp["/l2/b"]) * tm[1])
    logits = (h * tm[2]) @ tp["/l3/W"].T + tp["/l3/b"]
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor(y))
    loss.backward()
    np.testing.assert_allclose(O.mlp_loss(par, X, y, masks), loss.item(), rtol=1e-12)
    for k in par:
        np.testing.assert_allclose(g[k], tp[k].grad.numpy() + 0.5 * alpha * par[k], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("case", ["tiny", "tc", "ragged"])
def test_mlp_model_golden(case):
    """tests/golden/mlp_model.npz: outputs of the unmodified models/gpu/mlp.py run under oracle/chainer_shim.py
    (``python -m oracle.make_golden mlp``) with the three dropout keep-masks recorded."""
    g = load_golden("mlp_model.npz")[case]
    par = {k: g["par" + k] for k in O.MLP_KEYS}
    masks = [g["masks"][i].astype(np.float64) for i in range(3)]
    alpha = float(g["alpha"])
    ref = O.mlp_grad(par, g["X"], g["y"], masks, alpha)
    for k in O.MLP_KEYS:
        np.testing.assert_allclose(ref[k], g["grad" + k], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(O.mlp_loss(par, g["X"], g["y"], masks), g["loss"], rtol=1e-12)
    np.testing.assert_allclose(O.mlp_nlp(par, g["X"], g["y"], masks, alpha), g["nlp"], rtol=1e-12)
    np.testing.assert_allclose(O.mlp_log_prior(par, alpha), g["log_prior"], rtol=1e-12)
    np.testing.assert_array_equal(O.mlp_forward(par, g["X"], masks)[0].argmax(axis=1), g["pred"])


# ---- logistic model, hmc on it, sgd.fit_dropout ---------------------------------
@pytest.mark.parametrize("case", ["small", "d100", "clip"])
def test_logistic_model(case):
    g = load_golden("logistic_model.npz")[case]
    par = {"weights": g["W"], "bias": g["b"]}
    with np.errstate(all="ignore"):
        got = O.logistic_grad(par, g["X"], g["y"], g["alpha"])
        np.testing.assert_allclose(got["weights"], g["gW"], **TOL)
        np.testing.assert_allclose(got["bias"], g["gb"], **TOL)
        np.testing.assert_allclose(O.logistic_log_likelihood(par, g["X"], g["y"]), g["ll"], **TOL)
        np.testing.assert_allclose(O.logistic_nlp(par, g["X"], g["y"], g["alpha"]), g["nlp"], **TOL)
    np.testing.assert_allclose(O.logistic_log_prior(par, g["alpha"]), g["log_prior"], **TOL)
    np.testing.assert_allclose(O.logistic_net(par, g["X"]), g["yhat"], **TOL)
    np.testing.assert_array_equal(O.LogisticOracle({"alpha": g["alpha"]}).predict(par, g["X"], batchsize=32), g["pred"])


def test_hmc_sample_logistic():
    g = load_golden("hmc_sample_logistic.npz")
    d = g["W0"].shape[0]
    normals, used = split_tape(g["z"], [(d, 1), (1,)], 1 + g["niter"] + g["burnin"])
    assert used == g["z"].size
    draws = O.TapeDraws(normals, list(g["u"]))
    post, loss, info = O.hmc_sample(O.LogisticOracle({"alpha": g["alpha"]}), {"weights": g["W0"], "bias": g["b0"]},
                                    g["eps"], g["path"], g["niter"], g["burnin"], draws, X_train=g["X"], y_train=g["y"])
    assert draws.u_pos == g["u"].size
    np.testing.assert_allclose(post["weights"], g["postW"], **TOL)
    np.testing.assert_allclose(post["bias"], g["postb"], **TOL)
    np.testing.assert_allclose(loss, g["loss"], **TOL)
    moved = np.any(np.diff(g["postW"], axis=0) != 0, axis=(1, 2))
    assert moved.any() and not moved.all()  # accepts and rejects both occur


def test_sgd_fit_dropout():
    g = load_golden("sgd_fit_dropout_softmax.npz")
    d, k = g["W"].shape
    par, loss = O.sgd_fit_dropout(O.SoftmaxOracle({"alpha": g["alpha"]}), {"weights": np.zeros((d, k)), "bias": np.zeros(k)},
                                  g["eps"], g["epochs"], g["batch_size"], g["gamma"], g["X"], O.one_hot(g["y"], k),
                                  g["masks"].astype(np.float64))
    np.testing.assert_allclose(par["weights"], g["W"], **TOL)
    np.testing.assert_allclose(par["bias"], g["b"], **TOL)
    np.testing.assert_allclose(loss, g["loss"], **TOL)
    assert 0.6 < g["masks"].mean() < 0.8
