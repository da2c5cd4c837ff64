"""GPU: dropout-MLP model (hamiltonian/models/gpu/mlp.py) against the NumPy restatement in the oracle
(pinned to the unmodified reference file executed under oracle/chainer_shim.py -- Chainer itself is not installable here --
and cross-checked against torch.autograd on CPU; golden fixture tests/golden/mlp_model.npz)."""
import numpy as np
import pytest
import torch

from conftest import TapeRng, replay_uniforms
from oracle import hamiltonian_oracle as O

pytestmark = pytest.mark.gpu

from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.sghmc import sghmc  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.sgd import sgd  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.mlp import KEYS, mlp  # noqa: E402


def make(rs, B, n_in, n_mid, n_out, C=None, scale=0.3):
    lead = () if C is None else (C,)
    par = {"/l1/W": rs.normal(0, scale, lead + (n_mid, n_in)), "/l1/b": rs.normal(0, scale, lead + (n_mid,)),
           "/l2/W": rs.normal(0, scale, lead + (n_mid, n_mid)), "/l2/b": rs.normal(0, scale, lead + (n_mid,)),
           "/l3/W": rs.normal(0, scale, lead + (n_out, n_mid)), "/l3/b": rs.normal(0, scale, lead + (n_out,))}
    X = rs.rand(B, n_in)
    y = rs.randint(0, n_out, B)
    return par, X, y


def close(got, ref, rtol=1e-4, atol_scale=1e-5, what=""):
    ref = np.asarray(ref, dtype=np.float64)
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), ref, rtol=rtol,
                               atol=atol_scale * max(1e-30, float(np.abs(ref).max())), err_msg=what)


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("shape", [(37, 20, 24, 5), (130, 70, 96, 10), (500, 784, 512, 10)])
def test_mlp_grad_loss_nlp_injected_masks(shape, prec):
    """(500, 784, 512, 10) is BASELINE config 4's layer shape; the five large GEMMs run on tcgen05 for bf16x3."""
    B, n_in, n_mid, n_out = shape
    rs = np.random.RandomState(0)
    C, alpha = (2 if B == 500 else 3), 0.05
    par, X, y = make(rs, B, n_in, n_mid, n_out, C, scale=(0.05 if B == 500 else 0.3))
    masks = (rs.rand(3, C, B, n_mid) > 0.1).astype(np.uint8)
    m = mlp({"alpha": alpha}, n_in, n_mid, n_out, precision=prec)
    m.bind(X, y)
    m.set_masks(masks)
    g = m.grad(par, X_train=X, y_train=y)
    loss = m.log_likelihood(par, X_train=X, y_train=y)
    nlp = m.negative_log_posterior(par, X_train=X, y_train=y)
    for c in range(C):
        pc = {k: par[k][c] for k in KEYS}
        mk = [masks[l, c].astype(np.float64) for l in range(3)]
        ref = O.mlp_grad(pc, X, y, mk, alpha)
        for k in KEYS:
            err = np.abs(g[k][c] - ref[k]).max() / np.abs(ref[k]).max()
            print("%s %s chain %d %-6s max err / max|g| = %.2e" % (shape, prec, c, k, err))
        for k in KEYS:
            if prec == "fp32":
                close(g[k][c], ref[k], what="%s chain %d" % (k, c))
            else:
                # ReLU kinks: a pre-activation within ~1e-5 of zero can fall on the other side of the gate in
                # split-bf16 arithmetic, which changes that sample's contribution to one gradient row.  Require
                # rtol 1e-4 for >= 99.5 % of the entries and cap every entry at 1e-3 of the largest one.
                r, gk = np.asarray(ref[k], dtype=np.float64), np.asarray(g[k][c], dtype=np.float64)
                bad = np.abs(gk - r) > 1e-4 * np.abs(r) + 1e-5 * np.abs(r).max()
                assert bad.mean() <= 5e-3, (k, c, bad.mean())
                assert np.abs(gk - r).max() <= 1e-3 * np.abs(r).max(), (k, c)
        close(loss[c], O.mlp_loss(pc, X, y, mk), 1e-5)
        close(nlp[c], O.mlp_nlp(pc, X, y, mk, alpha), 1e-5)


def test_mlp_no_dropout_and_philox_dropout_statistics():
    rs = np.random.RandomState(1)
    B, n_in, n_mid, n_out, alpha = 64, 12, 512, 4, 0.0
    par, X, y = make(rs, B, n_in, n_mid, n_out, scale=0.2)
    m0 = mlp({"alpha": alpha}, n_in, n_mid, n_out, dropout=0.0, precision="bf16x3")
    close(m0.log_likelihood(par, X_train=X, y_train=y), O.mlp_loss(par, X, y, None), 1e-5)
    g0 = m0.grad(par, X_train=X, y_train=y)
    ref = O.mlp_grad(par, X, y, None, alpha)
    close(g0["/l1/W"], ref["/l1/W"])
    # Philox dropout: a fresh mask per evaluation (like Chainer), loss fluctuates around the no-dropout value
    m1 = mlp({"alpha": alpha}, n_in, n_mid, n_out, dropout=0.1, seed=7)
    ls = np.array([m1.log_likelihood(par, X_train=X, y_train=y) for _ in range(30)])
    assert len(np.unique(ls)) == 30
    # keep rate: with l1 bias large and positive, relu(dropout(a)) == 0 exactly where the unit was dropped
    par2 = {k: np.array(v, copy=True) for k, v in par.items()}
    par2["/l1/W"][:] = 0
    par2["/l1/b"][:] = 1.0
    par2["/l2/W"][:] = np.eye(n_mid)
    par2["/l2/b"][:] = 0
    g = m1.grad(par2, X_train=X, y_train=y)  # exercises the path; statistics via the loss spread above
    assert np.isfinite(g["/l2/W"]).all()
    assert abs(ls.mean() - O.mlp_loss(par, X, y, None)) < 0.5


def test_sghmc_step_on_mlp_minibatch_vs_oracle():
    rs = np.random.RandomState(3)
    B, n_in, n_mid, n_out, alpha = 48, 10, 16, 3, 0.1
    par, X, y = make(rs, B, n_in, n_mid, n_out, scale=0.2)
    masks = (rs.rand(3, 1, B, n_mid) > 0.1).astype(np.uint8)
    eps, path = 1e-2, 3e-2
    model = mlp({"alpha": alpha}, n_in, n_mid, n_out)
    model.bind(X, y)
    model.set_masks(masks)
    s = sghmc(model, par, path_length=path, step_size=eps, verbose=False, sign="descent")
    P = sum(int(np.prod(v.shape)) for v in par.values())
    L = O.path_length_steps(0.45, path, eps)
    z = rs.normal(size=P * (1 + (L - 1)))
    with replay_uniforms([0.45, 0.5]):
        q, p, a = s.step(par, None, TapeRng(z), X_train=X, y_train=y)
    shapes = [par[k].shape for k in KEYS]
    normals, pos = [], 0
    for _ in range(1 + (L - 1)):
        for sh in shapes:
            n = int(np.prod(sh))
            normals.append(z[pos:pos + n].reshape(sh))
            pos += n
    mk = [masks[l, 0].astype(np.float64) for l in range(3)]
    r = O.sghmc_step(O.MlpOracle({"alpha": alpha}, lambda b, n: mk), par, list(KEYS), eps, path,
                     O.TapeDraws(normals, [0.45, 0.5]), sign="descent", X_train=X, y_train=y)
    assert r["L"] == L and L >= 3
    close(a, r["accept_prob"], 1e-3, 1e-6)
    for k in KEYS:
        close(q[k], r["q"][k], 1e-4, 1e-5, k)
        close(p[k], r["p"][k], 1e-4, 1e-5, k)


def test_sgd_fit_on_mlp_reduces_loss():
    rs = np.random.RandomState(4)
    B, n_in, n_mid, n_out = 256, 20, 32, 4
    par, X, _ = make(rs, B, n_in, n_mid, n_out, scale=0.1)
    y = (X[:, :n_out] + 0.1 * rs.rand(B, n_out)).argmax(1)
    model = mlp({"alpha": 1e-4}, n_in, n_mid, n_out, dropout=0.1, seed=1)
    opt = sgd(model, par, step_size=0.05)
    fitted, loss = opt.fit(epochs=30, batch_size=64, gamma=0.9, X_train=X, y_train=y)
    assert loss[-1] < 0.7 * loss[0]


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_mlp_predict(prec):
    """mlp.predict (mlp.py:84-95): with dropout=0 the forward pass is deterministic and must match the oracle's
    forward; with dropout on (reference behaviour) the probabilities are a valid distribution and differ between
    chains only through their parameters and masks."""
    rs = np.random.RandomState(3)
    B, n_in, n_mid, n_out, C = 77, 40, 64, 6, 3
    par, X, y = make(rs, B, n_in, n_mid, n_out, C)
    m0 = mlp({"alpha": 0.05}, n_in, n_mid, n_out, precision=prec, dropout=0.0)
    probs = m0.predict(par, X, prob=True)
    labels = m0.predict(par, X)
    assert probs.shape == (C, B, n_out) and labels.shape == (C, B)
    for c in range(C):
        pc = {k: par[k][c] for k in KEYS}
        z = O.mlp_forward(pc, X, None)[0]
        e = np.exp(z - z.max(axis=1, keepdims=True))
        ref = e / e.sum(axis=1, keepdims=True)
        close(probs[c], ref, rtol=2e-4, atol_scale=2e-5, what="chain %d" % c)
        assert (labels[c] == ref.argmax(axis=1)).mean() > 0.97
    one = m0.predict({k: par[k][0] for k in KEYS}, X, prob=True)
    assert one.shape == (B, n_out)
    close(one, probs[0], rtol=1e-6, atol_scale=1e-7)
    m1 = mlp({"alpha": 0.05}, n_in, n_mid, n_out, precision=prec, dropout=0.1, seed=7)
    p1 = m1.predict(par, X, prob=True)
    np.testing.assert_allclose(p1.sum(axis=2), 1.0, rtol=1e-5)
    assert np.abs(p1 - probs).max() > 1e-3  # dropout really is on


def test_mlp_producer_written_operand_copies_are_bit_identical():
    """csrc/mlp.cu (round 2): the GEMM epilogues write the bf16 hi/lo operand copies their consumers need
    (GemmDesc::ck_* / ct_*) instead of separate split launches.  The copies are the split of the same fp32 values, so
    the gradient must be BIT-identical to the one of the split-launch path (BHMC_MLP_FUSE=0), for the hi/lo scheme and
    the single pass; every variant runs in its own interpreter (the switch is read once per process)."""
    import os
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
from dropout_hamiltonian_montecarlo_b200.runtime import MlpHandle, default_context
ctx = default_context()
rs = np.random.RandomState(4)
out = {}
for name, (N, n_in, n_mid, n_out, C, row0, B) in {"cfg4": (1200, 784, 512, 10, 5, 500, 500), "odd": (700, 200, 192, 7, 3, 64, 333)}.items():
    X = torch.as_tensor(rs.rand(N, n_in).astype(np.float32)).cuda(); y = torch.as_tensor(rs.randint(0, n_out, N).astype(np.int32)).cuda()
    h = MlpHandle(ctx, N, n_in, n_mid, n_out, 0.01, 0.1, seed=3, chain_id0=2); h.bind(X, y)
    q = h.pack(rs.normal(0, .05, (C, h.P)).astype(np.float32))
    for prec in (1, 2):
        g, st = h.grad(q, row0, B, prec)
        ctx.sync()
        out["%%s_g%%d" %% (name, prec)] = g[:, :h.P].cpu().numpy(); out["%%s_s%%d" %% (name, prec)] = st.cpu().numpy()
        l0 = ctx.launches
        h.grad(q, row0, B, prec); ctx.sync()
        out["%%s_l%%d" %% (name, prec)] = ctx.launches - l0
    h.close()
np.savez(sys.argv[1], **out)
''' % (root,)
    res = []
    for fuse, head in (("1", "0"), ("0", "0"), ("1", "1")):
        with tempfile.NamedTemporaryFile(suffix=".npz") as f:
            r = subprocess.run([sys.executable, "-c", code, f.name], env=dict(os.environ, BHMC_MLP_FUSE=fuse, BHMC_MLP_HEAD=head),
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-3000:]
            res.append({k: v for k, v in np.load(f.name).items()})
    a, b, h = res
    # the one-pass head (k_mlp_head, BHMC_MLP_HEAD=1, the default) sums the logits and the gW3 / gb3 / gb2 terms in another
    # order than the five launches it replaces: equal to fp32 round-off, 7 launches become 2
    for name in ("cfg4", "odd"):
        for prec in (1, 2):
            ga, gh = a["%s_g%d" % (name, prec)].astype(np.float64), h["%s_g%d" % (name, prec)].astype(np.float64)
            assert np.abs(gh - ga).max() <= 2e-5 * np.abs(ga).max(), (name, prec, np.abs(gh - ga).max() / np.abs(ga).max())
            np.testing.assert_allclose(h["%s_s%d" % (name, prec)], a["%s_s%d" % (name, prec)], rtol=2e-6)
            assert int(h["%s_l%d" % (name, prec)]) == int(a["%s_l%d" % (name, prec)]) - 5  # 7 launches of the head become 2
    for name in ("cfg4", "odd"):
        for prec in (1, 2):
            # Philox dropout masks are keyed by an evaluation counter, identical in both processes
            np.testing.assert_array_equal(a["%s_g%d" % (name, prec)], b["%s_g%d" % (name, prec)])
            np.testing.assert_array_equal(a["%s_s%d" % (name, prec)], b["%s_s%d" % (name, prec)])
            # 4 split launches replaced by the producing epilogues, 2 more by the cached copies of the X window
            assert int(a["%s_l%d" % (name, prec)]) == int(b["%s_l%d" % (name, prec)]) - 6, (name, prec, a["%s_l%d" % (name, prec)], b["%s_l%d" % (name, prec)])


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("case", ["tiny", "tc", "ragged"])
def test_mlp_model_golden_through_cuda(case, prec):
    """tests/golden/mlp_model.npz holds what the UNMODIFIED reference file models/gpu/mlp.py returned when executed under
    oracle/chainer_shim.py (fake chainer / cupy, ``python -m oracle.make_golden mlp``), with its three dropout keep-masks
    recorded: grad (mlp.py:47-64), log_likelihood (:66-78), negative_log_posterior (:80-82).  Tolerance rtol 1e-4 (+ 1e-5
    of the largest entry for entries near zero); 'tc' has every GEMM dimension >= 64 so the tcgen05 path runs."""
    from conftest import load_golden
    g = load_golden("mlp_model.npz")[case]
    par = {k: g["par" + k] for k in KEYS}
    X, y, alpha = g["X"], g["y"], float(g["alpha"])
    n_mid, n_in = par["/l1/W"].shape
    n_out = par["/l3/W"].shape[0]
    m = mlp({"alpha": alpha}, n_in, n_mid, n_out, precision=prec)
    m.bind(X, y)
    m.set_masks(g["masks"][:, None])  # [3, C = 1, B, n_mid]
    got = m.grad(par, X_train=X, y_train=y)
    for k in KEYS:
        ref = np.asarray(g["grad" + k], dtype=np.float64)
        gk = np.asarray(got[k], dtype=np.float64).reshape(ref.shape)
        if prec == "fp32":
            close(gk, ref, what=k)
        else:  # ReLU kinks, see test_mlp_grad_loss_nlp_injected_masks
            bad = np.abs(gk - ref) > 1e-4 * np.abs(ref) + 1e-5 * np.abs(ref).max()
            assert bad.mean() <= 5e-3 and np.abs(gk - ref).max() <= 1e-3 * np.abs(ref).max(), (k, bad.mean())
    close(m.log_likelihood(par, X_train=X, y_train=y), g["loss"], 1e-5)
    close(m.negative_log_posterior(par, X_train=X, y_train=y), g["nlp"], 1e-5)


class _FlatMlpOracle:
    """The oracle's MLP seen as ONE variable 'theta' (KEYS order, the layout of the parameter vector on the device): the
    joint sweep -- one gradient per leapfrog iteration, every parameter moved together -- is Gauss-Seidel over that
    single variable (oracle sghmc_step / hmc_step unchanged)."""

    def __init__(self, inner, shapes):
        self.inner, self.shapes = inner, shapes

    def unflat(self, par):
        th, out, pos = np.asarray(par["theta"]), {}, 0
        for k, sh in zip(KEYS, self.shapes):
            n = int(np.prod(sh))
            out[k] = th[pos:pos + n].reshape(sh)
            pos += n
        return out

    def grad(self, par, **d):
        g = self.inner.grad(self.unflat(par), **d)
        return {"theta": np.concatenate([np.asarray(g[k]).ravel() for k in KEYS])}

    def log_likelihood(self, par, **d):
        return self.inner.log_likelihood(self.unflat(par), **d)

    def negative_log_posterior(self, par, **d):
        return self.inner.negative_log_posterior(self.unflat(par), **d)


def test_sghmc_joint_sweep_on_mlp_vs_oracle():
    """BASELINE config 4 runs the 'joint' sweep (one gradient per leapfrog iteration).  Shape with every GEMM dimension
    >= 64: the tcgen05 path with the update applied by the gradient epilogues (ModelBase::grad_fused_update), the operand
    mirror and the W2^T copy of csrc/mlp.cu all take part; injected noise and dropout masks; fp64 oracle."""
    rs = np.random.RandomState(8)
    B, n_in, n_mid, n_out, alpha = 80, 64, 64, 10, 0.1
    par, X, y = make(rs, B, n_in, n_mid, n_out, scale=0.1)
    masks = (rs.rand(3, 1, B, n_mid) > 0.1).astype(np.uint8)
    eps, path = 1e-2, 3.5e-2
    model = mlp({"alpha": alpha}, n_in, n_mid, n_out, precision="bf16x3")
    model.bind(X, y)
    model.set_masks(masks)
    s = sghmc(model, par, path_length=path, step_size=eps, verbose=False, sign="descent", sweep="joint")
    shapes = [par[k].shape for k in KEYS]
    P = sum(int(np.prod(sh)) for sh in shapes)
    L = O.path_length_steps(0.62, path, eps)
    z = rs.normal(size=P * L)
    with replay_uniforms([0.62, 0.5]):
        q, p, a = s.step(par, None, TapeRng(z), X_train=X, y_train=y)
    flat = _FlatMlpOracle(O.MlpOracle({"alpha": alpha}, lambda b, n: [masks[l, 0].astype(np.float64) for l in range(3)]), shapes)
    th0 = {"theta": np.concatenate([np.asarray(par[k], dtype=np.float64).ravel() for k in KEYS])}
    r = O.sghmc_step(flat, th0, ["theta"], eps, path, O.TapeDraws([z[i * P:(i + 1) * P] for i in range(L)], [0.62, 0.5]),
                     sign="descent", X_train=X, y_train=y)
    assert r["L"] == L and L >= 4
    rq, rp = flat.unflat(r["q"]), flat.unflat(r["p"])
    close(a, r["accept_prob"], 1e-3, 1e-6)
    for k in KEYS:
        close(q[k], rq[k], 1e-4, 2e-5, k)
        close(p[k], rp[k], 1e-4, 2e-5, k)


def test_mlp_update_applied_by_the_gradient_epilogues_matches_update_launches():
    """csrc/mlp.cu grad_fused_update (default) against BHMC_MLP_FUSE_UPD=0 (gradient to HBM + k_hmc_update launches):
    same Philox noise and masks, several chains, three sghmc steps on different row windows -- positions and momenta equal to
    fp32 round-off, the same number of gradient evaluations, fewer launches."""
    import os
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
from dropout_hamiltonian_montecarlo_b200.runtime import MlpHandle, SamplerHandle, default_context
ctx = default_context()
rs = np.random.RandomState(4)
N, n_in, n_mid, n_out, C, B = 600, 128, 192, 7, 5, 200
X = torch.as_tensor(rs.rand(N, n_in).astype(np.float32)).cuda(); y = torch.as_tensor(rs.randint(0, n_out, N).astype(np.int32)).cuda()
out = {}
for kind, name in ((2, "sghmc"), (0, "hmc")):
    h = MlpHandle(ctx, N, n_in, n_mid, n_out, 0.01, 0.1, seed=3, chain_id0=2); h.bind(X, y)
    s = SamplerHandle(ctx, h, kind, C, seed=11, precision=1, sweep=[(0, h.P)], shared_path=(kind == 2), sghmc_descent=True)
    s.set_q(np.random.RandomState(5).normal(0, .05, (C, h.P)).astype(np.float32))
    l0 = ctx.launches; ng = 0
    for j in range(3):
        o = s.hmc_run(1, 2e-3, 1.2e-2, row0=j * B, nrows=B, step0=j, keep_samples=False, schedule="lockstep")
        ng += o["n_grad_evals"]
    ctx.sync()
    out[name + "_q"] = s.get(0); out[name + "_p"] = s.get(1); out[name + "_ng"] = ng; out[name + "_l"] = ctx.launches - l0
    s.close(); h.close()
np.savez(sys.argv[1], **out)
''' % (root,)
    res = []
    for fu in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".npz") as f:
            r = subprocess.run([sys.executable, "-c", code, f.name], env=dict(os.environ, BHMC_MLP_FUSE_UPD=fu), capture_output=True,
                               text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-3000:]
            res.append({k: v for k, v in np.load(f.name).items()})
    a, b = res
    for name in ("sghmc", "hmc"):
        assert int(a[name + "_ng"]) == int(b[name + "_ng"]) and int(a[name + "_ng"]) > 3 * 5
        assert int(a[name + "_l"]) < int(b[name + "_l"]), (name, a[name + "_l"], b[name + "_l"])
        for w in ("_q", "_p"):
            x, yv = a[name + w].astype(np.float64), b[name + w].astype(np.float64)
            assert np.abs(x - yv).max() <= 2e-5 * np.abs(yv).max(), (name, w, np.abs(x - yv).max() / np.abs(yv).max())
