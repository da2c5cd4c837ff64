"""GPU (B200): edge shapes and error behaviour of the model entry points, through the C ABI.
The reference has no error convention beyond Python exceptions; the ABI returns negative codes that the
Python mirror raises as BhmcError -- never a silent fallback."""
import numpy as np
import pytest
import torch

from oracle import hamiltonian_oracle as O

pytestmark = pytest.mark.gpu

from dropout_hamiltonian_montecarlo_b200._lib import BhmcError  # noqa: E402
from dropout_hamiltonian_montecarlo_b200.runtime import SoftmaxHandle, default_context  # noqa: E402


def _case(N, D, K, C, seed):
    rs = np.random.RandomState(seed)
    X = rs.rand(N, D).astype(np.float32)
    y = rs.randint(0, K, N).astype(np.int32)
    q = rs.normal(0, .2, (C, (D + 1) * K)).astype(np.float32)
    return X, y, q


def _check(h, X, y, q, r0, n, prec, alpha, rtol, scale):
    D, K = X.shape[1], h.K
    g, ll = h.grad(h.pack(q), r0, n, prec)
    g = g[:, :h.P].cpu().numpy().astype(np.float64)
    ll = ll.cpu().numpy()
    for c in range(q.shape[0]):
        par = {"weights": q[c, :D * K].reshape(D, K).astype(np.float64), "bias": q[c, D * K:].astype(np.float64)}
        Xs, Ys = X[r0:r0 + n].astype(np.float64), O.one_hot(y[r0:r0 + n], K)
        ref = O.flatten_par(O.softmax_grad(par, Xs, Ys, alpha), ["weights", "bias"])
        atol = scale * max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(g[c], ref, rtol=rtol, atol=atol, err_msg="N=%d D=%d K=%d prec=%d" % (n, D, K, prec))
        np.testing.assert_allclose(ll[c], O.softmax_log_likelihood(par, Xs, Ys), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("N,D,K,C", [(1, 5, 3, 1), (130, 3, 2, 3), (257, 70, 64, 2), (64, 784, 10, 1), (1000, 9, 38, 5),
                                     (129, 65, 11, 7)])
def test_edge_shapes(N, D, K, C):
    """one row, tiny feature counts, the largest class count of the tensor path, class counts that need padding,
    row counts one past a tile"""
    X, y, q = _case(N, D, K, C, 3)
    h = SoftmaxHandle(default_context(), N, D, K, 0.01)
    h.bind(torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda())
    _check(h, X, y, q, 0, N, 0, 0.01, 2e-5, 2e-6)
    _check(h, X, y, q, 0, N, 1, 0.01, 1e-4, 2e-5)
    if N > 2:
        _check(h, X, y, q, 1, N - 2, 1, 0.01, 1e-4, 2e-5)


def test_more_classes_than_the_tensor_path_supports():
    N, D, K = 200, 12, 65
    X, y, q = _case(N, D, K, 2, 4)
    ctx = default_context()
    h = SoftmaxHandle(ctx, N, D, K, 0.01)
    with pytest.raises(BhmcError):
        h.bind(torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda(), 0b111)
    h2 = SoftmaxHandle(ctx, N, D, K, 0.01)
    h2.bind(torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda(), 0b001)  # the fp32 path has no such limit
    _check(h2, X, y, q, 0, N, 0, 0.01, 2e-5, 2e-6)
    with pytest.raises(BhmcError):  # ... and a tensor precision that was not prepared is an error, not a fallback
        h2.grad(h2.pack(q), 0, N, 1)


def test_row_window_outside_the_bound_rows():
    N, D, K = 300, 10, 4
    X, y, q = _case(N, D, K, 2, 5)
    h = SoftmaxHandle(default_context(), N, D, K, 0.01)
    h.bind(torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda())
    for prec in (0, 1):
        with pytest.raises(BhmcError):
            h.grad(h.pack(q), 200, 200, prec)
        with pytest.raises(BhmcError):
            h.grad(h.pack(q), -1, 10, prec)
    _check(h, X, y, q, 200, 100, 1, 0.01, 1e-4, 2e-5)  # the handle is still usable after the errors


@pytest.mark.parametrize("sched", ["lockstep", "streaming"])
def test_grad_hook_path_matches_the_fused_path(sched):
    """The row-shard all-reduce hook (a no-op with one rank) forces the unfused launch sequence -- separate split-K
    reduce, update and operand preparation: same draws, same decisions, samples equal to fp32 round-off."""
    from dropout_hamiltonian_montecarlo_b200.parallel import RowShardHook
    from dropout_hamiltonian_montecarlo_b200.runtime import SamplerHandle
    rs = np.random.RandomState(12)
    N, D, K, C, n_steps = 700, 50, 10, 6, 3
    X, y, _ = _case(N, D, K, C, 12)
    q0 = rs.normal(0, .1, (C, (D + 1) * K)).astype(np.float32)
    h = SoftmaxHandle(default_context(), N, D, K, 0.01)
    h.bind(torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda())
    z = torch.as_tensor(rs.normal(size=(n_steps, C, h.P)), dtype=torch.float32)
    u1, u2 = rs.rand(n_steps, C), rs.rand(n_steps, C)
    outs = []
    for with_hook in (False, True):
        s = SamplerHandle(h.ctx, h, 0, C, precision=1)
        hook = RowShardHook(s) if with_hook else None
        s.set_q(q0)
        o = s.hmc_run(n_steps, 5e-4, 4e-3, z_momentum=z, u_path=u1, u_accept=u2, schedule=sched)
        outs.append((o["samples"].cpu().numpy(), o["accept_prob"].cpu().numpy(), o["n_grad_evals"]))
        if with_hook:
            assert hook.calls > 0
    (s0, a0, n0), (s1, a1, n1) = outs
    assert n0 == n1
    np.testing.assert_allclose(a1, a0, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(s1, s0, rtol=1e-5, atol=1e-6)


def test_energies_and_decisions_are_bit_reproducible():
    """Energy terms are accumulated with fp64 atomics in arrival order; every addend is rounded to a multiple of 2^-24
    first (csrc/internal.cuh quantize_addend), so the sums are exact and do not depend on that order: the same call
    twice gives the same bits -- log-likelihoods, acceptance probabilities, samples.  (Round 1: last-bit differences from
    run to run.)  Several blocks per chain on purpose: N = 5000 rows, P = 3030."""
    from dropout_hamiltonian_montecarlo_b200.runtime import SamplerHandle
    rs = np.random.RandomState(21)
    N, D, K, C, n_steps = 5000, 100, 30, 9, 4
    X, y, _ = _case(N, D, K, C, 21)
    q0 = rs.normal(0, .1, (C, (D + 1) * K)).astype(np.float32)
    h = SoftmaxHandle(default_context(), N, D, K, 0.01)
    h.bind(torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda())
    lls = [h.grad(h.pack(q0), 0, N, 1)[1].cpu().numpy().copy() for _ in range(6)]
    for ll in lls[1:]:
        assert np.array_equal(ll.view(np.int64), lls[0].view(np.int64))
    assert np.all(lls[0] * 2.0 ** 24 == np.rint(lls[0] * 2.0 ** 24))  # on the grid: the sum was exact
    outs = []
    for rep in range(3):
        s = SamplerHandle(h.ctx, h, 0, C, precision=1, seed=5)
        s.set_q(q0)
        o = s.hmc_run(n_steps, 2e-4, 2e-3, schedule="streaming")
        outs.append((o["samples"].cpu().numpy().copy(), o["accept_prob"].cpu().numpy().copy()))
    for smp, acc in outs[1:]:
        assert np.array_equal(acc.view(np.int64), outs[0][1].view(np.int64))
        assert np.array_equal(smp.view(np.int32), outs[0][0].view(np.int32))
