"""CPU: the ESS estimator on series with known effective sample size."""
import numpy as np

from dropout_hamiltonian_montecarlo_b200.ess import ess, ess_1d


def ar1(n, phi, rs):
    x = np.zeros(n)
    e = rs.normal(size=n)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    return x


def test_iid_and_ar1():
    rs = np.random.RandomState(0)
    n = 20000
    assert abs(ess_1d(rs.normal(size=n)) / n - 1.0) < 0.1
    for phi in (0.5, 0.9):
        expect = n * (1 - phi) / (1 + phi)
        got = ess_1d(ar1(n, phi, rs))
        assert abs(got / expect - 1.0) < 0.2, (phi, got, expect)
    assert ess_1d(np.ones(100)) == 0.0


def test_ess_sums_over_chains_and_params():
    rs = np.random.RandomState(1)
    s = np.stack([np.stack([ar1(4000, 0.8, rs) for _ in range(3)], axis=1) for _ in range(5)], axis=1)  # [n, 5 chains, 3]
    r = ess(s)
    expect = 5 * 4000 * 0.2 / 1.8
    assert abs(r["median"] / expect - 1) < 0.25 and r["min"] <= r["median"]


def test_sample_sink_roundtrip_cpu(tmp_path):
    """hamiltonian/sink.py is host logic: chunked appends land where backend_mean finds them."""
    import os
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.sink import SampleSink, backend_mean
    rs = np.random.RandomState(0)
    shapes = {"weights": (3, 2), "bias": (2,)}
    flat = rs.normal(size=(10, 4, 8)).astype(np.float32)
    backend = os.path.join(str(tmp_path), "run", "c0")
    sink = SampleSink(backend, shapes, 10, 4, squeeze=False)
    sink.append(flat[:6])
    sink.append(flat[6:])
    files = sink.close()
    w = np.load(files["weights"])
    assert w.shape == (10, 4, 3, 2) and w.dtype == np.float32
    np.testing.assert_array_equal(w.reshape(10, 4, 6), flat[:, :, :6])
    np.testing.assert_array_equal(np.load(files["bias"]), flat[:, :, 6:])
    m = backend_mean([backend], 10)
    np.testing.assert_allclose(m["weights"], flat[:, :, :6].sum(0).reshape(4, 3, 2) / 10, rtol=1e-6)
    one = SampleSink(os.path.join(str(tmp_path), "c1"), shapes, 4, 1, squeeze=True)
    one.append(flat[:4, :1])
    f1 = one.close()
    assert np.load(f1["bias"]).shape == (4, 2)


def test_reference_helper_methods_cpu():
    """The small host-side helpers of the reference classes (softmax.softmax / cross_entropy, logistic.sigmoid,
    hmc.potential_energy / draw_momentum, sgd.iterate_minibatches, sgld.draw_momentum) need no device."""
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.sgd import sgd
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.sgld import sgld
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.logistic import logistic
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
    from oracle import hamiltonian_oracle as O
    rs = np.random.RandomState(0)
    z = rs.normal(size=(7, 4))
    y = O.one_hot(rs.randint(0, 4, 7), 4)
    m = softmax({"alpha": 0.1})
    np.testing.assert_allclose(m.softmax(z), O.softmax_probs(z), rtol=1e-12)
    np.testing.assert_allclose(m.cross_entropy(z, y).sum(), np.sum(y * (z - O._logsumexp_rows(z)[:, None])), rtol=1e-12)
    np.testing.assert_allclose(logistic({"alpha": 0.1}).sigmoid(z), 1 / (1 + np.exp(-z)), rtol=1e-12)
    start = {"weights": np.zeros((3, 2)), "bias": np.zeros(2)}
    h = hmc(m, start, verbose=False)
    p = h.draw_momentum(np.random.RandomState(1))
    r = np.random.RandomState(1)
    np.testing.assert_array_equal(p["weights"], r.normal(0, 1, (3, 2)))
    np.testing.assert_array_equal(p["bias"], r.normal(0, 1, 2))
    np.testing.assert_allclose(h.potential_energy(p), O.kinetic_energy(p), rtol=1e-12)
    X, yy = rs.rand(10, 3), rs.rand(10)
    wins = [xb.shape[0] for xb, _ in sgd(m, start).iterate_minibatches(X, yy, 4)]
    assert wins == [4, 4]  # remainder of 2 rows dropped
    eta = sgld(m, start, step_size=0.01, verbose=False).draw_momentum(np.random.RandomState(2), 0.01)
    np.testing.assert_array_equal(eta["weights"], np.random.RandomState(2).normal(0, 0.02, (3, 2)))
