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
