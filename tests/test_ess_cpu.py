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
