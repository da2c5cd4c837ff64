"""Effective sample size (the reference computes none; BASELINE's metric asks for ESS/s).

Geyer's initial-positive-sequence estimator per chain and parameter: tau = -1 + 2 * sum of the
consecutive autocorrelation pairs rho[2k] + rho[2k+1] while they stay positive; ESS = n / tau, summed
over chains.  Host-side analysis of the returned samples (NumPy) -- not part of the GPU hot path."""
import numpy as np


def autocorr(x):
    """Autocorrelation of a 1-D series via FFT (biased normalisation), rho[0] == 1."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    x = x - x.mean()
    var = np.dot(x, x) / n
    if var <= 0:
        return np.zeros(n)
    f = np.fft.rfft(x, 2 * n)
    acov = np.fft.irfft(f * np.conj(f))[:n] / n
    return acov / var


def ess_1d(x):
    """Geyer IPS effective sample size of one chain of one scalar."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 4 or np.allclose(x, x[0]):
        return 0.0
    rho = autocorr(x)
    tau = -1.0
    for k in range(0, n - 1, 2):
        pair = rho[k] + rho[k + 1]
        if pair <= 0:
            break
        tau += 2.0 * pair
    return float(n / max(tau, 1e-12)) if tau > 0 else float(n)


def ess(samples, max_params=64, seed=0):
    """samples: [n_steps, n_chains, P] (or [n_steps, P]).  Returns dict(min, median, per_param) with ESS summed
    over chains, evaluated on at most ``max_params`` randomly chosen parameters."""
    s = np.asarray(samples)
    if s.ndim == 2:
        s = s[:, None, :]
    n, C, P = s.shape
    idx = np.arange(P) if P <= max_params else np.random.RandomState(seed).choice(P, max_params, replace=False)
    per = np.array([sum(ess_1d(s[:, c, j]) for c in range(C)) for j in idx])
    return {"min": float(per.min()), "median": float(np.median(per)), "per_param": per, "params": idx}
