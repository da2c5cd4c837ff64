"""B200-native (sm_100a) implementation of the sampler hot path of
sherna90/dropout_hamiltonian_montecarlo: batched-chain HMC / SGLD / SGHMC / SGD over the
softmax-regression (and MLP) log-posterior, behind the reference's own API.

    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc

or, to keep reference import paths (``import hamiltonian.inference.gpu.hmc``) working:

    import dropout_hamiltonian_montecarlo_b200 as b200; b200.install_as_hamiltonian()

All compute runs in ``csrc/libbhmc.so`` (hand-written CUDA, C ABI in ``include/bhmc.h``).
There is no CPU fallback.
"""
import sys

__version__ = "0.1.0"


def install_as_hamiltonian():
    """Alias this package's ``hamiltonian`` tree as the top-level ``hamiltonian`` package so
    that scripts written against the reference import the B200 implementation unchanged."""
    import importlib

    root = importlib.import_module(__name__ + ".hamiltonian")
    sys.modules["hamiltonian"] = root
    for sub in ("utils", "models", "models.gpu", "models.gpu.softmax", "models.gpu.mvn_gaussian", "models.gpu.mlp",
                "inference", "inference.gpu", "inference.gpu.hmc", "inference.gpu.sgmcmc",
                "inference.gpu.sgld", "inference.gpu.sghmc", "inference.gpu.sgd"):
        sys.modules["hamiltonian." + sub] = importlib.import_module(__name__ + ".hamiltonian." + sub)
    return root
