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
    import pkgutil

    # every module of the tree, discovered rather than listed (a hand-kept list once missed models.gpu.logistic)
    for info in pkgutil.walk_packages(root.__path__, root.__name__ + "."):
        mod = importlib.import_module(info.name)
        sys.modules["hamiltonian." + info.name[len(root.__name__) + 1:]] = mod
    return root
