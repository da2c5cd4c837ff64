"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (``/root/reference``).

The reference package cannot be imported as shipped on Python >= 3.10
(``hamiltonian/utils.py:2`` imports ``collections.Iterable``; every sampler imports
``h5py`` -- ``hamiltonian/inference/cpu/hmc.py:7`` -- which is not installed).  This module
applies the 3-line compatibility shim of SURVEY.md section 8(c) *before* importing the
reference, leaving the read-only tree untouched.

It is used in exactly two places:
  * ``oracle/make_golden.py``  -- mints the committed fixtures under ``tests/golden/``;
  * ``tests/test_oracle_vs_reference.py`` -- pins the numpy restatement
    (``oracle/hamiltonian_oracle.py``) against the unmodified reference.
Both are skipped when ``/root/reference`` is absent (e.g. on the GPU box).  Nothing in the
product package may import this file.
"""
import collections
import collections.abc
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BHMC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hamiltonian"))


def load_reference():
    """Import the unmodified reference and return a namespace of its hot-path classes."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np

    sys.dont_write_bytecode = True  # the reference tree is read-only
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    # our own package also ships a module tree called ``hamiltonian`` (under the product
    # package, never top-level), so a top-level ``hamiltonian`` is always the reference.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import hamiltonian.models.cpu.softmax as m_softmax
    import hamiltonian.models.cpu.mvn_gaussian as m_mvn
    import hamiltonian.models.cpu.logistic as m_logistic
    import hamiltonian.inference.cpu.hmc as i_hmc
    import hamiltonian.inference.cpu.sgmcmc as i_sgmcmc
    import hamiltonian.inference.cpu.sgld as i_sgld
    import hamiltonian.inference.cpu.sghmc as i_sghmc
    import hamiltonian.inference.cpu.sgd as i_sgd
    import hamiltonian.utils as utils

    ns = types.SimpleNamespace()
    ns.softmax = m_softmax.softmax
    ns.mvn_gaussian = m_mvn.mvn_gaussian
    ns.logistic = m_logistic.logistic
    ns.hmc = i_hmc.hmc
    ns.sgmcmc = i_sgmcmc.sgmcmc
    ns.sgld = i_sgld.sgld
    ns.sghmc = i_sghmc.sghmc
    # sghmc is not runnable as shipped: the MRO mixin supplies
    # draw_momentum / accept / potential_energy from hmc.
    ns.sghmc_runnable = type("sghmc_runnable", (i_sghmc.sghmc, i_hmc.hmc), {})
    ns.sgd = i_sgd.sgd
    ns.one_hot = utils.one_hot
    ns.DualAveragingStepSize = i_hmc.DualAveragingStepSize
    return ns
