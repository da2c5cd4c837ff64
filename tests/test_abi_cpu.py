"""CPU: the C-ABI library loads without a GPU, exports every symbol include/bhmc.h declares,
fails loudly (no fallback) when asked to compute, and its Philox matches the Random123 KATs."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from dropout_hamiltonian_montecarlo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "bhmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bhmc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(L):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "include/bhmc.h declares %s but libbhmc.so does not export it" % n
    assert set(names) == set(_lib.exported_symbols()), "ctypes prototypes drifted from the header"


def test_struct_layout_matches_header(L):
    # sizes the C compiler sees (gcc on the header) == ctypes mirrors
    prog = r'''
    #include <stdio.h>
    #include "bhmc.h"
    int main(){ printf("%zu %zu %zu\n", sizeof(bhmc_sampler_config), sizeof(bhmc_hmc_run), sizeof(bhmc_sg_run)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    assert [int(x) for x in out] == [C.sizeof(_lib.SamplerConfig), C.sizeof(_lib.HmcRun), C.sizeof(_lib.SgRun)]


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = L.bhmc_ctx_create(0, None, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in L.bhmc_last_error()
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
    m = softmax({"alpha": 0.01})
    with pytest.raises(_lib.BhmcError):
        m.grad({"weights": np.zeros((4, 3)), "bias": np.zeros(3)}, X_train=np.zeros((8, 4)), y_train=np.eye(3)[[0] * 8])


KAT = [  # Random123 kat_vectors, philox4x32 10 rounds: counter, key -> output
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,exp", KAT)
def test_philox_known_answers(L, ctr, key, exp):
    o = (C.c_uint32 * 4)()
    L.bhmc_philox4x32_host((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), o)
    assert list(o) == exp


def test_philox_uniform_host_range_and_determinism(L):
    u = np.array([L.bhmc_philox_uniform_host(7, c, s, 0x02000000) for c in range(64) for s in range(64)])
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.02 and len(np.unique(u)) == u.size
    assert L.bhmc_philox_uniform_host(7, 3, 5, 0x02000000) == L.bhmc_philox_uniform_host(7, 3, 5, 0x02000000)


def test_host_helpers_without_gpu():
    from dropout_hamiltonian_montecarlo_b200.hamiltonian import utils, _base
    assert (utils.one_hot([0, 2, 1], 3) == np.eye(3)[[0, 2, 1]]).all()
    m = _base.ChainModel()
    m.var_names = ("weights", "bias")
    shapes = {"weights": (4, 3), "bias": (3,)}
    par = {"weights": np.arange(12.).reshape(4, 3), "bias": np.arange(3.) + 100}
    flat, squeeze, like = m.flatten(par, shapes)
    assert flat.shape == (1, 15) and squeeze and like == "numpy"
    back = m.unflatten(flat, shapes, squeeze)
    assert (back["weights"] == par["weights"]).all() and (back["bias"] == par["bias"]).all()
    par2 = {"weights": np.zeros((5, 4, 3)), "bias": np.ones((5, 3))}
    flat2, squeeze2, _ = m.flatten(par2, shapes)
    assert flat2.shape == (5, 15) and not squeeze2
    back2 = m.unflatten(np.zeros((7, 5, 15), np.float32), shapes, False)
    assert back2["weights"].shape == (7, 5, 4, 3) and back2["bias"].shape == (7, 5, 3)


def test_install_as_hamiltonian_aliases_every_module():
    """`import hamiltonian.models.gpu.logistic` after install_as_hamiltonian() used to raise ImportError (the alias
    list was kept by hand); the tree is now walked."""
    import importlib
    import sys
    import dropout_hamiltonian_montecarlo_b200 as b200
    saved = {k: v for k, v in sys.modules.items() if k == "hamiltonian" or k.startswith("hamiltonian.")}
    try:
        b200.install_as_hamiltonian()
        for name in ("hamiltonian.models.gpu.logistic", "hamiltonian.models.gpu.softmax", "hamiltonian.models.gpu.mlp",
                     "hamiltonian.models.gpu.mvn_gaussian", "hamiltonian.inference.gpu.hmc", "hamiltonian.inference.gpu.sgld",
                     "hamiltonian.inference.gpu.sghmc", "hamiltonian.inference.gpu.sgd", "hamiltonian.sink", "hamiltonian.utils",
                     "hamiltonian._base"):
            mod = importlib.import_module(name)
            assert mod.__name__.startswith("dropout_hamiltonian_montecarlo_b200.hamiltonian"), name
        from hamiltonian.models.gpu.logistic import logistic  # noqa: F401
    finally:
        for k in [k for k in sys.modules if k == "hamiltonian" or k.startswith("hamiltonian.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_nccl_is_resolved_at_run_time_not_linked(L):
    """libbhmc.so must load without libnccl (single-GPU users never need it): no DT_NEEDED entry; the comm entry points
    exist and report a clean error / a version once the library can be dlopen'ed."""
    import subprocess
    from dropout_hamiltonian_montecarlo_b200 import _lib
    needed = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in needed.lower()
    assert L.bhmc_nccl_version() >= 0
    assert L.bhmc_comm_world(None) == 0


def test_step_trace_is_lazy_and_list_like():
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import _StepTrace
    made = []

    def make(i):
        made.append(i)
        return {"x": i}
    t = _StepTrace(5, make)
    assert len(t) == 5 and made == []
    assert t[1] == [{"x": 1}] and t[-1] == [{"x": 4}] and made == [1, 4]
    assert [e[0]["x"] for e in t] == [0, 1, 2, 3, 4]
    assert t[1:3] == [[{"x": 1}], [{"x": 2}]]
    with pytest.raises(IndexError):
        t[5]
