"""TEST INFRASTRUCTURE ONLY -- stand-ins for ``chainer`` and ``cupy`` so that the UNMODIFIED reference file
``hamiltonian/models/gpu/mlp.py`` can be executed in this container (neither package is installed, there is no network).

What this pins and what it does not.  The reference's MLP is 95 lines that *compose* Chainer primitives: which layers,
where the three dropouts sit, which loss, how the prior enters ``grad`` / ``negative_log_posterior``, the parameter
names and their order.  Running that file as shipped pins all of it -- the call sequence is the reference's, not a
restatement.  The primitives themselves are implemented here from Chainer's documented semantics (v7 docs):

* ``links.Linear(in, out)``: parameters ``W`` of shape (out, in) and ``b`` (out,); ``y = x W^T + b``;
* ``functions.dropout(x, ratio)`` (train mode): ``mask = rand >= ratio``, ``y = x * mask / (1 - ratio)``;
* ``functions.relu``; ``functions.softmax(x, axis)``;
* ``functions.softmax_cross_entropy(x, t)`` with its defaults ``normalize=True, reduce='mean'``: mean over the batch of
  ``-log softmax(x)[t]``;
* ``Chain.namedparams()``: (path, parameter) pairs, children in sorted name order, ``W`` before ``b``;
* ``Variable.backward()`` on a scalar loss, ``cleargrads()``, ``.data`` / ``.array`` / ``.grad``.

Derivatives come from ``torch.autograd`` (CPU, the dtype of the arrays handed in).  So the MLP oracle is pinned to the
reference's own code under these primitive semantics -- weaker than running real Chainer, stronger than "unpinned".

Dropout masks: ``MASKS`` is a queue of 0/1 keep-masks consumed by successive ``dropout`` calls (the tests inject the
same masks into the oracle and the CUDA path); with an empty queue a mask is drawn from ``RNG``.

Nothing in the product package may import this file.
"""
import contextlib
import sys
import types

import numpy as np
import torch

MASKS = []  # injected keep-masks, consumed front to back by functions.dropout
RNG = np.random.RandomState(0)
DROPOUT_CALLS = []  # (ratio, shape) of every dropout call, for the tests


def _t(x):
    if isinstance(x, Variable):
        return x._t
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x))


class Variable:
    def __init__(self, data=None, requires_grad=False):
        self._t = None
        self._requires_grad = requires_grad
        if data is not None:
            self.data = data

    @property
    def data(self):
        return None if self._t is None else self._t.detach().numpy()

    @data.setter
    def data(self, value):
        t = torch.as_tensor(np.array(value, copy=True))
        if self._requires_grad:
            t.requires_grad_(True)
        self._t = t

    array = data

    @property
    def grad(self):
        return None if self._t is None or self._t.grad is None else self._t.grad.numpy()

    def cleargrad(self):
        if self._t is not None:
            self._t.grad = None

    def backward(self):
        self._t.backward()

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return self.data.dtype


def _wrap(t):
    v = Variable()
    v._t = t
    return v


class Parameter(Variable):
    def __init__(self, shape):
        super().__init__(np.zeros(shape, dtype=np.float32), requires_grad=True)


class Link:
    def __init__(self):
        self._params = []
        self._children = []
        self._update_enabled = True

    @contextlib.contextmanager
    def init_scope(self):
        before = set(self.__dict__)
        yield
        for name in sorted(set(self.__dict__) - before):
            obj = self.__dict__[name]
            if isinstance(obj, Parameter):
                self._params.append(name)
            elif isinstance(obj, Link):
                self._children.append(name)

    def namedparams(self, include_uninit=True):
        for name in sorted(self._params):
            yield "/" + name, self.__dict__[name]
        for cname in sorted(self._children):
            for path, p in self.__dict__[cname].namedparams(include_uninit):
                yield "/" + cname + path, p

    def params(self, include_uninit=True):
        for _, p in self.namedparams(include_uninit):
            yield p

    def to_gpu(self, device=None):
        return self

    def to_cpu(self):
        return self

    def enable_update(self):
        self._update_enabled = True

    def cleargrads(self):
        for p in self.params():
            p.cleargrad()

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)


class Chain(Link):
    pass


class ChainList(Link):
    pass


class Linear(Link):
    def __init__(self, in_size, out_size=None):
        super().__init__()
        with self.init_scope():
            self.W = Parameter((out_size, in_size))
            self.b = Parameter((out_size,))

    def forward(self, x):
        return _wrap(_t(x) @ self.W._t.T + self.b._t)


def relu(x):
    return _wrap(torch.relu(_t(x)))


def dropout(x, ratio=.5, **kwargs):
    t = _t(x)
    DROPOUT_CALLS.append((ratio, tuple(t.shape)))
    if MASKS:
        mask = np.asarray(MASKS.pop(0))
        assert mask.shape == tuple(t.shape), (mask.shape, tuple(t.shape))
    else:
        mask = RNG.rand(*t.shape) >= ratio
    scale = 1.0 / (1.0 - ratio)
    return _wrap(t * torch.as_tensor(mask.astype(np.float64) * scale).to(t.dtype))


def softmax(x, axis=1):
    return _wrap(torch.softmax(_t(x), dim=axis))


def softmax_cross_entropy(x, t, normalize=True, reduce="mean"):
    z = _t(x)
    lab = torch.as_tensor(np.asarray(t).astype(np.int64))
    lse = torch.logsumexp(z, dim=1)
    nll = lse - z[torch.arange(z.shape[0]), lab]
    assert reduce == "mean" and normalize
    return _wrap(nll.mean())


def install():
    """Put fake ``chainer`` / ``cupy`` module trees into sys.modules (idempotent).  Returns (chainer, cupy)."""
    if "chainer" in sys.modules and getattr(sys.modules["chainer"], "__bhmc_shim__", False):
        return sys.modules["chainer"], sys.modules["cupy"]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    F = mod("chainer.functions", relu=relu, dropout=dropout, softmax=softmax, softmax_cross_entropy=softmax_cross_entropy)
    L = mod("chainer.links", Linear=Linear)
    cuda = mod("chainer.backends.cuda")
    backends = mod("chainer.backends", cuda=cuda)
    backend = mod("chainer.backend")
    extensions = mod("chainer.training.extensions")
    training = mod("chainer.training", extensions=extensions)
    mnist = mod("chainer.datasets.mnist")
    datasets = mod("chainer.datasets", mnist=mnist)
    empty = {n: mod("chainer." + n) for n in ("gradient_check", "utils", "initializers", "iterators", "optimizers", "serializers")}
    chainer = mod("chainer", functions=F, links=L, backends=backends, backend=backend, training=training, datasets=datasets,
                  Variable=Variable, Parameter=Parameter, Link=Link, Chain=Chain, ChainList=ChainList,
                  Function=type("Function", (), {}), FunctionNode=type("FunctionNode", (), {}),
                  report=lambda *a, **k: None, __bhmc_shim__=True, **empty)
    # cupy: the subset mlp.py touches, on host arrays
    cupy = mod("cupy", asarray=np.asarray, asnumpy=np.asarray, sum=np.sum, square=np.square, int=int, float=float,
               ndarray=np.ndarray, random=np.random, __bhmc_shim__=True)
    return chainer, cupy


def load_reference_mlp(reference_root="/root/reference"):
    """Execute the unmodified ``hamiltonian/models/gpu/mlp.py`` under the shim and return its module."""
    import importlib.util
    import os

    install()
    path = os.path.join(reference_root, "hamiltonian", "models", "gpu", "mlp.py")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    spec = importlib.util.spec_from_file_location("_reference_gpu_mlp", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
