"""Shared host-side helpers: dict-of-arrays <-> [chains, P] flattening."""
import numpy as np
import torch

from ..runtime import default_context


def array_key(a):
    """Identity of an array for the bind cache (object id + storage pointer + shape)."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return ("t", id(a), a.data_ptr(), tuple(a.shape))
    a = np.asarray(a)
    return ("n", id(a), a.__array_interface__["data"][0], a.shape, a.strides)


def to_numpy(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


class ChainModel:
    """Base of the model classes: owns the context and the dict <-> flat conversions."""

    var_names = ()

    def __init__(self, device=None):
        self._device = device
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = default_context(self._device)
        return self._ctx

    def flatten(self, par, shapes, names=None):
        """par -> (flat [C, P] float32 numpy, squeeze, like).  ``squeeze`` is True when par had no
        chain axis; ``like`` remembers the input container type for the way back."""
        names = names or self.var_names
        first = par[names[0]]
        like = "torch" if isinstance(first, torch.Tensor) else "numpy"
        cols, n_chains, squeeze = [], None, True
        for v in names:
            a = to_numpy(par[v]).astype(np.float32, copy=False)
            base = tuple(shapes[v])
            if a.shape == base:
                a = a.reshape(1, -1)
            elif a.shape[1:] == base:
                squeeze = False
                a = a.reshape(a.shape[0], -1)
            else:
                raise ValueError("parameter %r has shape %s, expected %s or (chains,)+%s" % (v, a.shape, base, base))
            cols.append(a)
        n_chains = max(c.shape[0] for c in cols)
        cols = [np.broadcast_to(c, (n_chains, c.shape[1])) for c in cols]
        return np.ascontiguousarray(np.concatenate(cols, axis=1)), squeeze, like

    def unflatten(self, flat, shapes, squeeze, like="numpy", names=None, dtype=np.float64):
        """flat [..., P] (numpy or torch) -> dict of arrays shaped (..., *shape)."""
        names = names or self.var_names
        if isinstance(flat, torch.Tensor):
            flat = flat.detach().cpu().numpy()
        out, pos = {}, 0
        lead = flat.shape[:-1]
        for v in names:
            n = int(np.prod(shapes[v]))
            a = flat[..., pos:pos + n].reshape(lead + tuple(shapes[v])).astype(dtype)
            pos += n
            if squeeze:
                a = a.reshape(lead[:-1] + tuple(shapes[v])) if lead and lead[-1] == 1 else a
            out[v] = torch.as_tensor(a) if like == "torch" else a
        return out
