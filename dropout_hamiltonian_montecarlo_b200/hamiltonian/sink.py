"""On-disk sample sink: where the samples go when they do not fit in host memory.

The reference's multicore samplers append every sample to one resizable fp32 HDF5 dataset per variable
(``inference/cpu/sghmc_multicore.py:36-53``) and ``hmc.backend_mean`` (``inference/cpu/hmc.py:132-138``) averages such
files.  h5py is not part of this image, so the sink writes the same layout as plain ``.npy`` files
(``<backend>.<var>.npy``, fp32, shape ``[niter, (chains,) *var_shape]``, memory-mapped so chunks stream from the device
ring straight to disk); with h5py importable and a ``.h5`` backend name it writes the HDF5 layout instead.
"""
import os

import numpy as np


def _h5py():
    try:
        import h5py
        return h5py if hasattr(h5py, "File") else None
    except Exception:
        return None


class SampleSink:
    def __init__(self, backend, var_shapes, n_iter, n_chains, squeeze):
        self.backend = str(backend)
        self.names = list(var_shapes)
        self.shapes = {v: tuple(var_shapes[v]) for v in self.names}
        self.n_iter, self.n_chains, self.squeeze = int(n_iter), int(n_chains), bool(squeeze)
        self.pos = 0
        lead = (self.n_iter,) if self.squeeze else (self.n_iter, self.n_chains)
        self._h5 = None
        h5 = _h5py() if self.backend.endswith((".h5", ".hdf5")) else None
        if h5 is not None:
            self._h5 = h5.File(self.backend, "w")
            self.arrays = {v: self._h5.create_dataset(v, lead + self.shapes[v], dtype=np.float32) for v in self.names}
            self.files = {v: self.backend for v in self.names}
        else:
            self.files = {v: "%s.%s.npy" % (self.backend, v.strip("/").replace("/", "_")) for v in self.names}
            d = os.path.dirname(self.backend)
            if d:
                os.makedirs(d, exist_ok=True)
            self.arrays = {v: np.lib.format.open_memmap(self.files[v], mode="w+", dtype=np.float32,
                                                        shape=lead + self.shapes[v]) for v in self.names}

    def append(self, flat):
        """flat: [m, C, P] float32 (numpy) in the model's variable order."""
        m = flat.shape[0]
        off = 0
        for v in self.names:
            n = int(np.prod(self.shapes[v]))
            blk = flat[:, :, off:off + n]
            blk = blk[:, 0].reshape((m,) + self.shapes[v]) if self.squeeze else blk.reshape((m, self.n_chains) + self.shapes[v])
            self.arrays[v][self.pos:self.pos + m] = blk
            off += n
        self.pos += m

    def close(self):
        if self._h5 is not None:
            self._h5.close()
        else:
            for a in self.arrays.values():
                a.flush()
        self.arrays = {}
        return dict(self.files)


def backend_mean(multi_backend, niter, start=None):
    """hmc.py:132-138: sum every file's samples over the iteration axis, add the files up, divide by ``niter``.
    ``multi_backend``: list of backend names as passed to ``sample(backend=...)``."""
    tot = {}
    for backend in multi_backend:
        h5 = _h5py() if str(backend).endswith((".h5", ".hdf5")) else None
        if h5 is not None and os.path.exists(backend):
            with h5.File(backend, "r") as f:
                for v in f.keys():
                    tot[v] = tot.get(v, 0) + np.sum(f[v], axis=0)
            continue
        prefix = os.path.basename(str(backend)) + "."
        d = os.path.dirname(str(backend)) or "."
        for fn in sorted(os.listdir(d)):
            if fn.startswith(prefix) and fn.endswith(".npy"):
                v = fn[len(prefix):-4]
                tot[v] = tot.get(v, 0) + np.sum(np.load(os.path.join(d, fn), mmap_mode="r"), axis=0, dtype=np.float64)
    out = {v: a / float(niter) for v, a in tot.items()}
    if start is not None:
        out = {v: out[v].reshape(np.asarray(start[v]).shape) if out[v].size == np.asarray(start[v]).size else out[v]
               for v in out}
    return out
