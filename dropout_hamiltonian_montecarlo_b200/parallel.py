"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Two ways the path shards; the reference has neither (host multiprocessing only):

* **chains** -- independent units, no data-path collective.  ``shard_chains`` gives each rank a
  contiguous block of global chain ids; Philox streams are keyed by the global id, so the union
  of all ranks' chains is the same set of chains for any world size.
* **rows** -- full-batch HMC on a row-sharded data matrix (BASELINE config 5): every rank holds
  all chains and ``N/G`` rows; each gradient evaluation is followed by ONE grouped NCCL all-reduce(sum)
  of the gradient (fp32) and the log-likelihoods (fp64), enqueued by the C driver on the context
  stream (``RowComm`` -> ``bhmc_sampler_set_row_comm``; csrc/comm.cu).  The Gaussian-prior term
  ``alpha*q`` is added once by giving every rank ``alpha/G``; energies are normalised by the global
  row count.  Everything after the all-reduce (update, Philox draws, accept) runs redundantly and
  identically on every rank.  ``RowShardHook`` (two torch.distributed collectives from a ctypes
  callback per evaluation) is the round-1 mechanism; it remains for process groups that are not NCCL
  (the gloo CPU tests) and as the A/B partner of the C path (``BHMC_ROW_COMM=hook``).
"""
import os
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_chains(n_chains_total, rank, world):
    """-> (global id of the first local chain, number of local chains); remainders go to low ranks."""
    base, rem = divmod(int(n_chains_total), int(world))
    n = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n


def shard_rows(n_rows_total, rank, world, align=8):
    """-> (first row, number of rows) of this rank; shard boundaries are multiples of ``align``
    (TMA needs 16-byte aligned inner coordinates for the transposed operand)."""
    per = -(-int(n_rows_total) // int(world))
    per = -(-per // align) * align
    start = min(rank * per, n_rows_total)
    stop = min(start + per, n_rows_total)
    return start, stop - start


def allreduce_sum_(tensors, group=None):
    """In-place all-reduce(sum) of a list of tensors (one collective per tensor; they are two)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        if t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class RowComm:
    """NCCL communicator owned by libbhmc.so (csrc/comm.cu) over the ranks of a torch.distributed group: rank 0 makes
    the NCCL unique id, the 128 bytes travel through the existing process group, every rank joins."""

    def __init__(self, ctx, group=None):
        self.ctx = ctx
        L = ctx.L
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(L.bhmc_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        h = C.c_void_p()
        _lib.check(L.bhmc_comm_create(ctx.handle, ident, rank, world, C.byref(h)))
        self.handle, self.rank, self.world = h, rank, world

    def allreduce(self, g=None, stat=None):
        """In-place sum over ranks of a device fp32 tensor and/or a device fp64 tensor (one grouped NCCL call)."""
        _lib.check(self.ctx.L.bhmc_comm_allreduce(self.handle, C.c_void_p(g.data_ptr() if g is not None else 0),
                                                  g.numel() if g is not None else 0,
                                                  C.c_void_p(stat.data_ptr() if stat is not None else 0),
                                                  stat.numel() if stat is not None else 0))

    def attach(self, sampler):
        _lib.check(self.ctx.L.bhmc_sampler_set_row_comm(sampler.handle, self.handle))

    def close(self):
        if self.handle is not None:
            self.ctx.L.bhmc_comm_destroy(self.handle)
            self.handle = None


_row_comms = {}


def row_comm_for(ctx, group=None):
    """One RowComm per (context, group), or None when the C path does not apply (not NCCL, or BHMC_ROW_COMM=hook)."""
    if os.environ.get("BHMC_ROW_COMM", "c") == "hook":
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(group) != "nccl":
        return None
    key = (id(ctx), id(group))
    if key not in _row_comms:
        _row_comms[key] = RowComm(ctx, group)
    return _row_comms[key]


class RowShardHook:
    """Gradient hook (bhmc_sampler_set_grad_hook) that all-reduces g and loglik across ranks.

    The hook receives raw device pointers; they are wrapped as tensors without copying through
    ``torch.frombuffer``-like views of the sampler's own buffers, registered up front."""

    def __init__(self, sampler, group=None):
        self.sampler = sampler
        self.group = group
        self.calls = 0
        ctx = sampler.ctx
        # views over the sampler's resident gradient / scalar buffers
        gptr = C.c_void_p()
        _lib.check(ctx.L.bhmc_sampler_state_ptr(sampler.handle, 2, C.byref(gptr)))
        self._views = {}
        self._cb = _lib.GRAD_HOOK(self._call)
        _lib.check(ctx.L.bhmc_sampler_set_grad_hook(sampler.handle, C.cast(self._cb, C.c_void_p), None))

    def _view(self, ptr, n, dtype):
        key = (ptr, n, dtype)
        v = self._views.get(key)
        if v is None:
            itemsize = torch.empty((), dtype=dtype).element_size()
            buf = _DevicePointer(ptr, n * itemsize, self.sampler.ctx.device.index)
            v = torch.as_tensor(buf, device=self.sampler.ctx.device).view(dtype)[:n]
            self._views[key] = v
        return v

    def _call(self, user, g_ptr, stat_ptr, rows, ld):
        try:
            self.calls += 1
            ts = [self._view(stat_ptr, rows, torch.float64)]
            if g_ptr:
                ts.append(self._view(g_ptr, rows * ld, torch.float32))
            allreduce_sum_(ts, self.group)
            return 0
        except Exception as e:  # never raise through the C frame
            print("bhmc row-shard hook failed:", repr(e))
            return 1


class _DevicePointer:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer."""

    def __init__(self, ptr, nbytes, device_index):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
