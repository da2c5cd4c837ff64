"""CPU, world_size 2, gloo: the host-side sharding logic of the multi-GPU paths.

* chain sharding: the ranks' blocks tile the global chain ids exactly, for any world size;
* row sharding: per-rank oracle gradients on row shards with alpha/G, all-reduced over gloo, equal the
  full-data gradient (this is exactly what RowShardHook does with the CUDA buffers over NCCL).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_chains_tiles_ids():
    from dropout_hamiltonian_montecarlo_b200.parallel import shard_chains
    for total in (1, 7, 64, 1024, 1000):
        for world in (1, 2, 3, 8):
            ids = []
            for r in range(world):
                s, n = shard_chains(total, r, world)
                ids += list(range(s, s + n))
            assert ids == list(range(total))


def test_shard_rows_aligned_and_complete():
    from dropout_hamiltonian_montecarlo_b200.parallel import shard_rows
    for total in (1000, 60000, 1000003):
        for world in (1, 2, 4, 8):
            pos = 0
            for r in range(world):
                s, n = shard_rows(total, r, world)
                assert s == pos and s % 8 == 0
                pos += n
            assert pos == total


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dropout_hamiltonian_montecarlo_b200.parallel import allreduce_sum_, shard_rows
    from oracle import hamiltonian_oracle as O
    rs = np.random.RandomState(0)
    N, D, K, C, alpha = 203, 17, 5, 3, 0.25
    X = rs.rand(N, D)
    y = rs.randint(0, K, N)
    Y = O.one_hot(y, K)
    W = rs.normal(0, .3, (C, D, K))
    b = rs.normal(0, .3, (C, K))
    s, n = shard_rows(N, rank, world)
    g = np.zeros((C, (D + 1) * K))
    ll = np.zeros(C)
    for c in range(C):
        par = {"weights": W[c], "bias": b[c]}
        gc = O.softmax_grad(par, X[s:s + n], Y[s:s + n], alpha / world)  # prior split over ranks
        g[c] = O.flatten_par(gc, ["weights", "bias"])
        ll[c] = O.softmax_log_likelihood(par, X[s:s + n], Y[s:s + n])
    tg, tl = torch.from_numpy(g), torch.from_numpy(ll)
    allreduce_sum_([tl, tg])
    ok = True
    for c in range(C):
        par = {"weights": W[c], "bias": b[c]}
        ref = O.flatten_par(O.softmax_grad(par, X, Y, alpha), ["weights", "bias"])
        ok &= np.allclose(tg[c].numpy(), ref, rtol=1e-12, atol=1e-12)
        ok &= np.isclose(tl[c].item(), O.softmax_log_likelihood(par, X, Y), rtol=1e-12)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_row_sharded_gradient_allreduce_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
