"""torchrun --nproc-per-node G tools/multi_gpu_check.py : multi-GPU parity of the two sharding modes.
Prints one line per check on rank 0; exits non-zero on failure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
from dropout_hamiltonian_montecarlo_b200.parallel import shard_chains, shard_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
rs = np.random.RandomState(0)
N, D, K, alpha = 4096 + 40, 96, 10, 0.05
X = rs.rand(N, D).astype(np.float32)
y = rs.randint(0, K, N).astype(np.int32)
eps, path = 2e-5, 2e-4

# ---- rows sharded: every rank holds all chains, N/G rows; all-reduce after every gradient -------------
C = 4
W0 = rs.normal(0, .05, (C, D, K)).astype(np.float32)
b0 = rs.normal(0, .05, (C, K)).astype(np.float32)
s0, n0 = shard_rows(N, rank, world)
m_sh = softmax({"alpha": alpha}, precision="bf16x3", row_sharded=True)
smp = hmc(m_sh, {"weights": W0, "bias": b0}, path_length=path, step_size=eps, verbose=False, seed=11)
post, loss, _, _ = smp.sample(niter=4, burnin=1, X_train=X[s0:s0 + n0], y_train=y[s0:s0 + n0])
t = torch.as_tensor(post["weights"]).cuda()
gathered = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
same = all(torch.equal(gathered[0], g) for g in gathered)
m_full = softmax({"alpha": alpha}, precision="bf16x3")
ref = hmc(m_full, {"weights": W0, "bias": b0}, path_length=path, step_size=eps, verbose=False, seed=11)
rpost, rloss, _, _ = ref.sample(niter=4, burnin=1, X_train=X, y_train=y)
err = np.abs(post["weights"] - rpost["weights"]).max() / np.abs(rpost["weights"]).max()
lerr = np.abs(loss - rloss).max()
moved = np.abs(rpost["weights"][-1] - W0).max() > 0
if rank == 0:
    print("row-sharded: replicas identical=%s  vs single-GPU max rel err=%.2e  loss err=%.2e moved=%s hook calls=%d"
          % (same, err, lerr, moved, getattr(getattr(smp._sampler[1], "_row_hook", None), "calls", -1)))
ok &= same and err < 1e-4 and lerr < 1e-5 and moved

# ---- chains sharded: no collective on the data path ---------------------------------------------------
Ct = 6 * world
Wc = rs.normal(0, .05, (Ct, D, K)).astype(np.float32)
bc = rs.normal(0, .05, (Ct, K)).astype(np.float32)
c0, cn = shard_chains(Ct, rank, world)
loc = hmc(softmax({"alpha": alpha}, precision="bf16x3"), {"weights": Wc[c0:c0 + cn], "bias": bc[c0:c0 + cn]},
          path_length=path, step_size=eps, verbose=False, seed=5, chain_id0=c0)
lpost, _, _, _ = loc.sample(niter=3, burnin=0, X_train=X, y_train=y)
allc = hmc(softmax({"alpha": alpha}, precision="bf16x3"), {"weights": Wc, "bias": bc}, path_length=path, step_size=eps,
           verbose=False, seed=5)
apost, _, _, _ = allc.sample(niter=3, burnin=0, X_train=X, y_train=y)
cerr = np.abs(lpost["weights"] - apost["weights"][:, c0:c0 + cn]).max()
flag = torch.tensor([float(cerr)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    print("chain-sharded: max |local - global run| over ranks = %.2e" % flag.item())
ok &= flag.item() < 1e-6
res = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if res.item() == 1.0 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if res.item() == 1.0 else 1)
