"""Short profiling target: a few gradient evaluations of the cfg2 workload (what ncu wraps)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import WORKLOADS, synth
from dropout_hamiltonian_montecarlo_b200._lib import PREC
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--evals", type=int, default=3)
ap.add_argument("--chains", type=int, default=0)
ap.add_argument("--loglik", action="store_true")
ap.add_argument("--data", default="dense", choices=["dense", "pixels"])
a = ap.parse_args()
import bench
bench.DATA_KIND = a.data
wl = WORKLOADS[a.workload]
C = a.chains or wl["C"]
dev = torch.device("cuda", 0)
X, y = synth(wl["N"], wl["D"], wl["K"], 0, device=dev)
m = softmax({"alpha": wl["alpha"]}, precision=a.precision)
h = m.bind(X, y, n_classes=wl["K"])
q = h.pack(np.random.RandomState(0).normal(0, 0.01, (C, h.P)).astype(np.float32))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
h.ctx.timing(True)
for i in range(a.evals):
    e0.record()
    g, ll = h.grad(q, 0, wl["N"], PREC[a.precision], not a.loglik)
    e1.record()
    torch.cuda.synchronize()
    print("eval %d: %.3f ms  ll[0]=%.3f |g|=%.4e" % (i, e0.elapsed_time(e1), ll[0].item(), g.norm().item() if g is not None else 0.0))

tf, nf = h.ctx.kernel_time(0)
tb, nb = h.ctx.kernel_time(1)
print("kernel-only (CUDA events around the launches): fwd %.1f us x%d   bwd+reduce %.1f us x%d" % (1e3 * tf / max(nf, 1), nf, 1e3 * tb / max(nb, 1), nb))
