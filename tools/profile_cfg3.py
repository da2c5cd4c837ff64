"""Short profiling target: a few SGLD epochs of the cfg3 workload (softmax 60000 x 784 x 10, minibatch 500, 128 chains)
through the library call the bench block times (bhmc_sampler_sg_run -> k_sg_persistent2, one cooperative launch per
epoch).  What ncu wraps; nothing printed here is a bench value."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import WORKLOADS, synth
from dropout_hamiltonian_montecarlo_b200._lib import KIND, PREC
from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
from dropout_hamiltonian_montecarlo_b200.runtime import SamplerHandle

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--chains", type=int, default=128)
ap.add_argument("--batch", type=int, default=500)
ap.add_argument("--warm-epochs", type=int, default=0, help="untimed epochs before the timed ones (a fresh box idles at low clocks)")
a = ap.parse_args()
wl = WORKLOADS["cfg2"]
dev = torch.device("cuda", 0)
X, y = synth(wl["N"], wl["D"], wl["K"], 0, device=dev)
m = softmax({"alpha": wl["alpha"]}, precision=a.precision)
h = m.bind(X, y, n_classes=wl["K"])
C, B, N = a.chains, a.batch, wl["N"]
nb = N // B
s = SamplerHandle(h.ctx, h, KIND["sgld"], C, seed=1, chain_id0=0, precision=PREC[a.precision])
s.set_q(np.zeros((C, h.P), np.float32))
s.sg_run(1, 1, B, 1e-5, n_rows=N)  # one burn-in and one sampling epoch, as the bench block's warm-up
if a.warm_epochs:
    s.sg_run(a.warm_epochs, 0, B, 1e-5, n_rows=N, step0=2 * nb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = h.ctx.launches
e0.record()
o = s.sg_run(a.epochs, 0, B, 1e-5, n_rows=N, step0=2 * nb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("sgld: %d epochs x %d minibatch steps, %d chains: %.3f ms = %.2f us per step, %.3g grad-evals/s, %d launches" % (
    a.epochs, nb, C, ms, 1e3 * ms / (a.epochs * nb), o["n_grad_evals"] / (ms * 1e-3), h.ctx.launches - l0))
s.close()
