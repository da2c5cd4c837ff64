"""Secondary workloads (not the driver's headline): BASELINE configs 3 and 4 and the fused-update roofline.

    python tools/bench_extra.py sgld   [--chains 128] [--epochs 2]      cfg3: SGLD softmax, minibatch 500
    python tools/bench_extra.py mlp    [--chains 16]  [--steps 20]      cfg4: SGHMC MLP 784-512-512-10, minibatch 500
    python tools/bench_extra.py update [--chains 64]                    fused update kernels vs HBM peak (cfg4-sized state)
Prints one JSON line each."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import peaks, synth
from dropout_hamiltonian_montecarlo_b200._lib import KIND, PREC
from dropout_hamiltonian_montecarlo_b200.runtime import MlpHandle, SamplerHandle, SoftmaxHandle, default_context

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["sgld", "sghmc", "mlp", "update", "rows", "stream"])
ap.add_argument("--rows", type=int, default=1000000)
ap.add_argument("--features", type=int, default=2048)
ap.add_argument("--classes", type=int, default=38)
ap.add_argument("--chains", type=int, default=0)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--precision", default="bf16x3")
a = ap.parse_args()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
ctx = default_context()
dev = ctx.device
peak_tf, peak_bw, src = peaks()
N, D, K, B = 60000, 784, 10, 500
if a.what != "rows":
    X, y = synth(N, D, K, 0, device=dev)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3, out


if a.what == "stream":
    # profiling target for the elementwise kernels of the streaming schedule at the cfg4-sized state (P = 669 706)
    # where they are HBM-bound: a toy MVN model would not do, so a wide softmax stands in: D*K + K ~ 670k parameters
    Ns, Ds, Ks, C = 256, 65535, 10, 32
    g = torch.Generator(device=dev).manual_seed(1)
    Xs = torch.rand(Ns, Ds, generator=g, device=dev)
    ys = torch.randint(0, Ks, (Ns,), generator=g, device=dev, dtype=torch.int32)
    h = SoftmaxHandle(ctx, Ns, Ds, Ks, 0.01)
    h.bind(Xs, ys, 1 | (1 << PREC["bf16"]))
    s = SamplerHandle(ctx, h, KIND["hmc"], C, seed=1, precision=PREC["bf16"], sweep=list(zip(h.var_off, h.var_len)))
    s.set_q(np.zeros((C, h.P), np.float32))
    s.hmc_run(2, 1e-4, 5e-4, keep_samples=False, schedule="streaming")
    ctx.timing(True)
    dt, out = timed(lambda: s.hmc_run(a.steps, 1e-4, 1e-3, step0=2, keep_samples=True, schedule="streaming"))
    t_upd, n_upd = ctx.kernel_time(3)
    print(json.dumps({"workload": "streaming-schedule elementwise kernels, P=%d, %d chains" % (h.P, C), "phases": out["n_phases"],
                      "update_group_ms": t_upd, "update_group_launches": n_upd, "seconds": dt}))
elif a.what == "rows":
    # BASELINE config 5: full-batch HMC, rows sharded over the ranks, one all-reduce (NCCL) per gradient evaluation.
    # Launch with torchrun; every rank synthesises only its own shard (abs(N(0,1)) features).
    import torch.distributed as dist
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
    from dropout_hamiltonian_montecarlo_b200.parallel import shard_rows
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Nt, Dt, Kt, C = a.rows, a.features, a.classes, a.chains or 8
    r0, nloc = shard_rows(Nt, rank, world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    Xs = torch.randn(nloc, Dt, generator=g, device=dev).abs_()
    ys = torch.randint(0, Kt, (nloc,), generator=g, device=dev, dtype=torch.int32)
    m = softmax({"alpha": 0.01}, precision=a.precision, row_sharded=world > 1)
    eps, path = 1e-7, 1e-6  # E[L] = 10
    smp = hmc(m, {"weights": np.zeros((Dt, Kt), np.float32), "bias": np.zeros(Kt, np.float32)}, path_length=path,
              step_size=eps, verbose=False, n_chains=C, seed=3, path_length_mode="shared")
    smp.sample(niter=1, burnin=0, X_train=Xs, y_train=ys)  # warm-up (binds, allocates)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    smp.sample(niter=a.steps, burnin=0, X_train=Xs, y_train=ys)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    n_grad = smp.last_run["n_grad_evals"]
    if rank == 0:
        print(json.dumps({"workload": "cfg5 full-batch HMC softmax %dx%dx%d, %d chains replicated, rows sharded over %d GPU(s)"
                          % (Nt, Dt, Kt, C, world), "precision": a.precision, "grad_evals_per_s": n_grad / dt,
                          "ms_per_grad_eval_all_chains": 1e3 * dt / (n_grad / C),
                          "algorithmic_tflops_total": n_grad * 4.0 * Nt * Dt * Kt / dt / 1e12, "n_gpus": world,
                          "allreduce_bytes_per_eval": 4 * C * m._bound[1].ld + 8 * C}))
    if world > 1:
        dist.destroy_process_group()
elif a.what == "sgld":
    C = a.chains or 128
    h = SoftmaxHandle(ctx, N, D, K, 0.01)
    h.bind(X, y, 1 | (1 << PREC[a.precision]))
    s = SamplerHandle(ctx, h, KIND["sgld"], C, seed=1, precision=PREC[a.precision])
    s.set_q(np.zeros((C, h.P), np.float32))
    s.sg_run(0, 1, B, 1e-5, n_rows=N)  # warm-up epoch
    l0 = ctx.launches
    dt, out = timed(lambda: s.sg_run(a.epochs, 0, B, 1e-5, n_rows=N))
    nb = N // B
    print(json.dumps({"workload": "cfg3 SGLD softmax 60000x784x10, minibatch 500, %d chains/GPU" % C, "precision": a.precision,
                      "grad_evals_per_s": out["n_grad_evals"] / dt, "us_per_minibatch_step": 1e6 * dt / (a.epochs * nb),
                      "launches_per_step": (ctx.launches - l0) / (a.epochs * nb),
                      "algorithmic_tflops": out["n_grad_evals"] * 4.0 * B * D * K / dt / 1e12}))
elif a.what == "sghmc":
    # BASELINE config 3, second half: SGHMC softmax on minibatches of 500 (sghmc.py:19-39 per minibatch: momentum draw,
    # L-1 friction+noise leapfrog iterations on the window, Metropolis test), shared path lengths, E[L] = 10
    C = a.chains or 128
    h = SoftmaxHandle(ctx, N, D, K, 0.01)
    h.bind(X, y, 1 | (1 << PREC[a.precision]))
    s = SamplerHandle(ctx, h, KIND["sghmc"], C, seed=1, precision=PREC[a.precision], shared_path=True, sghmc_descent=True)
    s.set_q(np.zeros((C, h.P), np.float32))
    eps, path = 1e-5, 1e-4
    nb = N // B
    for j in range(8):
        s.hmc_run(1, eps, path, row0=j * B, nrows=B, step0=j, keep_samples=False, keep_stats=False)
    n_grad = 0
    l0 = ctx.launches

    def run():
        global n_grad
        for j in range(a.steps):
            o = s.hmc_run(1, eps, path, row0=(j % nb) * B, nrows=B, step0=8 + j, keep_samples=False, keep_stats=False)
            n_grad += o["n_grad_evals"]
    dt, _ = timed(run)
    print(json.dumps({"workload": "cfg3 SGHMC softmax 60000x784x10, minibatch 500, %d chains/GPU, E[L]=10, shared path lengths" % C,
                      "precision": a.precision, "grad_evals_per_s": n_grad / dt, "us_per_sghmc_step": 1e6 * dt / a.steps,
                      "us_per_grad_eval_all_chains": 1e6 * dt / (n_grad / C), "launches_per_step": (ctx.launches - l0) / a.steps,
                      "algorithmic_tflops": n_grad * 4.0 * B * D * K / dt / 1e12}))
elif a.what == "mlp":
    C = a.chains or 16
    n_mid = 512
    h = MlpHandle(ctx, N, D, n_mid, K, 0.01, 0.1, seed=3)
    h.bind(X, y)
    s = SamplerHandle(ctx, h, KIND["sghmc"], C, seed=1, precision=PREC[a.precision], sweep=[(0, h.P)], shared_path=True,
                      sghmc_descent=True)
    rs = np.random.RandomState(0)
    s.set_q(rs.normal(0, 0.05, (C, h.P)).astype(np.float32))
    eps, path = 1e-3, 5e-3  # E[L] = 5
    s.hmc_run(2, eps, path, row0=0, nrows=B, keep_samples=False)
    n_grad = 0

    def run():
        global n_grad
        for i in range(a.steps):
            o = s.hmc_run(1, eps, path, row0=(i % (N // B)) * B, nrows=B, step0=10 + i, keep_samples=False)
            n_grad += o["n_grad_evals"]
    dt, _ = timed(run)
    flops = 6.0 * B * (D * n_mid + n_mid * n_mid + n_mid * K) - 2.0 * B * D * n_mid
    print(json.dumps({"workload": "cfg4 SGHMC MLP 784-512-512-10 dropout 0.1, minibatch 500, %d chains/GPU, joint sweep" % C,
                      "precision": a.precision, "grad_evals_per_s": n_grad / dt,
                      "algorithmic_tflops": n_grad * flops / dt / 1e12, "ms_per_sghmc_step": 1e3 * dt / a.steps}))
else:
    import ctypes as Ct
    from dropout_hamiltonian_montecarlo_b200._lib import check
    C = a.chains or 64
    out = {"peak_hbm_gbs": peak_bw, "peak_source": src, "kernels": {}}
    names = {0: ("hmc kick+drift", 20), 1: ("sghmc friction+philox+drift", 20), 2: ("sgld philox", 16),
             3: ("accept/select+sample sink", 20), 4: ("momentum draw philox", 16)}
    for label, P in (("cfg4 MLP P=669706", 669706), ("cfg2 softmax P=7850", 7850)):
        for which, (nm, bpp) in names.items():
            ms = Ct.c_double()
            check(ctx.L.bhmc_bench_update(ctx.handle, which, C, P, 50, Ct.byref(ms)))
            gbs = bpp * C * P / (ms.value * 1e-3) / 1e9
            out["kernels"]["%s | %s" % (label, nm)] = {"ms": ms.value, "bytes_per_param": bpp, "achieved_gbs": gbs,
                                                        "frac_of_hbm_peak": gbs / peak_bw}
    out["chains"] = C
    print(json.dumps(out))
