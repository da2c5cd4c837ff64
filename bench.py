#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: gradient evaluations / second
(chains x leapfrog) on the MNIST-shaped softmax BNN; ESS / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

Headline workload at N=1 = BASELINE configs[1]: full-batch HMC, softmax regression, synthetic
60 000 x 784, 10 classes, 64 chains (per GPU; weak scaling: chains are independent units and
shard over ranks with no data-path collective).  A "step" is one HMC transition of every chain:
momentum draw, L-1 Gauss-Seidel leapfrog sweeps (each sub-step one full gradient = forward +
backward GEMM), Metropolis test, sample store.

`value`  : chain-gradient evaluations applied / s, inputs resident in HBM, CUDA-event timed,
           max over ranks.
`e2e`    : same metric through the public API (`hmc.sample`) with HOST buffers: every step binds the
           data from pinned host memory (H2D + operand preparation inside the timed region) and
           reads the samples / losses back (D2H).
`roofline`: dominant kernel group of the timed region (the backward GEMM -- k_tc_bwd_sk for launches of >= ~20
           chains, k_tc_gemm<bwd> below -- in every run so far; the line names whichever group took longer),
           algorithmic flops 2*N*D*K per chain it carried / its
           duration (CUDA events on the launching stream, sampled live inside the timed region)
           against the measured sustained bf16 peak.
`cpu_baseline`: the NumPy oracle port of the reference path (fp64, BLAS threads = host cores) on a
           bounded sample of the same workload, same box; plus ESS/s of the CPU arm with the same
           estimator on the small-N variant SURVEY 8(d) names.  Rank 0 at N = 1 only.

The other BASELINE configs ride on the same JSON line as secondary blocks (driver-visible, each with its
own value / ms_per_step / roofline):
`cfg3`   : configs[2] -- SGLD and SGHMC softmax, minibatch 500, 128 chains per GPU (1024 over 8).
`cfg4`   : configs[3] -- SGHMC Bayesian MLP 784-512-512-10, minibatch 500, Philox noise + dropout.
`cfg5_row_sharded`: configs[4] -- full-batch HMC softmax 1 000 000 x 2048 x 38, 8 chains replicated,
           rows sharded over the N ranks, one grouped NCCL all-reduce per gradient evaluation
           (STRONG scaling: the total work is fixed); with N > 1 the run asserts that the replicas
           stay identical and that a row-sharded gradient matches a 1-rank gradient.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: N, D, K, chains per GPU, eps, path_length
    "cfg2": dict(N=60000, D=784, K=10, C=64, eps=1e-4, path=1e-2, alpha=0.01,
                 desc="HMC softmax 60000x784x10 full batch, 64 chains/GPU"),
    # BASELINE configs[4] (rows shard over GPUs; the 1-GPU shapes are profiling targets of tools/profile_grad.py)
    "cfg5": dict(N=1000000, D=2048, K=38, C=8, eps=1e-7, path=1e-6, alpha=0.01,
                 desc="HMC softmax 1000000x2048x38 full batch, 8 chains"),
    "cfg5-half": dict(N=500000, D=2048, K=38, C=8, eps=1e-7, path=1e-6, alpha=0.01,
                      desc="HMC softmax 500000x2048x38 full batch, 8 chains (one of two row shards)"),
    "cfg2-small": dict(N=2048, D=784, K=10, C=16, eps=1e-4, path=2e-3, alpha=0.01,
                       desc="HMC softmax 2048x784x10 (debug size)"),
}


DATA_KIND = "dense"  # set from --data: "dense" = X ~ U[0,1) fp32; "pixels" = 8-bit pixels / 255 (what MNIST holds)


def _quantize(X):
    """U[0,1) -> k/255 with k uniform in 0..255: the value grid of MNIST images scaled as the reference's scripts do."""
    import torch
    return torch.div(torch.floor(X * 256.0).clamp_(max=255.0), torch.tensor(255.0, device=X.device))  # IEEE division


def use_all_host_threads():
    """The CPU arms use every host core the process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers,
    which pinned the NumPy oracle to one BLAS thread (measured: 6.3 instead of 27-35 grad-evals/s at --gpus 2)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def data_desc():
    return "synthetic" if DATA_KIND == "dense" else "synthetic (8-bit pixels / 255, MNIST's value grid)"


def synth(N, D, K, seed, device=None):
    """MNIST-shaped synthetic data: X ~ U[0,1) (or 8-bit pixels / 255), labels = argmax(X W* + Gumbel)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    Wt = torch.randn(D, K, generator=g) * 0.1
    if device is not None:
        gd = torch.Generator(device=device).manual_seed(seed)
        X = torch.rand(N, D, generator=gd, device=device, dtype=torch.float32)
        if DATA_KIND == "pixels":
            X = _quantize(X)
        u = torch.rand(N, K, generator=gd, device=device).clamp_(1e-12, 1 - 1e-7)
        y = (X @ Wt.to(device) - torch.log(-torch.log(u))).argmax(1).to(torch.int32)
        return X, y
    X = torch.rand(N, D, generator=g, dtype=torch.float32)
    if DATA_KIND == "pixels":
        X = _quantize(X)
    u = torch.rand(N, K, generator=g).clamp_(1e-12, 1 - 1e-7)
    y = (X @ Wt - torch.log(-torch.log(u))).argmax(1).to(torch.int32)
    return X, y


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, every 100 ms, in-process through NVML
    (nvidia_ml_py).  A polling `nvidia-smi -lms 100` subprocess was measured to stall kernel launches for 0.2-0.3 s
    at a time (stage21 logs), which lands in the timed region; it remains the fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []     # (sm_mhz, max_mhz, set of reasons)
        self.first = 0
        self.proc = None
        self._stop = False
        self.source = None

    # ---- NVML path ----
    def _nvml_loop(self, nv, h):
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(nv, a, None) or getattr(nv, b, 0)) for n, a, b in names]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(h))
                self.rows.append((sm, mx, {n for n, b in bits if b and (r & b)}))
            except Exception:
                pass
            time.sleep(0.1)

    # ---- nvidia-smi fallback ----
    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                reasons = {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8])
                           if v.lower().startswith("active")}
                self.rows.append((float(r[1]), float(r[2]), reasons))
            except Exception:
                pass

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.source = "nvml"
            threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True).start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._smi_loop, daemon=True).start()
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported (the sampler is started before
        the warm-up so that its initialisation is over by then)."""
        self.first = len(self.rows)

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        time.sleep(0.12)
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        rows = self.rows[self.first:]
        sm = [r[0] for r in rows]
        mx = [r[1] for r in rows]
        reasons = set().union(*[r[2] for r in rows]) if rows else set()
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.8), d.get("hbm_gbs", 6552.3), "measured"
    return 1400.0, 6650.0, "fallback"


def ncu_traffic(kernel, prec, wl, chains_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` captures (profiles/ncu_traffic.json).  Trajectories are ragged, so the timed launches carry fewer
    chains on average than a full launch: the captures are taken at several launch widths (chains per launch) and the
    figure for this run's mean width is interpolated linearly between the two nearest ones -- the traffic of these
    GEMMs is affine in the width (the X operand is streamed once whatever the width, the (P-Y)^T operand and the
    partials grow with it).  Returns (bytes, description)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        pts = sorted((e["chains_per_launch"], e["dram_bytes"], e["source"]) for k, e in t.items()
                     if k.startswith("%s/%s/%s" % (wl["desc"], prec, kernel)))
        if not pts:
            return None, "no ncu capture for this workload/precision"
        lo = max([q for q in pts if q[0] <= chains_per_launch] or [pts[0]], key=lambda q: q[0])
        hi = min([q for q in pts if q[0] >= chains_per_launch] or [pts[-1]], key=lambda q: q[0])
        if lo[0] == hi[0]:
            if abs(lo[0] - chains_per_launch) > 8:
                return None, "nearest ncu capture is of %d-chain launches; this run averaged %.1f" % (lo[0], chains_per_launch)
            return lo[1], "%s (captured at %d chains per launch; run mean %.1f)" % (lo[2], lo[0], chains_per_launch)
        w = (chains_per_launch - lo[0]) / (hi[0] - lo[0])
        return (1 - w) * lo[1] + w * hi[1], ("interpolated at the run's mean of %.1f chains per launch between the ncu "
                                            "captures at %d and %d chains (%s; %s)" % (chains_per_launch, lo[0], hi[0], lo[2], hi[2]))
    except Exception as e:
        return None, "ncu_traffic.json unreadable: %r" % (e,)


# ----------------------------------------------------------------------------------------------------
# ESS comparison workload ("small-N variant"): the reference's energy is the MEAN NLP while its dynamics
# use the SUM gradient, so at N = 60 000 the chain only moves for eps <~ 3e-7 and 500 CPU steps would take half an
# hour.  At N = 2 000 the oracle does a transition in ~30 ms, so BOTH arms can run >= 300 transitions with the same
# settings and the same estimator (Geyer initial positive sequence, dropout_hamiltonian_montecarlo_b200/ess.py).
ESS_SMALL = dict(N=2000, D=784, K=10, eps=1e-5, path=2e-4, alpha=0.01, burnin=50,
                 desc="HMC softmax 2000x784x10 full batch (small-N ESS variant), eps 1e-5, E[L] 20")


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return int(max([i.get("num_threads", 1) for i in threadpool_info()] + [1]))
    except Exception:
        return int(os.cpu_count() or 1)


def cpu_hmc_steps(wl, n_steps, warmup, L=None, seconds_target=None, blas_threads_cap=None):
    """Oracle port of hmc.step (hamiltonian/inference/cpu/hmc.py:39-64 + models/cpu/softmax.py) on the host cores:
    one chain, fp64, the workload's data shape.  L pins the path length of every step (bounded sample); None draws it
    as the reference does.  Returns (per-step seconds of the timed steps, grad evals of the timed steps)."""
    from oracle import hamiltonian_oracle as O
    if blas_threads_cap:
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=blas_threads_cap)
        except Exception:
            pass
    else:
        use_all_host_threads()
    X, y = synth(wl["N"], wl["D"], wl["K"], 0)
    X = X.numpy().astype(np.float64)
    Y = O.one_hot(y.numpy(), wl["K"])
    model = O.SoftmaxOracle({"alpha": wl["alpha"]})
    q = {"weights": np.zeros((wl["D"], wl["K"])), "bias": np.zeros(wl["K"])}
    rs = np.random.RandomState(0)
    per_step, n_grad = [], 0
    t_begin = time.perf_counter()
    for i in range(warmup + n_steps):
        t0 = time.perf_counter()
        u_len = (L - 0.5) * wl["eps"] / (2 * wl["path"]) if L else rs.rand()
        draws = O.TapeDraws([rs.normal(size=q["weights"].shape), rs.normal(size=q["bias"].shape)], [u_len, rs.rand()])
        r = O.hmc_step(model, q, ["weights", "bias"], wl["eps"], wl["path"], draws, X_train=X, y_train=Y)
        q = r["q"]
        if i >= warmup:
            per_step.append(time.perf_counter() - t0)
            n_grad += r["n_grad"]
            if seconds_target and time.perf_counter() - t_begin > seconds_target:
                break
    return per_step, n_grad


def cpu_ess_small(seconds_target=12.0, max_steps=400):
    """ESS/s of the CPU arm: the oracle port runs the small-N variant for up to max_steps transitions (bounded by
    seconds_target) and its samples go through the SAME estimator as the device run."""
    from oracle import hamiltonian_oracle as O
    from dropout_hamiltonian_montecarlo_b200.ess import ess as ess_fn
    use_all_host_threads()
    w = ESS_SMALL
    X, y = synth(w["N"], w["D"], w["K"], 0)
    X = X.numpy().astype(np.float64)
    Y = O.one_hot(y.numpy(), w["K"])
    model = O.SoftmaxOracle({"alpha": w["alpha"]})
    q = {"weights": np.zeros((w["D"], w["K"])), "bias": np.zeros(w["K"])}
    rs = np.random.RandomState(1)
    samples, acc, n_grad = [], [], 0
    t0 = None
    for i in range(w["burnin"] + max_steps):
        if i == w["burnin"]:
            t0 = time.perf_counter()
        draws = O.TapeDraws([rs.normal(size=q["weights"].shape), rs.normal(size=q["bias"].shape)], [rs.rand(), rs.rand()])
        r = O.hmc_step(model, q, ["weights", "bias"], w["eps"], w["path"], draws, X_train=X, y_train=Y)
        q = r["q"]
        if i >= w["burnin"]:
            samples.append(np.concatenate([q["weights"].ravel(), q["bias"].ravel()]))
            acc.append(r["accept_prob"])
            n_grad += r["n_grad"]
            if time.perf_counter() - t0 > seconds_target and len(samples) >= 100:
                break
    dt = time.perf_counter() - t0
    e = ess_fn(np.asarray(samples)[:, None, :], max_params=48)
    return {"ess_min_per_s": e["min"] / dt, "ess_median_per_s": e["median"] / dt, "steps": len(samples),
            "burnin": w["burnin"], "chains": 1, "mean_accept_prob": float(np.mean(acc)), "grad_evals_per_s": n_grad / dt,
            "seconds": dt, "cores": blas_threads(), "workload": w["desc"],
            "estimator": "Geyer initial positive sequence, 48 random parameters (same code as the device run)"}


def cpu_chains_over_processes(wl_name, L, seconds_target=10.0, procs=None):
    """The other way to use the host cores (what hamiltonian/inference/cpu/hmc_multicore.py intended):
    one single-threaded process per core, each running its own chain of the oracle port on the same workload.
    Whole-host rate = sum of the gradient evaluations / the longest process time.  Returns None if a worker failed."""
    try:
        procs = procs or len(os.sched_getaffinity(0))
    except Exception:
        procs = procs or (os.cpu_count() or 1)
    code = ("import json, sys; sys.path.insert(0, %r); import bench; bench.DATA_KIND = %r; "
            "ps, n = bench.cpu_hmc_steps(bench.WORKLOADS[%r], 1000, 1, L=%d, seconds_target=%r, blas_threads_cap=1); "
            "print(json.dumps([sum(ps), n]))" % (ROOT, DATA_KIND, wl_name, L, seconds_target))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    t0 = time.perf_counter()
    ws = [subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
          for _ in range(procs)]
    res = []
    for w in ws:
        try:
            out, _ = w.communicate(timeout=10 * seconds_target + 120)
            res.append(json.loads(out.strip().splitlines()[-1]))
        except Exception:
            w.kill()
            return None
    if not res or any(r[1] <= 0 for r in res):
        return None
    dt = max(r[0] for r in res)
    n = sum(r[1] for r in res)
    return {"value": n / dt, "unit": "grad-evals/s", "processes": procs, "threads_per_process": 1, "seconds": dt,
            "wall_s_incl_startup": time.perf_counter() - t0,
            "sample": "%d single-threaded processes, one chain each, L=%d, %.0f s of HMC steps per process (%d grad evals in all)"
                      % (procs, L, seconds_target, n)}


def cpu_reference_rate(wl, seconds_target=12.0, max_steps=20, wl_name="cfg2"):
    """cpu_baseline of the headline: bounded sample (path length pinned to L=11 per step) of the same workload, the host
    cores used both ways -- one chain on multi-threaded BLAS, and one single-threaded chain per core; `value` is the
    better of the two."""
    L = 11
    per_step, n_grad = cpu_hmc_steps(wl, max_steps, 0, L=L, seconds_target=seconds_target)
    dt = sum(per_step)
    out = dict(value=n_grad / dt, unit="grad-evals/s", cores=blas_threads(), kind="port",
               sample="%d HMC steps of 1 chain, L=%d (%d grad evals), fp64 NumPy oracle port of hmc.step, %s"
                      % (len(per_step), L, n_grad, wl["desc"]), seconds=dt)
    out["one_chain_blas_threads"] = {"value": out["value"], "threads": out["cores"]}
    mp = cpu_chains_over_processes(wl_name, L, seconds_target=min(10.0, seconds_target))
    if mp is not None:
        out["chains_over_processes"] = mp
        if mp["value"] > out["value"]:
            out.update(value=mp["value"], cores=mp["processes"], seconds=mp["seconds"],
                       sample=mp["sample"] + ", fp64 NumPy oracle port of hmc.step, " + wl["desc"])
    return out


def base_config(args, wl):
    """The part of `config` both arms share: what the workload IS (the arms differ in how much of it they sample)."""
    return {"workload": wl["desc"], "N": wl["N"], "D": wl["D"], "K": wl["K"], "chains_per_gpu": wl["C"],
            "step_size": wl["eps"], "path_length": wl["path"], "path_length_mode": args.path_mode,
            "sweep": "reference (Gauss-Seidel, 2 gradients per leapfrog iteration)"}


def run_reference(args, wl):
    """Reference arm: the reference's own CPU algorithm for the path (oracle port: the reference tree is not on the
    GPU box) on the box's host cores, every step a bounded sample of the workload: one chain, path length pinned to
    L = 6 (11 gradient evaluations per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = 6
    per_step, n_grad = cpu_hmc_steps(wl, args.steps, args.warmup, L=L)
    total = sum(per_step)
    threads = blas_threads()
    val = n_grad / total
    one_chain = {"value": val, "threads": threads}
    sample_how = "%d timed HMC steps of 1 chain on %d BLAS threads" % (len(per_step), threads)
    # the same cores as independent single-threaded chains (what the reference's hmc_multicore intended); the arm's value is
    # the better use of the host
    mp = cpu_chains_over_processes(args.workload, L, seconds_target=8.0)
    if mp is not None and mp["value"] > val:
        val, threads = mp["value"], mp["processes"]
        sample_how = mp["sample"]
    ess_cpu = None if args.no_ess else cpu_ess_small()
    line = {"impl": "reference", "metric": "grad evals/sec (chains x leapfrog) on MNIST-shape softmax BNN",
            "value": val, "unit": "grad-evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (1 + (L - 1) * 2) / val,  # one HMC step of one chain at the arm's whole-host rate
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": data_desc(), "config": base_config(args, wl),
            "cpu_baseline": {"value": val, "unit": "grad-evals/s", "cores": threads, "kind": "port",
                             "sample": "%s, path length pinned to L=%d (%d grad evals per step), "
                                       "fp64 NumPy oracle port (the reference tree is not present on the GPU box)"
                                       % (sample_how, L, 1 + (L - 1) * 2),
                             "one_chain_blas_threads": one_chain, "chains_over_processes": mp},
            "e2e": {"value": val, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if ess_cpu is not None:
        line["ess"] = ess_cpu
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# secondary blocks: BASELINE configs 3, 4, 5 (one JSON object each, merged into the headline line)
def _agg_time_count(dev, world, ms, count):
    """(max over ranks of the device time in ms, sum over ranks of the work count)."""
    import torch
    import torch.distributed as dist
    st = torch.tensor([ms, float(count)], dtype=torch.float64, device=dev)
    if world > 1:
        mx, sm = st.clone(), st.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return float(mx[0]), float(sm[1])
    return float(st[0]), float(st[1])


def _timed(world, fn):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return e0.elapsed_time(e1), out


def _group_roofline(ctx, run_once, flops_per_launch, peak_tf, names=("fwd", "bwd", "prep", "update"), persistent_flops=None):
    """Per-launch-group device times of ONE extra untimed pass with every kernel group bracketed by CUDA events
    (bracketing costs stream overlap, so it is never on inside a timed region): the dominant group, its mean
    duration, and the algorithmic rate of the GEMM group that took longest against the tensor peak."""
    ctx.timing(1)
    run_once()
    ctx.sync()
    t = [ctx.kernel_time(g) for g in range(4)]
    ctx.timing(0)
    groups = {names[g]: {"ms_total": t[g][0], "launch_groups": int(t[g][1]),
                         "avg_ms": (t[g][0] / t[g][1]) if t[g][1] else None} for g in range(4)}
    dom = max((0, 1), key=lambda g: t[g][0])
    avg = (t[dom][0] / t[dom][1]) * 1e-3 if t[dom][1] else 0.0
    flops, label = flops_per_launch[dom], names[dom]
    if persistent_flops is not None and t[1][1] == 0 and t[0][1] > 0:
        # the persistent minibatch kernel (csrc/softmax_persist.cuh): ONE launch per epoch holds the forward GEMMs, the
        # backward GEMMs and the updates of all its steps -- it is timed under the forward group and there is no
        # separate backward launch; its algorithmic work is both GEMMs of every step of the launch
        flops = persistent_flops / t[0][1]
        label = "k_sg_persistent2 (cta_group::2 forward + backward GEMM + update of every step of an epoch in one cooperative launch; k_sg_persistent where the shape has no pair form)"
    achieved = flops / avg / 1e12 if avg > 0 else 0.0
    return {"bound": "tensor", "kernel_group": label, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved / peak_tf, "avg_launch_ms": avg * 1e3, "algorithmic_flops_per_launch": flops,
            "groups": groups, "timing": "one extra pass, every kernel group bracketed by CUDA events on the launching stream"}


def bench_cfg3(ctx, h, X, y, rank, world, dev, prec, peak_tf, epochs=20, warm=4):
    """BASELINE configs[2]: SGLD and SGHMC softmax on sequential minibatches of 500, 128 chains per GPU (1024 chains
    over 8 GPUs; chains shard with no collective -> weak scaling).  sgld.py:31-46 / sghmc.py:19-39 + sgmcmc.py:40-86."""
    from dropout_hamiltonian_montecarlo_b200._lib import KIND, PREC
    from dropout_hamiltonian_montecarlo_b200.runtime import SamplerHandle
    N, D, K, B, C = h.N, h.D, h.K, 500, 128
    nb = N // B
    flops_gemm = 2.0 * B * D * K * C  # one GEMM of one minibatch evaluation of all chains
    out = {"workload": "SGLD / SGHMC softmax %dx%dx%d, minibatch %d, %d chains/GPU" % (N, D, K, B, C), "n_gpus": world,
           "scaling": "weak", "precision": prec, "unit": "grad-evals/s", "algorithmic_flops_per_chain_eval": 4.0 * B * D * K}
    # ---- SGLD: one library call runs whole epochs
    s = SamplerHandle(ctx, h, KIND["sgld"], C, seed=1, chain_id0=rank * C, precision=PREC[prec])
    s.set_q(np.zeros((C, h.P), np.float32))
    # warm-up: one burn-in and `warm` sampling epochs (sample store and per-epoch NLP touched; a 3 ms epoch is too short to
    # settle the clocks after the host-side gap between the blocks: single 5-epoch timings spread 24-32 us per step)
    s.sg_run(warm, 1, B, 1e-5, n_rows=N)
    l0 = ctx.launches
    ms, o = _timed(world, lambda: s.sg_run(epochs, 0, B, 1e-5, n_rows=N, step0=(warm + 1) * nb))
    launches = ctx.launches - l0
    ms, n = _agg_time_count(dev, world, ms, o["n_grad_evals"])
    roof = _group_roofline(ctx, lambda: s.sg_run(1, 0, B, 1e-5, n_rows=N, step0=nb * (epochs + warm + 1)),
                           {0: flops_gemm, 1: flops_gemm}, peak_tf, persistent_flops=2.0 * flops_gemm * nb)
    out["sgld"] = {"value": n / (ms * 1e-3), "ms_per_step": ms / (epochs * nb), "steps": epochs * nb,
                   "step": "one minibatch update of every chain (gradient + Philox noise + update)",
                   "gpu_launches_per_step": launches / (epochs * nb),
                   "algorithmic_tflops": n * 4.0 * B * D * K / (ms * 1e-3) / 1e12 / world, "roofline": roof}
    s.close()
    # ---- SGHMC: every minibatch is one sghmc.step (momentum draw, L-1 friction + noise iterations, Metropolis test)
    s = SamplerHandle(ctx, h, KIND["sghmc"], C, seed=1, chain_id0=rank * C, precision=PREC[prec], shared_path=True,
                      sghmc_descent=True)
    s.set_q(np.zeros((C, h.P), np.float32))
    eps, path, steps = 1e-5, 1e-4, 60
    for j in range(8):
        s.hmc_run(1, eps, path, row0=j * B, nrows=B, step0=j, keep_samples=False, keep_stats=False)

    def run(j0, n_steps):
        tot = 0
        for j in range(j0, j0 + n_steps):
            tot += s.hmc_run(1, eps, path, row0=(j % nb) * B, nrows=B, step0=j, keep_samples=False, keep_stats=False)["n_grad_evals"]
        return tot
    l0 = ctx.launches
    ms, n_local = _timed(world, lambda: run(8, steps))
    launches = ctx.launches - l0
    ms, n = _agg_time_count(dev, world, ms, n_local)
    roof = _group_roofline(ctx, lambda: run(8 + steps, 10), {0: flops_gemm, 1: flops_gemm}, peak_tf)
    out["sghmc"] = {"value": n / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
                    "step": "one sghmc.step of every chain on one minibatch, E[L] = 10, shared path lengths, descent sign",
                    "ms_per_grad_eval_all_chains": ms / (n_local / C), "gpu_launches_per_step": launches / steps,
                    "algorithmic_tflops": n * 4.0 * B * D * K / (ms * 1e-3) / 1e12 / world, "roofline": roof}
    s.close()
    out["value"] = out["sgld"]["value"]
    out["ms_per_step"] = out["sgld"]["ms_per_step"]
    out["roofline"] = {k: roof_v for k, roof_v in out["sgld"]["roofline"].items() if k != "groups"}
    return out


def bench_cfg4(ctx, X, y, rank, world, dev, prec, peak_tf, steps=20, C=16):
    """BASELINE configs[3]: SGHMC on the dropout MLP 784-512-512-10 (models/gpu/mlp.py:19-82), minibatch 500, Philox
    noise + dropout masks, chains shard over GPUs (weak scaling)."""
    from dropout_hamiltonian_montecarlo_b200._lib import KIND, PREC
    from dropout_hamiltonian_montecarlo_b200.runtime import MlpHandle, SamplerHandle
    N, D = X.shape
    K, B, n_mid = int(y.max().item()) + 1, 500, 512
    h = MlpHandle(ctx, N, D, n_mid, K, 0.01, 0.1, seed=3, chain_id0=rank * C)
    h.bind(X, y)
    s = SamplerHandle(ctx, h, KIND["sghmc"], C, seed=1, chain_id0=rank * C, precision=PREC[prec], sweep=[(0, h.P)],
                      shared_path=True, sghmc_descent=True)
    s.set_q(np.random.RandomState(0).normal(0, 0.05, (C, h.P)).astype(np.float32))
    eps, path = 1e-3, 5e-3  # E[L] = 5
    nb = N // B
    s.hmc_run(2, eps, path, row0=0, nrows=B, keep_samples=False)

    def run(i0, n_steps):
        tot = 0
        for i in range(i0, i0 + n_steps):
            tot += s.hmc_run(1, eps, path, row0=(i % nb) * B, nrows=B, step0=i, keep_samples=False)["n_grad_evals"]
        return tot
    l0 = ctx.launches
    ms, n_local = _timed(world, lambda: run(10, steps))
    launches = ctx.launches - l0
    ms, n = _agg_time_count(dev, world, ms, n_local)
    sw = D * n_mid + n_mid * n_mid + n_mid * K
    flops = 6.0 * B * sw - 2.0 * B * D * n_mid       # 1.606 GFLOP at B = 500
    f_fwd, f_bwd = 2.0 * B * sw * C, (4.0 * B * sw - 2.0 * B * D * n_mid) * C
    roof = _group_roofline(ctx, lambda: run(10 + steps, 5), {0: f_fwd, 1: f_bwd}, peak_tf,
                           names=("forward (3 GEMMs + epilogues)", "backward (5 GEMMs + reductions)", "operand splits", "update"))
    out = {"workload": "SGHMC MLP %d-%d-%d-%d dropout 0.1, minibatch %d, %d chains/GPU, joint sweep, E[L] = 5"
                       % (D, n_mid, n_mid, K, B, C), "n_gpus": world, "scaling": "weak", "precision": prec,
           "value": n / (ms * 1e-3), "unit": "grad-evals/s", "ms_per_step": ms / steps, "steps": steps,
           "step": "one sghmc.step of every chain on one minibatch", "gpu_launches_per_step": launches / steps,
           "algorithmic_flops_per_chain_eval": flops, "algorithmic_tflops": n * flops / (ms * 1e-3) / 1e12 / world,
           "roofline": roof, "parity": "pinned under a shim: the unmodified reference models/gpu/mlp.py runs under oracle/chainer_shim.py (Chainer's documented primitive semantics, torch.autograd); real Chainer / CuPy are not installable here"}
    s.close()
    h.close()
    return out


def _cfg4_both(ctx, X, y, rank, world, dev, prec, peak_tf):
    """The cfg4 block at 16 chains per GPU (its `value`; the width round 1 reported) with the same run at 64 chains per
    GPU beside it (SURVEY 8(d) gives 8-64 chains per GPU for this config: wider chain batches fill the machine better)."""
    out = bench_cfg4(ctx, X, y, rank, world, dev, prec, peak_tf)
    wide = bench_cfg4(ctx, X, y, rank, world, dev, prec, peak_tf, steps=8, C=64)
    out["chains_64"] = {k: wide[k] for k in ("workload", "value", "ms_per_step", "steps", "algorithmic_tflops")}
    out["chains_64"]["roofline_frac"] = wide["roofline"]["frac"]
    return out


def bench_cfg5_rows(ctx, rank, world, dev, prec, peak_tf, steps=4, rows=1000000, D=2048, K=38, C=8):
    """BASELINE configs[4]: full-batch HMC softmax, N = 1 M rows x 2048 features x 38 classes, 8 chains replicated on
    every rank, rows sharded N/G per rank, ONE grouped NCCL all-reduce (gradient + log-lik) per evaluation enqueued by
    the C driver.  Strong scaling: the total work is the same at every N."""
    import torch
    import torch.distributed as dist
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
    from dropout_hamiltonian_montecarlo_b200.parallel import shard_rows
    out = {"workload": "HMC softmax %dx%dx%d full batch, %d chains replicated, rows sharded over %d GPU(s)" % (rows, D, K, C, world),
           "n_gpus": world, "scaling": "strong", "precision": prec, "unit": "grad-evals/s",
           "collective": "none (1 rank)" if world == 1 else
           "ncclAllReduce(sum) of [%d x %d] fp32 + [%d] fp64 grouped into one call per evaluation, enqueued from C" % (C, (D + 1) * K, C)}
    # ---- in-run check (N > 1): a row-sharded gradient equals the 1-rank gradient of the same rows
    if world > 1:
        nchk = 65536
        g = torch.Generator(device=dev).manual_seed(7)
        Xc = torch.randn(nchk, D, generator=g, device=dev).abs_()
        yc = torch.randint(0, K, (nchk,), generator=g, device=dev, dtype=torch.int32)
        rs = np.random.RandomState(3)
        par = {"weights": rs.normal(0, 0.02, (C, D, K)).astype(np.float32), "bias": rs.normal(0, 0.02, (C, K)).astype(np.float32)}
        r0, nl = shard_rows(nchk, rank, world)
        m_sh = softmax({"alpha": 0.01}, precision=prec, row_sharded=True)
        g_sh = m_sh.grad(par, X_train=Xc[r0:r0 + nl].contiguous(), y_train=yc[r0:r0 + nl].contiguous())
        ll_sh = m_sh.log_likelihood(par, X_train=m_sh._bound[2][0], y_train=m_sh._bound[2][1])
        m_sh.unbind()
        err = torch.zeros(2, dtype=torch.float64, device=dev)
        if rank == 0:
            m_1 = softmax({"alpha": 0.01}, precision=prec)
            g_1 = m_1.grad(par, X_train=Xc, y_train=yc)
            ll_1 = m_1.log_likelihood(par, X_train=Xc, y_train=yc)
            m_1.unbind()
            err[0] = float(np.abs(g_sh["weights"] - g_1["weights"]).max() / np.abs(g_1["weights"]).max())
            err[1] = float(np.abs(ll_sh - ll_1).max() / np.abs(ll_1).max())
        dist.broadcast(err, src=0)
        out["check_grad_vs_1rank"] = {"rows": nchk, "max_rel_err_grad": float(err[0]), "max_rel_err_loglik": float(err[1]),
                                      "ok": bool(err[0] < 1e-4 and err[1] < 1e-9)}
        del Xc, yc
    # ---- the workload: every rank synthesises ITS shard (post-ReLU-like features abs(N(0,1)))
    r0, nloc = shard_rows(rows, rank, world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    Xs = torch.randn(nloc, D, generator=g, device=dev).abs_()
    ys = torch.randint(0, K, (nloc,), generator=g, device=dev, dtype=torch.int32)
    m = softmax({"alpha": 0.01}, precision=prec, row_sharded=world > 1)
    eps, path = 1e-7, 1e-6  # E[L] = 10; one path length shared by the chains (every launch carries all 8)
    smp = hmc(m, {"weights": np.zeros((D, K), np.float32), "bias": np.zeros(K, np.float32)}, path_length=path,
              step_size=eps, verbose=False, n_chains=C, seed=3, path_length_mode="shared")
    smp.sample(niter=1, burnin=0, X_train=Xs, y_train=ys)  # warm-up: binds (operand copies), allocates
    l0 = ctx.launches
    ms, res = _timed(world, lambda: smp.sample(niter=steps, burnin=0, X_train=Xs, y_train=ys))
    launches = ctx.launches - l0
    n_grad = smp.last_run["n_grad_evals"]  # identical on every rank (replicated chains)
    st = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
    ms = float(st[0])
    flops_eval = 4.0 * rows * D * K
    h = m._bound[1]
    f_gemm = 2.0 * nloc * D * K * C
    roof = _group_roofline(ctx, lambda: smp.sample(niter=1, burnin=0, X_train=Xs, y_train=ys), {0: f_gemm, 1: f_gemm}, peak_tf)
    out.update({"value": n_grad / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
                "step": "one HMC transition of the 8 chains (E[L] = 10: ~19 full-data gradient evaluations)",
                "ms_per_grad_eval_all_chains": ms / (n_grad / C), "rows_per_gpu": nloc, "gpu_launches_per_step": launches / steps,
                "algorithmic_flops_per_chain_eval": flops_eval, "algorithmic_tflops_total": n_grad * flops_eval / (ms * 1e-3) / 1e12,
                "allreduce_bytes_per_eval": 4 * C * h.ld + 8 * C, "roofline": roof})
    if world > 1:  # replicas must stay bit-identical: same draws, same all-reduced gradients
        post = res[0]
        t = torch.as_tensor(np.ascontiguousarray(post["weights"][-1])).to(dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        out["check_replicas_identical"] = bool(all(torch.equal(gathered[0], x) for x in gathered))
        out["check_moved"] = bool(np.abs(post["weights"][-1]).max() > 0)
    m.unbind()
    return out


# ----------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.inference.gpu.hmc import hmc
    from dropout_hamiltonian_montecarlo_b200.hamiltonian.models.gpu.softmax import softmax
    from dropout_hamiltonian_montecarlo_b200.runtime import SamplerHandle
    from dropout_hamiltonian_montecarlo_b200._lib import PREC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    N, D, K, C = wl["N"], wl["D"], wl["K"], wl["C"]
    prec = args.precision
    X, y = synth(N, D, K, 0, device=dev)
    model = softmax({"alpha": wl["alpha"]}, precision=prec)
    h = model.bind(X, y, n_classes=K)
    ctx = h.ctx
    shared = args.path_mode == "shared"
    s = SamplerHandle(ctx, h, 0, C, seed=1234, chain_id0=rank * C, precision=PREC[prec], shared_path=shared)
    s.set_q(np.zeros((C, h.P), np.float32))

    def run_steps(i0, n):
        """n HMC transitions of every chain in ONE library call (what hmc.sample(niter=n) issues): with per-chain path
        lengths the call runs the streaming schedule -- a chain starts its next transition as soon as its own
        trajectory ends -- so the K timed steps are K transitions per chain, not K barriers."""
        return s.hmc_run(n, wl["eps"], wl["path"], step0=i0, keep_samples=True, keep_stats=True, schedule=args.schedule)

    clocks = ClockSampler(local)
    if not os.environ.get("BENCH_NO_CLOCKS"):  # diagnosis only: a line without clocks is not a valid bench line
        clocks.start()
    warm_groups = None
    if args.warmup > 0:
        ctx.timing(1)  # every kernel group, for the share table (the timed region records the GEMM groups only)
        ow = run_steps(0, args.warmup)
        ctx.sync()
        tw = [ctx.kernel_time(g)[0] for g in range(4)]
        warm_groups = {"fwd": tw[0], "bwd": tw[1], "prep": tw[2], "update": tw[3],
                       "per": "warm-up region (%d steps, %d chain-evals), all kernel groups timed" % (args.warmup, ow["n_grad_evals"])}
        ctx.timing(0)
    ctx.sync()
    # L2 note: one gradient evaluation streams X (94-376 MB) + (P-Y)^T (77-245 MB) -- far larger than the
    # 126 MB L2 -- so consecutive launches cannot be served from cache; no explicit flush is needed.
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.mark()
    # live kernel timing for the roofline: CUDA events around every 8th forward / backward launch (every record
    # costs stream overlap; sampling keeps the perturbation of `value` below 1 %); the chains those launches carried
    # are counted with them, so flops / time is exact for the sample
    ctx.timing_stride(int(os.environ.get("BENCH_KTIMING_STRIDE", "8")))
    ctx.timing(0 if os.environ.get("BENCH_NO_KTIMING") else 2)
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_applied = n_launched = 0
    e0.record()
    o = run_steps(args.warmup, args.steps)
    n_applied += o["n_grad_evals"]
    n_launched += o["n_grad_launched"]
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t_fwd, n_fwd = ctx.kernel_time(0)
    t_bwd, n_bwd = ctx.kernel_time(1)
    u_fwd, u_bwd = ctx.kernel_units(0), ctx.kernel_units(1)
    ctx.timing_stride(1)
    ctx.timing(0)
    launches = ctx.launches - launches0
    clk = clocks.stop()
    n_launched_local = float(n_launched)  # chain-gradient evaluations this rank's GEMM launches processed
    stride = int(os.environ.get("BENCH_KTIMING_STRIDE", "8"))
    stats = torch.tensor([ms, float(n_applied), float(n_launched)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, n_applied, n_launched = float(mx[0]), float(sm[1]), float(sm[2])
    value = n_applied / (ms / 1e3)

    # ---- end to end through the public API with host buffers (rank-local; aggregated like value) ----
    e2e = None
    if not args.no_e2e:
        Xh = X.cpu().pin_memory()
        yh = y.cpu().pin_memory()
        api = hmc(softmax({"alpha": wl["alpha"]}, precision=prec, cache_data=False), {"weights": np.zeros((D, K), np.float32),
                                                                     "bias": np.zeros(K, np.float32)},
                  path_length=wl["path"], step_size=wl["eps"], verbose=False, n_chains=C, seed=99,
                  chain_id0=rank * C, path_length_mode=args.path_mode)
        n_e2e = 0
        d2h = 0
        for i in range(1 + args.steps):  # first call = warm-up (allocations)
            if i == 1:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
            # cache_data=False: every call uploads X / y from pinned HOST memory (H2D + bf16 operand
            # preparation inside the timed region) and returns samples / losses to the host (D2H)
            post, loss, _, _ = api.sample(niter=1, burnin=0, X_train=Xh, y_train=yh)
            if i >= 1:
                n_e2e += api.last_run["n_grad_evals"]
                d2h = post["weights"].size * 4 + post["bias"].size * 4 + loss.size * 8 * 2 + loss.size * 4
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = torch.tensor([dt, float(n_e2e)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = st.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = st.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            dt, n_e2e = float(mx[0]), float(sm[1])
        e2e = {"value": n_e2e / dt, "unit": "grad-evals/s", "h2d_bytes_per_step": int(N * D * 4 + N * 4),
               "d2h_bytes_per_step": int(d2h), "api": "hmc.sample(niter=1, X_train=<pinned host>, y_train=<pinned host>)"}

    # ---- ESS / s (second half of BASELINE's metric), Geyer initial-positive-sequence estimator.  Two runs:
    # (1) the headline data at settings where proposals are accepted (the sum-gradient / mean-energy
    #     mismatch of the reference makes the chain move only for eps <~ 3e-7 at N = 60 000), >= 500 transitions;
    # (2) the small-N variant (ESS_SMALL) that the CPU arm can also run for hundreds of transitions: both arms, same
    #     settings, same estimator -> `ess.small_n` here, `cpu_baseline.ess` / the reference arm's `ess` there.
    ess_info = None
    if not args.no_ess:
        from dropout_hamiltonian_montecarlo_b200.ess import ess as ess_fn

        def device_ess(hh, e_eps, e_path, n_steps, n_burn, seed):
            s2 = SamplerHandle(ctx, hh, 0, C, seed=seed, chain_id0=rank * C, precision=PREC[prec], shared_path=shared)
            s2.set_q(np.zeros((C, hh.P), np.float32))
            s2.hmc_run(n_burn, e_eps, e_path, step0=0, keep_samples=False)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            oe = s2.hmc_run(n_steps, e_eps, e_path, step0=n_burn, keep_samples=True)
            ev1.record()
            torch.cuda.synchronize()
            t_ess = ev0.elapsed_time(ev1) * 1e-3
            idx = np.sort(np.random.RandomState(0).choice(hh.P, 48, replace=False))  # the estimator runs on 48 parameters
            sub = oe["samples"][:, :, torch.as_tensor(idx, device=dev)].cpu().numpy()
            r = ess_fn(sub, max_params=48)
            st = torch.tensor([t_ess, r["min"], r["median"], float(oe["accept_prob"].mean().item()), float(oe["n_grad_evals"])],
                              dtype=torch.float64, device=dev)
            if world > 1:
                mx = st.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                sm = st.clone()
                dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                st = torch.stack([mx[0], sm[1], sm[2], sm[3] / world, sm[4]])
            s2.close()
            return {"ess_min_per_s": float(st[1] / st[0]), "ess_median_per_s": float(st[2] / st[0]), "steps": n_steps,
                    "burnin": n_burn, "chains": C * world, "step_size": e_eps, "path_length": e_path,
                    "mean_accept_prob": float(st[3]), "grad_evals_per_s": float(st[4] / st[0]), "seconds": float(st[0]),
                    "estimator": "Geyer initial positive sequence per chain, summed over chains, 48 random parameters"}

        ess_info = device_ess(h, args.ess_eps, args.ess_eps * args.ess_L, args.ess_steps, args.ess_burnin, 4321)
        ws = ESS_SMALL
        Xs_, ys_ = synth(ws["N"], ws["D"], ws["K"], 0, device=dev)
        ms_ = softmax({"alpha": ws["alpha"]}, precision=prec)
        hs_ = ms_.bind(Xs_, ys_, n_classes=ws["K"])
        ess_info["small_n"] = dict(device_ess(hs_, ws["eps"], ws["path"], 400, ws["burnin"], 777), workload=ws["desc"])
        ms_.unbind()
        del Xs_, ys_

    # ---- same workload on MNIST's value grid (8-bit pixels / 255): the bind-time check finds X exact in bf16 after
    # scaling by 255 and bf16x3 issues 2 MMAs per product instead of 3.  Reported beside the headline, not as it.
    pixel_info = None
    if DATA_KIND == "dense" and not args.no_pixels and N * D * 4 <= (2 << 30):
        Xp = _quantize(X)
        mp = softmax({"alpha": wl["alpha"]}, precision=prec)
        hp = mp.bind(Xp, y, n_classes=K)
        sp = SamplerHandle(ctx, hp, 0, C, seed=1234, chain_id0=rank * C, precision=PREC[prec], shared_path=shared)
        sp.set_q(np.zeros((C, hp.P), np.float32))
        sp.hmc_run(args.warmup, wl["eps"], wl["path"], step0=0, keep_samples=True, keep_stats=True, schedule=args.schedule)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        op = sp.hmc_run(args.steps, wl["eps"], wl["path"], step0=args.warmup, keep_samples=True, keep_stats=True,
                        schedule=args.schedule)
        p1.record()
        torch.cuda.synchronize()
        st = torch.tensor([p0.elapsed_time(p1), float(op["n_grad_evals"])], dtype=torch.float64, device=dev)
        if world > 1:
            mx = st.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = st.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            st = torch.stack([mx[0], sm[1]])
        pex, psc = hp.operand_info()
        pixel_info = {"value": float(st[1] / (st[0] * 1e-3)), "unit": "grad-evals/s", "ms_per_step": float(st[0]) / max(1, args.steps),
                      "data": "same workload, X = 8-bit pixels / 255 (MNIST's value grid)",
                      "x_operand": ("exact in bf16 after scaling by %g: 2 MMAs per product" % psc) if pex else "not exact"}
        del sp, hp, mp, Xp

    # ---- BASELINE configs 3, 4, 5 as secondary blocks of the same line.  A failure in one of them must not lose the
    # headline: it is reported in the block instead ("error").
    secondary = {}
    peak_tf_all = peaks()[0]
    want = set(args.blocks.split(",")) if args.blocks else set()
    for name, fn in (("cfg3", lambda: bench_cfg3(ctx, h, X, y, rank, world, dev, prec, peak_tf_all)),
                     ("cfg4", lambda: _cfg4_both(ctx, X, y, rank, world, dev, prec, peak_tf_all)),
                     ("cfg5_row_sharded", lambda: bench_cfg5_rows(ctx, rank, world, dev, prec, peak_tf_all,
                                                                  rows=args.cfg5_rows))):
        if name not in want:
            continue
        try:
            t0_ = time.perf_counter()
            blk = fn()
            blk["block_wall_s"] = time.perf_counter() - t0_
            secondary[name] = blk
        except Exception as e:  # noqa: BLE001 -- the block is reported as failed, the line still prints
            import traceback
            traceback.print_exc()
            secondary[name] = {"error": repr(e)[:400]}
            if world > 1:
                raise  # ranks would deadlock in the next collective if only one of them failed
        torch.cuda.empty_cache()

    if rank == 0:
        peak_tf, peak_bw, src = peaks()
        # Algorithmic work of the dominant GEMM over the timed region: 2*N*D*K flops per chain-gradient evaluation
        # and per GEMM (4*N*D*K for forward + backward).  Ragged trajectories make the launches
        # process between 1 and C chains, so the figure is total flops / total kernel time, not a full-launch
        # figure divided by the mean duration.
        flops_eval = 2.0 * N * D * K
        dom = "fwd" if t_fwd >= t_bwd else "bwd"
        t_dom, n_dom, u_dom = (t_fwd, n_fwd, u_fwd) if dom == "fwd" else (t_bwd, n_bwd, u_bwd)
        avg = (t_dom / max(1, n_dom)) * 1e-3
        flops_launch = flops_eval * u_dom / max(1, n_dom)   # chains carried by the bracketed launches
        achieved = flops_launch / avg / 1e12 if avg > 0 else 0.0
        traffic, traffic_src = ncu_traffic(dom, prec, wl, u_dom / max(1, n_dom))
        x_exact, x_scale = h.operand_info()
        mma_mult = (2.0 if x_exact else 3.0) if prec == "bf16x3" else 1.0
        # the CPU leg runs at N = 1 only (rank 0 alone owns the host cores there; at N > 1 the other ranks would wait for it)
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_reference_rate(wl)
        if cpu is not None and not args.no_ess:
            cpu["ess"] = cpu_ess_small()
        line = {
            "metric": "grad evals/sec (chains x leapfrog) on MNIST-shape softmax BNN",
            "value": value, "unit": "grad-evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (bf16 hi/lo split x3 on tcgen05, fp32 accumulate)" if prec == "bf16x3" else prec,
            "data": data_desc(),
            "config": dict(base_config(args, wl), **{
                       "precision": prec,
                       "schedule": ("streaming (asynchronous chains, %d gradient launches)" % o["n_phases"]) if o["n_phases"]
                       else "lockstep",
                       "l2": "inputs larger than L2 (X 94-376 MB + (P-Y)^T 77-245 MB per evaluation)",
                       "grad_evals_launched_incl_masked": n_launched,
                       "x_operand": ("exact in bf16 after scaling by %g: 2 MMAs per product" % x_scale) if x_exact
                       else "fp32 values: hi/lo split, 3 MMAs per product"}),
            "roofline": {"bound": "tensor", "kernel": "backward GEMM X^T (P-Y) + split-K reduce (k_tc_bwd_sk: swapped operand roles, interleaved stream-K, "
                         "cta_group::2; k_tc_gemm<bwd> for launches of few chains)" if dom == "bwd"
                         else "forward GEMM (X W + softmax epilogue)", "kernel_group": dom, "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": src, "algorithmic_flops_per_launch": flops_launch,
                         "chains_per_launch_mean": u_dom / max(1, n_dom), "avg_launch_ms": avg * 1e3,
                         "launches_timed": int(n_dom), "timing": "CUDA events around every %d-th launch of the GEMM groups "
                         "inside the timed region" % stride,
                         "mma_flops_issued_over_algorithmic": mma_mult,
                         "group_ms": {"fwd_sampled": t_fwd, "bwd_sampled": t_bwd, "step_total": ms},
                         "warmup_group_ms": warm_groups},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if ess_info is not None:
            line["ess"] = ess_info
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if pixel_info is not None:
            line["pixel_data"] = pixel_info
        for k_, v_ in secondary.items():
            line[k_] = v_
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--path-mode", default="per_chain", choices=["per_chain", "shared"])
    ap.add_argument("--schedule", default="auto", choices=["auto", "lockstep", "streaming"])
    ap.add_argument("--data", default="dense", choices=["dense", "pixels"],
                    help="dense: X ~ U[0,1) fp32; pixels: 8-bit pixels / 255 (MNIST's value grid; exact-operand path)")
    ap.add_argument("--no-pixels", action="store_true", help="skip the secondary run on 8-bit pixel data")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ess", action="store_true")
    ap.add_argument("--ess-steps", type=int, default=500)
    ap.add_argument("--ess-burnin", type=int, default=50)
    ap.add_argument("--blocks", default="cfg3,cfg4,cfg5_row_sharded",
                    help="secondary blocks to run after the headline (comma separated; empty = none)")
    ap.add_argument("--cfg5-rows", type=int, default=1000000, help="total rows of the row-sharded workload")
    # ESS settings: swept on B200 (gpurun_out/stage19.log): eps in {5e-8..3e-7} x E[L] in {100, 30}; the shorter
    # trajectories give ~3x the ESS/s (195-226 vs 63-72) because a transition costs 1/3 and accepts more often
    ap.add_argument("--ess-eps", type=float, default=2e-7)
    ap.add_argument("--ess-L", type=float, default=30.0, help="path_length / step_size of the ESS run (E[L])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global DATA_KIND
    DATA_KIND = args.data
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
