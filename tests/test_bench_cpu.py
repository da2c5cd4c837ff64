"""CPU: host-side pieces of bench.py -- both arms describe the same workload, the CPU-arm ESS uses the shared
estimator, the traffic interpolation reads the committed ncu captures."""
import argparse
import json
import os

import numpy as np

import bench


def _args():
    return argparse.Namespace(path_mode="per_chain")


def test_both_arms_share_the_workload_config():
    wl = bench.WORKLOADS["cfg2"]
    c = bench.base_config(_args(), wl)
    assert c["workload"] == wl["desc"] and c["N"] == 60000 and c["D"] == 784 and c["K"] == 10 and c["chains_per_gpu"] == 64
    assert c["step_size"] == wl["eps"] and c["path_length"] == wl["path"]


def test_cpu_arm_steps_and_grad_count():
    wl = dict(bench.WORKLOADS["cfg2-small"], N=256, D=20, K=4)
    per_step, n_grad = bench.cpu_hmc_steps(wl, 3, 1, L=4)
    assert len(per_step) == 3 and n_grad == 3 * (1 + 3 * 2)  # hmc.py:47-54: 1 + (L-1) * n_vars evaluations per step


def test_cpu_ess_small_runs_the_shared_estimator(monkeypatch):
    monkeypatch.setitem(bench.ESS_SMALL, "N", 200)
    monkeypatch.setitem(bench.ESS_SMALL, "D", 12)
    monkeypatch.setitem(bench.ESS_SMALL, "K", 3)
    monkeypatch.setitem(bench.ESS_SMALL, "burnin", 5)
    r = bench.cpu_ess_small(seconds_target=0.5, max_steps=120)
    assert r["steps"] >= 100 and r["chains"] == 1 and r["ess_min_per_s"] >= 0 and r["ess_median_per_s"] >= r["ess_min_per_s"]
    assert 0.0 <= r["mean_accept_prob"] <= 1.0 and r["grad_evals_per_s"] > 0


def test_ncu_traffic_interpolates_between_captured_widths(tmp_path, monkeypatch):
    wl = bench.WORKLOADS["cfg2"]
    prof = tmp_path / "profiles"
    prof.mkdir()
    key = "%s/bf16x3/bwd" % wl["desc"]
    json.dump({key + "@64": {"chains_per_launch": 64, "dram_bytes": 400e6, "source": "a"},
               key + "@32": {"chains_per_launch": 32, "dram_bytes": 300e6, "source": "b"}}, open(prof / "ncu_traffic.json", "w"))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    t, src = bench.ncu_traffic("bwd", "bf16x3", wl, 48.0)
    assert abs(t - 350e6) < 1 and "interpolated" in src
    t, src = bench.ncu_traffic("bwd", "bf16x3", wl, 64.0)
    assert t == 400e6
    t, src = bench.ncu_traffic("fwd", "bf16x3", wl, 64.0)
    assert t is None


def test_committed_traffic_file_is_readable():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    t = json.load(open(p))
    for k, e in t.items():
        assert {"chains_per_launch", "dram_bytes", "source"} <= set(e), k


def test_reference_arm_under_torchrun_only_rank0_works():
    """Contract of the reference arm at N > 1: rank 0 alone runs and prints, the other ranks exit 0 without work."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
