"""bench.py on the CPU box: the reference arm (oracle port on the host cores) runs without a GPU and prints ONE JSON line
with the keys the driver reads; the synthetic data generators have the documented shapes and value grids."""
import json
import os
import subprocess
import sys

import numpy as np

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports to its workers
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2-small",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "grad-evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    assert cb["cores"] == avail  # all host threads, in spite of OMP_NUM_THREADS=1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2-small",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_synthetic_data():
    X, y = bench.synth(300, 20, 5, 0)
    assert tuple(X.shape) == (300, 20) and tuple(y.shape) == (300,)
    assert float(X.min()) >= 0.0 and float(X.max()) < 1.0 and int(y.min()) >= 0 and int(y.max()) < 5
    X2, y2 = bench.synth(300, 20, 5, 0)
    assert np.array_equal(X.numpy(), X2.numpy()) and np.array_equal(y.numpy(), y2.numpy())  # seeded
    old = bench.DATA_KIND
    try:
        bench.DATA_KIND = "pixels"
        Xp, _ = bench.synth(300, 20, 5, 0)
        k = np.rint(Xp.numpy() * np.float32(255))
        assert np.array_equal(Xp.numpy(), (k / np.float32(255)).astype(np.float32)) and "pixels" in bench.data_desc()
    finally:
        bench.DATA_KIND = old
    assert bench.data_desc() == "synthetic"
