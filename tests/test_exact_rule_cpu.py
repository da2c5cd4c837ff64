"""Host-side statement of the bind-time exact-operand rule (csrc/softmax_tc.cu:k_detect_exact): which inputs let the
bf16x3 GEMMs drop the lo copy of X.  The CUDA kernel is tested on the GPU (test_gpu_parity.py); here the rule itself is
pinned against the ways 8-bit pixels get scaled in practice, and bench.py's pixel generator is checked to satisfy it."""
import numpy as np
import torch

import bench


def bf16_round(x):
    """round-to-nearest-even fp32 -> bf16 -> fp32"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def rule(x):
    """(exact, scale) as tc_softmax_bind decides it"""
    x = np.asarray(x, np.float32)
    if np.all(bf16_round(x) == x):
        return True, 1.0
    k = np.rint(x * np.float32(255.0))
    ok = (k >= 0) & (k <= 255) & ((x == k / np.float32(255.0)) | (x == (k.astype(np.float64) / 255.0).astype(np.float32)) |
                                 (x == k * (np.float32(1.0) / np.float32(255.0))))
    return (True, 255.0) if np.all(ok) else (False, 1.0)


def test_rule_accepts_every_pixel_scaling():
    k = np.arange(256)
    assert rule(k / 255.0) == (True, 255.0)                                        # float64 quotient, then cast
    assert rule(k.astype(np.float32) / np.float32(255)) == (True, 255.0)           # fp32 quotient
    assert rule(k.astype(np.float32) * (np.float32(1) / np.float32(255))) == (True, 255.0)  # reciprocal multiply
    # the operand the GEMM sees is bf16(255 * x) = k exactly, for all three
    for x in (k / 255.0, k.astype(np.float32) * (np.float32(1) / np.float32(255))):
        x = np.asarray(x, np.float32)
        assert np.array_equal(bf16_round(x * np.float32(255.0)), k.astype(np.float32))


def test_rule_scale_one_and_rejections():
    assert rule(np.array([0.0, 1.0, -2.0, 0.5, 96.0])) == (True, 1.0)
    x = np.arange(256) / 255.0
    x[17] += 1e-6
    assert rule(x) == (False, 1.0)
    assert rule(np.random.RandomState(0).rand(1000)) == (False, 1.0)
    assert rule(np.array([0.5, 256.0 / 255.0, 0.3])) == (False, 1.0)               # k out of range / off the grid
    assert rule(np.array([np.nan, 0.5])) == (False, 1.0)


def test_bench_pixel_generator_is_on_the_grid():
    x = bench._quantize(torch.rand(4096, generator=torch.Generator().manual_seed(0))).numpy()
    assert rule(x) == (True, 255.0)
    assert x.min() >= 0.0 and x.max() <= 1.0 and len(np.unique(x)) == 256
