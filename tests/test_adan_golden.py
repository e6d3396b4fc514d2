"""Adan against tests/golden/adan.npz — parameters produced by the reference's own optimizer.py (foreach=False) on CPU
(tests/golden/make_golden_adan.py).  CPU: the fp64 restatement the GPU tests use as their checker reproduces the reference.
GPU: the fused kernel (csrc/adan.cu through sdf_b200.optimizer.Adan) reproduces it too, including the clipped step."""
import os

import numpy as np
import pytest
import torch

PATH = os.path.join(os.path.dirname(__file__), "golden", "adan.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/adan.npz not generated")


def _load():
    g = np.load(PATH)
    n = len([k for k in g.files if k.startswith("p0.")])
    steps = len([k for k in g.files if k.endswith(".0") and k.startswith("g")])
    return g, n, steps


def test_restatement_matches_reference_optimizer():
    from test_gpu_adan import ref_adan_steps
    g, n, steps = _load()
    p0 = [torch.from_numpy(g[f"p0.{i}"]) for i in range(n)]
    grads = [[torch.from_numpy(g[f"g{s}.{i}"]) for i in range(n)] for s in range(steps)]
    for upto in range(1, steps + 1):
        got = ref_adan_steps(p0, grads[:upto], [float(x) for x in g["lrs"]])
        for i in range(n):
            ref = g[f"p{upto}.{i}"].astype(np.float64)
            assert np.abs(got[i].numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (upto, i)


@pytest.mark.gpu
def test_fused_kernel_matches_reference_optimizer(device):
    from sdf_b200.optimizer import Adan
    g, n, steps = _load()
    params = [torch.nn.Parameter(torch.from_numpy(g[f"p0.{i}"]).to(device)) for i in range(n)]
    opt = Adan([{"params": [p], "lr": float(lr)} for p, lr in zip(params, g["lrs"])], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
    for s in range(steps):
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(g[f"g{s}.{i}"]).to(device)
        opt.step()
        for i, p in enumerate(params):
            ref = g[f"p{s + 1}.{i}"]
            err = np.abs(p.detach().cpu().numpy() - ref).max()
            assert err <= 2e-5 * max(1.0, np.abs(ref).max()), (s, i, err)


def test_reference_arm_optimizer_matches_reference_optimizer():
    """oracle/ref_cuda_path.TorchAdan (the optimizer of both reference arms of bench.py) against the same golden trajectory"""
    from oracle.ref_cuda_path import TorchAdan
    g, n, steps = _load()
    params = [torch.nn.Parameter(torch.from_numpy(g[f"p0.{i}"]).clone()) for i in range(n)]
    opt = TorchAdan([{"params": [p], "lr": float(lr)} for p, lr in zip(params, g["lrs"])], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0)
    for s in range(steps):
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(g[f"g{s}.{i}"]).clone()
        opt.step()
        for i, p in enumerate(params):
            ref = g[f"p{s + 1}.{i}"]
            assert np.abs(p.detach().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (s, i)
