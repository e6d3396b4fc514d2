"""GPU: the fused background kernels (csrc/render_aux.cu) against torch autograd of the same network — freq_encode(d, 6) -> Linear 39-32 + ReLU ->
Linear 32-3 -> sigmoid (nerf/network_grid.py:141-147) mixed as image + (1 - weights_sum) * bg (nerf/renderer.py:796-808) — at a ray count large
enough that every block of the backward walks several 64-ray chunks (the 512x512 DMTet stage) and at a ragged small one.
Tolerance: 2e-3 absolute on values (fast sine in the encoding), 2e-3 relative L2 on gradients."""
import math

import pytest
import torch

from sdf_b200 import _lib

pytestmark = pytest.mark.gpu
P = _lib.ptr


def torch_bg(d, w1, b1, w2, b2):
    enc = [d]
    for f in range(6):
        enc += [torch.sin(d * 2.0 ** f), torch.cos(d * 2.0 ** f)]
    h = torch.relu(torch.cat(enc, -1) @ w1.t() + b1)
    return torch.sigmoid(h @ w2.t() + b2)


@pytest.mark.parametrize("N,HW", [(70001, 70001), (389, 389)])
def test_background_forward_backward(device, N, HW):
    g = torch.Generator(device=device).manual_seed(N)
    d = torch.nn.functional.normalize(torch.randn(N, 3, device=device, generator=g), dim=-1)
    w1 = (torch.randn(32, 39, device=device, generator=g) * 0.3).requires_grad_(True)
    b1 = (torch.randn(32, device=device, generator=g) * 0.1).requires_grad_(True)
    w2 = (torch.randn(3, 32, device=device, generator=g) * 0.3).requires_grad_(True)
    b2 = (torch.randn(3, device=device, generator=g) * 0.1).requires_grad_(True)
    image_c = torch.rand(N, 3, device=device, generator=g).requires_grad_(True)
    ws = torch.rand(N, device=device, generator=g).requires_grad_(True)
    bg, pred = torch.empty(N, 3, device=device), torch.empty(1, 3, HW, device=device)
    st = _lib.stream()
    _lib.call("sdf_background_forward", P(d), N, P(w1), P(b1), P(w2), P(b2), None, 0, P(image_c), P(ws), P(bg), None, P(pred), HW, 3, st)
    ref_bg = torch_bg(d, w1, b1, w2, b2)
    ref = image_c + (1 - ws)[:, None] * ref_bg
    assert (bg - ref_bg).abs().max() < 2e-3 and (pred[0].t() - ref).abs().max() < 2e-3
    G = torch.randn(1, 3, HW, device=device, generator=g)
    (ref * G[0].t()).sum().backward()
    g_ic, g_ws = torch.empty(N, 3, device=device), torch.empty(N, device=device)
    gw = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
    _lib.call("sdf_background_backward", None, P(G), HW, 3, P(d), N, P(w1), P(b1), P(w2), P(b2), None, 0, P(ws.detach()), P(g_ic), P(g_ws), *[P(t) for t in gw], st)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(g_ic, image_c.grad) < 1e-6 and rel(g_ws, ws.grad) < 2e-3
    for mine, t in zip(gw, (w1, b1, w2, b2)):
        assert rel(mine, t.grad) < 2e-3, rel(mine, t.grad)
