"""GPU parity of csrc/gridenc.cu and csrc/freq_sh.cu through the drop-in packages,
against the CPU oracle and the reference's CUDA extensions (oracle/_ref).

Bars: fp16 hash-grid outputs bit-exact vs the reference kernel (same fp16 rounding sequence);
fp32 outputs rtol 1e-5; table gradients vs an fp64 scatter (oracle) rtol 1e-3 of the max (atomic order);
freq: atol 2e-6 vs reference kernel (__sinf on both), 5e-4 vs libm oracle at 2^5 x; SH: rtol 2e-5."""
import numpy as np
import pytest
import torch

from helpers import ref
from oracle import oracle as O
from sdf_b200 import _lib

pytestmark = pytest.mark.gpu


def T(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return t if dtype is None else t.to(dtype)


def device_resolutions(offsets_t, S, H):
    L = offsets_t.shape[0] - 1
    out = np.zeros(L, np.uint32)
    _lib.call("sdf_grid_level_resolutions", _lib.ptr(offsets_t), L, float(S), int(H), out.ctypes.data, _lib.stream())
    return out


GRID_CASES = [
    # D, C, L, log2_hash, desired_res, gridtype, align_corners, interp, max_level_frac
    (3, 2, 16, 19, 2048, 0, False, 1, None),     # the -O backbone configuration (nerf/network_grid.py:49)
    (3, 2, 16, 19, 2048, 0, False, 0, 0.5),
    (3, 2, 8, 14, 256, 1, True, 0, None),
    (2, 4, 6, 12, 128, 0, False, 1, None),
    (3, 1, 5, 10, 64, 0, True, 1, None),
    (3, 8, 4, 12, 64, 1, False, 0, None),
]


@pytest.mark.parametrize("case", GRID_CASES, ids=[f"D{c[0]}C{c[1]}L{c[2]}h{c[3]}g{c[5]}a{int(c[6])}i{c[7]}" for c in GRID_CASES])
@pytest.mark.parametrize("half", [False, True])
def test_grid_encode_forward_backward(device, case, half):
    import gridencoder
    D, C, L, log2h, dres, gridtype, ac, interp, mlf = case
    if half and C % 2:
        pytest.skip("fp16 path is only taken for even C (grid.py:46)")
    offsets, pls = O.grid_offsets(D, L, C, 2.0, 16, log2h, dres)
    n = int(offsets[-1])
    rng = np.random.default_rng(0)
    B = 20011
    x = rng.random((B, D), dtype=np.float32)
    x[:50] = rng.random((50, D), dtype=np.float32) * 1.2 - 0.1        # some out-of-range points -> zeros
    x[50:60] = 0.0; x[60:70] = 1.0                                        # exact borders
    table = (rng.random((n, C), dtype=np.float32) - 0.5)
    S = np.float32(np.log2(pls))
    t_x, t_tab, t_off = T(x, device), T(table, device), T(offsets, device)
    res_dev = device_resolutions(t_off, S, 16)
    res_host = np.array([O.grid_resolution(l, S, 16) for l in range(L)], np.uint32)
    if not np.array_equal(res_dev, res_host):
        print("NOTE device exp2f resolutions differ from host:", res_dev, res_host)
    import math
    max_level = L if mlf is None else max(min(int(math.ceil(mlf * L)), L), 1)

    t_x.requires_grad_(True)
    emb = t_tab.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16, enabled=half):
        out = gridencoder.grid_encode(t_x, emb, t_off, pls, 16, True, gridtype, ac, interp, mlf)
    assert out.dtype == (torch.float16 if half else torch.float32) and out.shape == (B, L * C)
    o_ref, dd_ref = O.grid_encode_forward(x, table, offsets, pls, 16, True, gridtype, ac, interp, max_level, half, res_dev)
    o_ref = o_ref.transpose(1, 0, 2).reshape(B, L * C)
    got = out.detach().float().cpu().numpy()
    if half:
        assert np.array_equal(got, o_ref), np.abs(got - o_ref).max()
    else:
        np.testing.assert_allclose(got, o_ref, rtol=1e-5, atol=1e-6)

    g = rng.normal(size=(B, L * C)).astype(np.float32)
    out.backward(T(g, device).to(out.dtype))
    g_used = g.astype(np.float16).astype(np.float32) if half else g
    g_LBC = g_used.reshape(B, L, C).transpose(1, 0, 2)
    gg_ref = O.grid_encode_backward(g_LBC, x, offsets, n, C, pls, 16, gridtype, ac, interp, max_level, False, res_dev)
    gg = emb.grad.double().cpu().numpy()
    scale = np.abs(gg_ref).max()
    assert np.abs(gg - gg_ref).max() / scale < 1e-3, np.abs(gg - gg_ref).max() / scale
    # input gradient through dy_dx
    gi_ref = O.grid_input_backward(g_LBC, dd_ref, B, D, C, L, half)
    gi = t_x.grad.cpu().numpy()
    tol = 2e-2 if half else 1e-4
    assert np.abs(gi - gi_ref).max() / (np.abs(gi_ref).max() + 1e-9) < tol

    r = ref.load("_gridencoder")
    if r is not None:
        dt = torch.float16 if half else torch.float32
        tab_r = t_tab.to(dt)
        out_r = torch.zeros(L, B, C, device=device, dtype=dt)
        dd_r = torch.zeros(B, L * D * C, device=device, dtype=dt)
        r.grid_encode_forward(T(x, device), tab_r, t_off, out_r, B, D, C, L, max_level, float(S), 16, dd_r, gridtype, ac, interp)
        torch.cuda.synchronize()
        out_r2 = out_r.permute(1, 0, 2).reshape(B, L * C).float().cpu().numpy()
        if half:
            assert np.array_equal(got, out_r2), np.abs(got - out_r2).max()
        else:
            np.testing.assert_allclose(got, out_r2, rtol=1e-6, atol=1e-7)
        # oracle vs reference kernel
        if half:
            assert np.array_equal(o_ref, out_r2)
        g_r = T(g, device).to(dt).view(B, L, C).permute(1, 0, 2).contiguous()
        gg_r = torch.zeros(n, C, device=device, dtype=dt)
        gi_r = torch.zeros(B, D, device=device, dtype=dt)
        r.grid_encode_backward(g_r, T(x, device), tab_r, t_off, gg_r, B, D, C, L, max_level, float(S), 16, dd_r, gi_r, gridtype, ac, interp)
        torch.cuda.synchronize()
        tolg = 3e-2 if half else 1e-3       # the reference accumulates fp16 with half2 atomics
        assert np.abs(gg_r.double().cpu().numpy() - gg_ref).max() / scale < tolg
        assert np.abs(gi_r.float().cpu().numpy() - gi).max() / (np.abs(gi).max() + 1e-9) < tol


def test_grid_module_autocast_cache_and_tv_wd(device):
    import gridencoder
    torch.manual_seed(0)
    enc = gridencoder.GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                  desired_resolution=2048, interpolation='smoothstep').to(device)
    with torch.no_grad():
        enc.embeddings.uniform_(-0.5, 0.5)
    x = torch.rand(4099, 3, device=device) * 2 - 1
    with torch.autocast("cuda", dtype=torch.float16):
        y1 = enc(x, bound=1)
        y2 = enc(x, bound=1)
    assert y1.dtype == torch.float16 and torch.equal(y1, y2)
    y32 = enc(x, bound=1)
    assert y32.dtype == torch.float32
    assert (y1.float() - y32).abs().max() < 4e-3
    with torch.no_grad():
        enc.embeddings.data.mul_(2.0)   # a .data write (no version bump) must still be seen by the fp16 path
    with torch.autocast("cuda", dtype=torch.float16):
        y3 = enc(x, bound=1)
    assert (y3.float() - 2 * y32).abs().max() < 8e-3
    # TV / WD injectors vs oracle
    y32 = enc(x, bound=1); y32.sum().backward()
    g0 = enc.embeddings.grad.clone()
    offsets = enc.offsets.cpu().numpy(); tab = enc.embeddings.detach().cpu().numpy()
    res_dev = device_resolutions(enc.offsets, np.float32(np.log2(enc.per_level_scale)), 16)
    xin = torch.rand(3001, 3, device=device) * 2 - 1
    enc.grad_total_variation(1e-3, xin, bound=1)
    tv = O.grad_total_variation(((xin + 1) / 2).cpu().numpy(), tab, offsets, 1e-3, enc.per_level_scale, 16, 0, False, res_dev)
    got = (enc.embeddings.grad - g0).double().cpu().numpy()
    assert np.abs(got - tv).max() / (np.abs(tv).max() + 1e-12) < 1e-3
    g1 = enc.embeddings.grad.clone()
    enc.grad_weight_decay(0.1)
    wd = O.grad_weight_decay(tab, g1.cpu().numpy(), offsets, 0.1)
    np.testing.assert_allclose(enc.embeddings.grad.cpu().numpy(), wd, rtol=1e-5, atol=1e-9)
    r = ref.load("_gridencoder")
    if r is not None:
        g_r = g0.clone()
        r.grad_total_variation(((xin + 1) / 2).contiguous(), enc.embeddings.detach(), g_r, enc.offsets, 1e-3, 3001, 3, 2, 16,
                               float(np.log2(enc.per_level_scale)), 16, 0, False)
        torch.cuda.synchronize()
        assert ((g_r - g0).double().cpu().numpy() - tv).__abs__().max() / (np.abs(tv).max() + 1e-12) < 1e-3


@pytest.mark.parametrize("degree", [4, 6, 12])
def test_freq_encode(device, degree):
    import freqencoder
    rng = np.random.default_rng(0)
    x = (rng.random((5003, 3), dtype=np.float32) * 2 - 1)
    t = T(x, device).requires_grad_(True)
    enc = freqencoder.FreqEncoder(3, degree)
    y = enc(t)
    yo = O.freq_encode_forward(x, degree)
    # __sinf absolute error grows with |argument| (2^(deg-1) * |x|)
    atol = 5e-7 * 2 ** degree + 2e-6
    np.testing.assert_allclose(y.detach().cpu().numpy(), yo, rtol=0, atol=atol)
    g = rng.normal(size=yo.shape).astype(np.float32)
    y.backward(T(g, device))
    go = O.freq_encode_backward(g, y.detach().cpu().numpy(), 3, degree)      # backward uses the saved outputs
    np.testing.assert_allclose(t.grad.cpu().numpy(), go, rtol=2e-4, atol=1e-4)      # fp32 sum of 2^f-scaled terms, FMA order differs
    r = ref.load("_freqencoder")
    if r is not None:
        y2 = torch.empty_like(y); r.freq_encode_forward(T(x, device), 5003, 3, degree, yo.shape[1], y2)
        gi2 = torch.zeros(5003, 3, device=device); r.freq_encode_backward(T(g, device), y2, 5003, 3, degree, yo.shape[1], gi2)
        torch.cuda.synchronize()
        assert torch.equal(y.detach(), y2)
        np.testing.assert_allclose(t.grad.cpu().numpy(), gi2.cpu().numpy(), rtol=2e-4, atol=1e-4)


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_encode(device, degree):
    import shencoder
    rng = np.random.default_rng(degree)
    x = rng.normal(size=(3001, 3)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x[:100] *= rng.random((100, 1), dtype=np.float32)         # points off the unit sphere: polynomials, not angles
    t = T(x, device).requires_grad_(True)
    y = shencoder.SHEncoder(3, degree)(t)
    yo, ddo = O.sh_encode_forward(x, degree, True)
    # fp32 evaluation of degree-(l) polynomials with coefficients up to ~2e2: cancellation error grows with the degree
    atol = 2e-6 * 4 ** max(0, degree - 4)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yo, rtol=5e-5, atol=atol)
    g = rng.normal(size=yo.shape).astype(np.float32)
    y.backward(T(g, device))
    go = O.sh_encode_backward(g, ddo, 3, degree)
    np.testing.assert_allclose(t.grad.cpu().numpy(), go, rtol=1e-3, atol=1e-3 * degree)
    r = ref.load("_shencoder")
    if r is not None:
        y2 = torch.empty_like(y); dd2 = torch.empty(3001, 3 * degree * degree, device=device)
        r.sh_encode_forward(T(x, device), y2, 3001, 3, degree, dd2)
        gi2 = torch.zeros(3001, 3, device=device)
        r.sh_encode_backward(T(g, device), T(x, device), 3001, 3, degree, dd2, gi2)
        torch.cuda.synchronize()
        np.testing.assert_allclose(y.detach().cpu().numpy(), y2.cpu().numpy(), rtol=5e-5, atol=atol)
        np.testing.assert_allclose(ddo, dd2.cpu().numpy(), rtol=1e-4, atol=1e-4 * 4 ** max(0, degree - 4))      # oracle dy_dx vs reference tables
        np.testing.assert_allclose(t.grad.cpu().numpy(), gi2.cpu().numpy(), rtol=1e-3, atol=1e-3 * degree)
