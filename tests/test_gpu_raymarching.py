"""GPU parity of csrc/raymarch.cu, called through the drop-in `raymarching` package (C ABI underneath),
against (1) the CPU oracle oracle/sdf_oracle.c and (2) the reference's own CUDA extension oracle/_ref/_raymarching.so.

Bars: bit-exact for counts, offsets (canonical order), xyzs/dirs/ts sample lists, morton, packbits, flatten;
rtol 1e-4 / atol 1e-5 for composited image/depth/weights and gradients (__expf + reduction order)."""
import numpy as np
import pytest
import torch

from helpers import ref, scenes
from oracle import oracle as O
from sdf_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def T(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return t if dtype is None else t.to(dtype)


def canon(rays, *arrs):
    """Re-pack per-sample arrays so that ray n's samples sit at the exclusive prefix sum of counts (ray order)."""
    rays = rays.cpu().numpy()
    cnt = rays[:, 1].astype(np.int64)
    off_new = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    idx = np.concatenate([np.arange(o, o + c) for o, c in zip(rays[:, 0], cnt)]) if cnt.sum() else np.zeros(0, np.int64)
    outs = [a.cpu().numpy()[idx] for a in arrs]
    r2 = np.stack([off_new, cnt], 1).astype(np.int32)
    return r2, outs


def test_near_far_sph(device):
    import raymarching
    for seed, bound in [(0, 1.0), (1, 2.0)]:
        ro, rd, aabb, nears, fars, _ = scenes.make_rays(48, 48, bound, 40.0, seed)
        n, f = raymarching.near_far_from_aabb(T(ro, device), T(rd, device), T(aabb, device), 0.2)
        assert np.array_equal(n.cpu().numpy(), nears) and np.array_equal(f.cpu().numpy(), fars)
        r = ref.load("_raymarching")
        if r is not None:
            n2, f2 = torch.empty_like(n), torch.empty_like(f)
            r.near_far_from_aabb(T(ro, device), T(rd, device), T(aabb, device), ro.shape[0], 0.2, n2, f2)
            torch.cuda.synchronize()
            assert torch.equal(n, n2) and torch.equal(f, f2)
        ro_in = ro * 0.1   # origins inside the sphere
        c = raymarching.sph_from_ray(T(ro_in, device), T(rd, device), 1.4)
        np.testing.assert_allclose(c.cpu().numpy(), O.sph_from_ray(ro_in, rd, 1.4), rtol=1e-5, atol=1e-5)
        if r is not None:
            c2 = torch.empty_like(c)
            r.sph_from_ray(T(ro_in, device), T(rd, device), 1.4, ro.shape[0], c2)
            torch.cuda.synchronize()
            np.testing.assert_allclose(c.cpu().numpy(), c2.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_morton_packbits_flatten(device):
    import raymarching
    rng = np.random.default_rng(0)
    coords = rng.integers(0, 128, (100003, 3), dtype=np.int32)
    ind = raymarching.morton3D(T(coords, device))
    assert np.array_equal(ind.cpu().numpy(), O.morton3D(coords))
    back = raymarching.morton3D_invert(ind)
    assert np.array_equal(back.cpu().numpy(), coords)
    grid = rng.random((2, 128 ** 3), dtype=np.float32)
    bits = raymarching.packbits(T(grid, device), 0.37)
    assert np.array_equal(bits.cpu().numpy(), O.packbits(grid, 0.37))
    # empty / ragged rays
    cnt = rng.integers(0, 70, 5000).astype(np.int32)
    cnt[::7] = 0
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
    rays = np.stack([off, cnt], 1)
    M = int(cnt.sum())
    res = raymarching.flatten_rays(T(rays, device), M)
    assert np.array_equal(res.cpu().numpy(), O.flatten_rays(rays, M))
    r = ref.load("_raymarching")
    if r is not None:
        ind2 = torch.empty_like(ind); r.morton3D(T(coords, device), coords.shape[0], ind2)
        bits2 = torch.empty_like(bits); r.packbits(T(grid, device), bits.numel(), 0.37, bits2)
        res2 = torch.zeros_like(res); r.flatten_rays(T(rays, device), rays.shape[0], M, res2)
        torch.cuda.synchronize()
        assert torch.equal(ind, ind2) and torch.equal(bits, bits2) and torch.equal(res, res2)


@pytest.mark.parametrize("case", scenes.MARCH_CASES, ids=[f"{c[0]}-b{c[1]}-g{c[3]:.4f}-c{int(c[5])}" for c in scenes.MARCH_CASES])
def test_march_rays_train_bit_exact(device, case):
    import raymarching
    kind, bound, cas, dtg, max_steps, contract, fovy = case
    bf = synth.occupancy_bitfield(kind, 128, cas, bound, seed=1)
    ro, rd, aabb, nears, fars, noises = scenes.make_rays(40, 40, bound, fovy, seed=3)
    N = ro.shape[0]
    # oracle
    xo, do, to, ro_rays = O.march_rays_train(ro, rd, bound, bf, cas, 128, nears, fars, noises, dtg, max_steps, contract)
    # ours, through the C ABI directly so the same noises are used
    from sdf_b200 import _lib
    d = lambda a: T(a, device)
    t_ro, t_rd, t_bf, t_n, t_f, t_nz = d(ro), d(rd), d(bf), d(nears), d(fars), d(noises)
    rays = torch.empty(N, 2, dtype=torch.int32, device=device)
    counter = torch.zeros(1, dtype=torch.int32, device=device)
    args = (_lib.ptr(t_ro), _lib.ptr(t_rd), _lib.ptr(t_bf), bound, int(contract), dtg, max_steps, N, cas, 128, _lib.ptr(t_n), _lib.ptr(t_f), _lib.ptr(t_nz))
    _lib.call("sdf_march_rays_train_count", *args, _lib.ptr(rays), _lib.ptr(counter), None, _lib.stream())
    M = int(counter.item())
    assert M == int(ro_rays[:, 1].sum())
    assert np.array_equal(rays.cpu().numpy(), ro_rays)
    xyzs = torch.zeros(M, 3, device=device); dirs = torch.zeros(M, 3, device=device); ts = torch.zeros(M, 2, device=device)
    _lib.call("sdf_march_rays_train_write", *args, _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ts), _lib.ptr(rays), M, _lib.stream())
    assert np.array_equal(xyzs.cpu().numpy(), xo) and np.array_equal(dirs.cpu().numpy(), do) and np.array_equal(ts.cpu().numpy(), to)

    # reference extension (atomic offsets -> canonicalise)
    r = ref.load("_raymarching")
    if r is not None:
        rays2 = torch.empty(N, 2, dtype=torch.int32, device=device)
        cnt2 = torch.zeros(1, dtype=torch.int32, device=device)
        r.march_rays_train(t_ro, t_rd, t_bf, bound, contract, dtg, max_steps, N, cas, 128, t_n, t_f, None, None, None, rays2, cnt2, t_nz)
        M2 = int(cnt2.item())
        assert M2 == M
        x2 = torch.zeros(M2, 3, device=device); d2 = torch.zeros(M2, 3, device=device); s2 = torch.zeros(M2, 2, device=device)
        r.march_rays_train(t_ro, t_rd, t_bf, bound, contract, dtg, max_steps, N, cas, 128, t_n, t_f, x2, d2, s2, rays2, cnt2, t_nz)
        torch.cuda.synchronize()
        rc, (xc, dc, sc) = canon(rays2, x2, d2, s2)
        assert np.array_equal(rc, rays.cpu().numpy())
        assert np.array_equal(xc, xyzs.cpu().numpy()) and np.array_equal(dc, dirs.cpu().numpy()) and np.array_equal(sc, ts.cpu().numpy())


def test_march_wrapper_and_empty(device):
    import raymarching
    bf = synth.occupancy_bitfield("blob", 128, 1, 1.0, seed=1)
    ro, rd, aabb, nears, fars, _ = scenes.make_rays(32, 32, 1.0, 20.0, seed=5, default_view=True)
    d = lambda a: T(a, device)
    xyzs, dirs, ts, rays = raymarching.march_rays_train(d(ro), d(rd), 1.0, d(bf), 1, 128, d(nears), d(fars), False, 0, 1024)
    xo, do, to, ro_rays = O.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, None, 0.0, 1024)
    assert np.array_equal(rays.cpu().numpy(), ro_rays) and np.array_equal(xyzs.cpu().numpy(), xo) and np.array_equal(ts.cpu().numpy(), to)
    # all rays miss
    bf0 = synth.occupancy_bitfield("empty", 128, 1, 1.0)
    x, dd, t, r = raymarching.march_rays_train(d(ro), d(rd), 1.0, d(bf0), 1, 128, d(nears), d(fars))
    assert x.shape[0] == 0 and int(r[:, 1].sum()) == 0
    w, ws, dep, img = raymarching.composite_rays_train(torch.zeros(0, device=device), torch.zeros(0, 3, device=device), t, r)
    assert float(ws.abs().sum()) == 0 and float(img.abs().sum()) == 0


@pytest.mark.parametrize("binarize", [False, True])
def test_composite_train_fwd_bwd(device, binarize):
    import raymarching
    bf = synth.occupancy_bitfield("blob", 128, 1, 1.0, seed=1)
    ro, rd, aabb, nears, fars, noises = scenes.make_rays(40, 40, 1.0, 20.0, seed=7)
    xo, do, to, rays = O.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, noises, 0.0, 1024)
    M, N = xo.shape[0], ro.shape[0]
    rng = np.random.default_rng(1)
    # densities that terminate some rays early and leave others translucent
    sig = (np.exp(rng.normal(2.0, 2.0, M))).astype(np.float32)
    rgb = rng.random((M, 3), dtype=np.float32)
    d = lambda a: T(a, device)
    s_t = d(sig).requires_grad_(True); c_t = d(rgb).requires_grad_(True)
    w, ws, dep, img = raymarching.composite_rays_train(s_t, c_t, d(to), d(rays), 1e-4, binarize)
    wo, wso, depo, imgo = O.composite_rays_train_forward(sig, rgb, to, rays, 1e-4, binarize)
    tol = dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(ws.detach().cpu().numpy(), wso, **tol)
    np.testing.assert_allclose(img.detach().cpu().numpy(), imgo, **tol)
    np.testing.assert_allclose(dep.detach().cpu().numpy(), depo, rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(w.detach().cpu().numpy(), wo, rtol=2e-4, atol=2e-5)
    gw = rng.normal(size=M).astype(np.float32) * 0.1
    gws = rng.normal(size=N).astype(np.float32)
    gd = rng.normal(size=N).astype(np.float32) * 0.1
    gi = rng.normal(size=(N, 3)).astype(np.float32)
    torch.autograd.backward([w, ws, dep, img], [d(gw), d(gws), d(gd), d(gi)])
    gso, gro = O.composite_rays_train_backward(gw, gws, gd, gi, sig, rgb, to, rays, wso, depo, imgo, 1e-4, binarize)
    gs = s_t.grad.cpu().numpy(); gr = c_t.grad.cpu().numpy()
    np.testing.assert_allclose(gr, gro, rtol=2e-4, atol=2e-5)
    scale = np.abs(gso).max() + 1e-6
    assert np.abs(gs - gso).max() / scale < 2e-4, np.abs(gs - gso).max() / scale
    r = ref.load("_raymarching")
    if r is not None:
        w2 = torch.zeros(M, device=device); ws2 = torch.empty(N, device=device); dep2 = torch.empty(N, device=device); img2 = torch.empty(N, 3, device=device)
        r.composite_rays_train_forward(d(sig), d(rgb), d(to), d(rays), M, N, 1e-4, binarize, w2, ws2, dep2, img2)
        gs2 = torch.zeros(M, device=device); gr2 = torch.zeros(M, 3, device=device)
        r.composite_rays_train_backward(d(gw), d(gws), d(gd), d(gi), d(sig), d(rgb), d(to), d(rays), ws2, dep2, img2, M, N, 1e-4, binarize, gs2, gr2)
        torch.cuda.synchronize()
        np.testing.assert_allclose(img.detach().cpu().numpy(), img2.cpu().numpy(), **tol)
        np.testing.assert_allclose(w.detach().cpu().numpy(), w2.cpu().numpy(), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(gr, gr2.cpu().numpy(), rtol=2e-4, atol=2e-5)
        assert np.abs(gs - gs2.cpu().numpy()).max() / scale < 2e-4
        # the oracle restatement agrees with the reference kernel too
        np.testing.assert_allclose(imgo, img2.cpu().numpy(), **tol)
        assert np.abs(gso - gs2.cpu().numpy()).max() / scale < 2e-4


@pytest.mark.parametrize("perturb", [False, True])
def test_inference_loop(device, perturb):
    """march_rays + composite_rays driven like nerf/renderer.py:759-794: ours, the CPU oracle AND the reference's own extension
    (oracle/_ref/_raymarching.so: raymarching.cu:714-829, :843-925) step by step on identical state — sample positions / ts bit-exact,
    alive lists equal, accumulated image / weights_sum / depth within 2e-4."""
    import raymarching
    rm = ref.load("_raymarching")
    bf = synth.occupancy_bitfield("blob", 128, 1, 1.0, seed=1)
    ro, rd, aabb, nears, fars, _ = scenes.make_rays(24, 24, 1.0, 20.0, seed=9)
    N = ro.shape[0]
    d = lambda a: T(a, device)
    rng = np.random.default_rng(2)
    t_ro, t_rd, t_bf, t_n, t_f = d(ro), d(rd), d(bf), d(nears), d(fars)
    ws = torch.zeros(N, device=device); dep = torch.zeros(N, device=device); img = torch.zeros(N, 3, device=device)
    ws_r = torch.zeros(N, device=device); dep_r = torch.zeros(N, device=device); img_r = torch.zeros(N, 3, device=device)
    ws_o = np.zeros(N, np.float32); dep_o = np.zeros(N, np.float32); img_o = np.zeros((N, 3), np.float32)
    alive = torch.arange(N, dtype=torch.int32, device=device); rt = t_n.clone()
    alive_r = alive.clone(); rt_r = t_n.clone()
    alive_o = np.arange(N, dtype=np.int32); rt_o = nears.copy()
    step = 0
    while step < 1024:
        n_alive = alive.shape[0]
        assert n_alive == alive_o.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        noises = rng.random(n_alive, dtype=np.float32) if (perturb and step == 0) else None
        M = n_alive * n_step
        x = torch.zeros(M, 3, device=device); dd = torch.zeros(M, 3, device=device); t = torch.zeros(M, 2, device=device)
        _lib.call("sdf_march_rays", n_alive, n_step, alive.data_ptr(), rt.data_ptr(), t_ro.data_ptr(), t_rd.data_ptr(), 1.0, 0, 0.0, 1024, 1, 128,
                  t_bf.data_ptr(), t_n.data_ptr(), t_f.data_ptr(), x.data_ptr(), dd.data_ptr(), t.data_ptr(), None if noises is None else d(noises).data_ptr(),
                  _lib.stream())
        xo, do, to = O.march_rays(n_alive, n_step, alive_o, rt_o, ro, rd, 1.0, bf, 1, 128, nears, fars, noises, 0.0, 1024)
        assert np.array_equal(x.cpu().numpy(), xo) and np.array_equal(t.cpu().numpy(), to)
        if rm is not None:
            xr = torch.zeros(M, 3, device=device); dr = torch.zeros(M, 3, device=device); tr = torch.zeros(M, 2, device=device)
            nz = d(noises) if noises is not None else torch.zeros(n_alive, device=device)
            rm.march_rays(n_alive, n_step, alive_r, rt_r, t_ro, t_rd, 1.0, False, 0.0, 1024, 1, 128, t_bf, t_n, t_f, xr, dr, tr, nz)
            assert torch.equal(xr, x) and torch.equal(tr, t) and torch.equal(dr, dd), "march_rays differs from the reference extension"
        sig = np.exp(rng.normal(1.0, 2.0, M)).astype(np.float32)
        rgb = rng.random((M, 3), dtype=np.float32)
        raymarching.composite_rays(n_alive, n_step, alive, rt, d(sig), d(rgb), t, ws, dep, img, 1e-2)
        O.composite_rays(n_alive, n_step, alive_o, rt_o, sig, rgb, to, ws_o, dep_o, img_o, 1e-2)
        assert np.array_equal(alive.cpu().numpy(), alive_o)
        if rm is not None:
            rm.composite_rays(n_alive, n_step, 1e-2, False, alive_r, rt_r, d(sig), d(rgb), tr, ws_r, dep_r, img_r)
            assert torch.equal(alive_r, alive), "composite_rays kills a different set of rays than the reference extension"
            np.testing.assert_allclose(rt.cpu().numpy(), rt_r.cpu().numpy(), rtol=0, atol=0)
            alive_r = alive_r[alive_r >= 0]
        alive = alive[alive >= 0]; alive_o = alive_o[alive_o >= 0]
        step += n_step
    np.testing.assert_allclose(img.cpu().numpy(), img_o, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(ws.cpu().numpy(), ws_o, rtol=2e-4, atol=2e-5)
    if rm is not None:
        np.testing.assert_allclose(img.cpu().numpy(), img_r.cpu().numpy(), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(ws.cpu().numpy(), ws_r.cpu().numpy(), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(dep.cpu().numpy(), dep_r.cpu().numpy(), rtol=2e-4, atol=2e-5)


def test_device_side_inference_loop_matches_host_driven_loop(device):
    """sdf_infer_begin / _march / _composite / _compact (loop state and alive-ray compaction on the device, no host sync) against the
    host-driven reference protocol above on the same synthetic densities: same image, weights_sum, depth (rays are independent of their slot)."""
    bf = synth.occupancy_bitfield("blob", 128, 1, 1.0, seed=1)
    ro, rd, aabb, nears, fars, _ = scenes.make_rays(32, 32, 1.0, 20.0, seed=10)
    N = ro.shape[0]
    d = lambda a: T(a, device)
    t_ro, t_rd, t_bf, t_n, t_f = d(ro), d(rd), d(bf), d(nears), d(fars)
    P = lambda t: t.data_ptr()

    def field(x):           # a smooth synthetic density / colour of the sample position, evaluated identically in both loops
        r2 = (x * x).sum(-1)
        return (40.0 * torch.exp(-r2 / 0.08)).contiguous(), (0.5 + 0.5 * torch.sin(7.0 * x)).contiguous()

    # host-driven loop (reference protocol)
    import raymarching
    ws = torch.zeros(N, device=device); dep = torch.zeros(N, device=device); img = torch.zeros(N, 3, device=device)
    alive = torch.arange(N, dtype=torch.int32, device=device); rt = t_n.clone()
    step = 0
    while step < 1024 and alive.shape[0] > 0:
        n_alive = alive.shape[0]
        n_step = max(min(N // n_alive, 8), 1)
        x, dd, t = raymarching.march_rays(n_alive, n_step, alive, rt, t_ro, t_rd, 1.0, t_bf, 1, 128, t_n, t_f, False, 0, 1024)
        s, c = field(x)
        raymarching.composite_rays(n_alive, n_step, alive, rt, s, c, t, ws, dep, img, 1e-4)
        alive = alive[alive >= 0]
        step += n_step
    # device-driven loop
    state = torch.zeros(8, dtype=torch.int32, device=device)
    al = [torch.empty(N, dtype=torch.int32, device=device) for _ in range(2)]
    rt2 = torch.empty(N, device=device)
    ws2 = torch.empty(N, device=device); dep2 = torch.empty(N, device=device); img2 = torch.empty(N, 3, device=device)
    xs = torch.empty(N, 3, device=device); ds = torch.empty(N, 3, device=device); ts = torch.empty(N, 2, device=device)
    st = _lib.stream()
    _lib.call("sdf_infer_begin", P(state), N, 1024, P(al[0]), P(rt2), P(t_n), P(ws2), P(dep2), P(img2), None, st)
    cur, iters = 0, 0
    for it in range(1024):
        _lib.call("sdf_infer_march", P(state), N, P(al[cur]), P(rt2), P(t_ro), P(t_rd), 1.0, 0, 0.0, 1024, 1, 128, P(t_bf), P(t_f), P(xs), P(ds), P(ts), None, st)
        s, c = field(xs)                    # capacity-sized: rows >= M are ignored by the compositor
        _lib.call("sdf_infer_composite", P(state), N, 1e-4, 0, P(al[cur]), P(rt2), P(s), P(c), P(ts), P(ws2), P(dep2), P(img2), st)
        _lib.call("sdf_infer_compact", P(state), N, P(al[cur]), P(al[1 - cur]), None, st)
        cur = 1 - cur
        iters += 1
        if it % 16 == 15 and int(state[0].item()) == 0:
            break
    stt = state.cpu().numpy()
    assert stt[0] == 0, f"rays still alive after {iters} iterations: {stt}"
    np.testing.assert_allclose(img2.cpu().numpy(), img.cpu().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ws2.cpu().numpy(), ws.cpu().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(dep2.cpu().numpy(), dep.cpu().numpy(), rtol=2e-5, atol=2e-6)
