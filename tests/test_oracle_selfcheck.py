"""CPU (T0): closed-form self-checks of the oracle restatement."""
import numpy as np

from oracle import oracle as O
from sdf_b200 import synth
from helpers import scenes


def test_morton_roundtrip_and_numpy():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (20000, 3), dtype=np.int32)
    m = O.morton3D(c)
    assert np.array_equal(O.morton3D_invert(m), c)
    assert np.array_equal(m.astype(np.uint32), synth.morton3d_np(c[:, 0], c[:, 1], c[:, 2]))
    # interleave definition
    x, y, z = 5, 3, 6   # 101, 011, 110 -> bits (z y x) per position: pos0: z0 y1 x1 = 0 1 1, pos1: 1 1 0, pos2: 1 0 1
    assert int(O.morton3D(np.array([[x, y, z]], np.int32))[0]) == (0b011) | (0b110 << 3) | (0b101 << 6)


def test_packbits_vs_numpy():
    rng = np.random.default_rng(1)
    g = rng.random(8 * 4097, dtype=np.float32)
    assert np.array_equal(O.packbits(g, 0.5), np.packbits(g.reshape(-1, 8) > 0.5, axis=1, bitorder="little").reshape(-1))
    assert np.array_equal(O.packbits(g, 0.5), synth.pack_bitfield(g, 0.5))


def test_sh_orthonormal_on_sphere():
    # Gauss-Legendre x uniform-phi quadrature integrates degree<=14 products exactly enough
    nz, nphi = 32, 64
    zs, wz = np.polynomial.legendre.leggauss(nz)
    phis = (np.arange(nphi) + 0.5) * 2 * np.pi / nphi
    Z, P = np.meshgrid(zs, phis, indexing="ij")
    s = np.sqrt(1 - Z ** 2)
    pts = np.stack([s * np.cos(P), s * np.sin(P), Z], -1).reshape(-1, 3).astype(np.float32)
    w = (wz[:, None] * (2 * np.pi / nphi) * np.ones_like(P)).reshape(-1)
    Y, _ = O.sh_encode_forward(pts, 8)
    G = (Y.astype(np.float64) * w[:, None]).T @ Y.astype(np.float64)
    assert np.abs(G - np.eye(64)).max() < 2e-5


def test_sh_dy_dx_is_the_gradient():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(50, 3)).astype(np.float32) * 0.7
    _, dd = O.sh_encode_forward(x, 8, True)
    dd = dd.reshape(50, 3, 64)
    eps = 1e-3
    for d in range(3):
        xp, xm = x.copy(), x.copy()
        xp[:, d] += eps; xm[:, d] -= eps
        fd = (O.sh_encode_forward(xp, 8)[0].astype(np.float64) - O.sh_encode_forward(xm, 8)[0]) / (2 * eps)
        assert np.abs(fd - dd[:, d]).max() < 5e-2 * max(1.0, np.abs(dd[:, d]).max()) * 1e-1


def test_hash_index_python_rederivation():
    """get_grid_index (gridencoder.cu:62-79) re-derived in Python for one hashed and one dense level."""
    offsets, pls = O.grid_offsets(desired_resolution=2048)
    S = np.float32(np.log2(pls))
    table = np.zeros((int(offsets[-1]), 2), np.float32)
    rng = np.random.default_rng(0)
    for level in (2, 9):
        res = O.grid_resolution(level, S, 16)
        size = int(offsets[level + 1] - offsets[level])
        pg = rng.integers(0, res - 1, 3)
        # a point exactly on the node pg (align_corners=False: pos = x*res - 0.5)
        x = ((pg + 0.5) / res).astype(np.float32)[None]
        if res ** 3 <= size:
            idx = int(pg[0] + pg[1] * res + pg[2] * res * res) % size
        else:
            idx = int((np.uint32(pg[0]) * np.uint32(1)) ^ (np.uint32(pg[1]) * np.uint32(2654435761)) ^ (np.uint32(pg[2]) * np.uint32(805459861))) % size
        table[:] = 0
        table[offsets[level] + idx] = [1.0, 2.0]
        out, _ = O.grid_encode_forward(x, table, offsets, pls, 16)
        np.testing.assert_allclose(out[level, 0], [1.0, 2.0], atol=2e-3)
        assert np.abs(out[np.arange(16) != level]).max() == 0


def test_grid_oob_and_partial_levels():
    offsets, pls = O.grid_offsets(desired_resolution=2048)
    rng = np.random.default_rng(0)
    table = rng.random((int(offsets[-1]), 2), dtype=np.float32)
    x = np.array([[0.5, 0.5, 1.0001], [-1e-6, 0.2, 0.2], [0.3, 0.3, 0.3]], np.float32)
    out, dd = O.grid_encode_forward(x, table, offsets, pls, 16, True, max_level=8)
    assert np.all(out[:, 0] == 0) and np.all(out[:, 1] == 0) and np.all(out[8:] == 0) and np.all(out[:8, 2] != 0)
    assert np.all(dd[:2] == 0)


def test_composite_weights_sum_leq_one_and_grad_fd():
    bf = synth.occupancy_bitfield("blob", 128, 1, 1.0, seed=1)
    ro, rd, aabb, nears, fars, noises = scenes.make_rays(12, 12, 1.0, 20.0, seed=7)
    x, d, t, rays = O.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, noises)
    M = x.shape[0]
    rng = np.random.default_rng(0)
    sig = np.exp(rng.normal(0.0, 1.5, M)).astype(np.float32)
    rgb = rng.random((M, 3), dtype=np.float32)
    w, ws, dep, img = O.composite_rays_train_forward(sig, rgb, t, rays)
    assert ws.max() <= 1.0 + 1e-5 and w.min() >= 0
    # backward is the gradient of forward (finite differences in float64 on a few samples)
    gi = rng.normal(size=img.shape).astype(np.float32)
    gws = rng.normal(size=ws.shape).astype(np.float32)
    gs, gr = O.composite_rays_train_backward(np.zeros(M, np.float32), gws, np.zeros_like(dep), gi, sig, rgb, t, rays, ws, dep, img)
    def loss(s):
        _, ws_, _, im_ = O.composite_rays_train_forward(s, rgb, t, rays, T_thresh=0.0)
        return float((im_.astype(np.float64) * gi).sum() + (ws_.astype(np.float64) * gws).sum())
    gs0, _ = O.composite_rays_train_backward(np.zeros(M, np.float32), gws, np.zeros_like(dep), gi, sig, rgb, t, rays,
                                              *O.composite_rays_train_forward(sig, rgb, t, rays, T_thresh=0.0)[1:], T_thresh=0.0)
    for i in rng.integers(0, M, 6):
        e = max(1e-2 * sig[i], 1e-3)
        sp, sm = sig.copy(), sig.copy(); sp[i] += e; sm[i] -= e
        fd = (loss(sp) - loss(sm)) / (2 * e)
        assert abs(fd - gs0[i]) < 5e-2 * max(1e-3, abs(gs0[i])) + 2e-3, (fd, gs0[i])


def test_near_far_against_numpy_slab():
    ro, rd, aabb, nears, fars, _ = scenes.make_rays(16, 16, 1.0, 30.0, seed=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (aabb[:3] - ro) / rd; t1 = (aabb[3:] - ro) / rd
    tn = np.minimum(t0, t1).max(1); tf = np.maximum(t0, t1).min(1)
    hit = tn <= tf
    assert np.array_equal(hit, nears < 1e30)
    np.testing.assert_allclose(nears[hit], np.maximum(tn[hit], 0.2), rtol=1e-6)
    np.testing.assert_allclose(fars[hit], tf[hit], rtol=1e-6)


def test_freq_layout_and_backward():
    rng = np.random.default_rng(0)
    x = rng.random((7, 3), dtype=np.float32) * 2 - 1
    y = O.freq_encode_forward(x, 6)
    assert y.shape == (7, 39)
    np.testing.assert_array_equal(y[:, :3], x)
    np.testing.assert_allclose(y[:, 3:6], np.sin(x), atol=1e-6)
    np.testing.assert_allclose(y[:, 6:9], np.cos(x), atol=1e-6)
    np.testing.assert_allclose(y[:, 33:36], np.sin(32 * x), atol=1e-5)
    g = rng.normal(size=y.shape).astype(np.float32)
    gi = O.freq_encode_backward(g, y, 3, 6)
    ref = g[:, :3].astype(np.float64).copy()
    for f in range(6):
        ref += 2 ** f * (g[:, 3 + 6 * f:6 + 6 * f] * np.cos(2 ** f * x) - g[:, 6 + 6 * f:9 + 6 * f] * np.sin(2 ** f * x))
    np.testing.assert_allclose(gi, ref, rtol=1e-4, atol=1e-4)
