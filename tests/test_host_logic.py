"""CPU: host-side decisions that never touch the GPU — GEMM tile / split-K selection, the GEGLU row interleave, the bench's host
core detection, the reference-facing option defaults."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]


def test_gemm_config_choice_is_valid_and_sane():
    from sdf_b200.gemm import choose_config, estimate_us
    shapes = [(8192, 320, 45), (8192, 2560, 5), (2048, 640, 270), (512, 1280, 180), (128, 1280, 360), (262144, 128, 18), (65536, 256, 36),
              (16384, 512, 72), (2, 1280, 5), (154, 640, 12), (4096, 4, 72), (4096, 8, 72)]
    for M, N, kb in shapes:
        bn, pair, sk = choose_config(M, N, kb)
        assert (bn, pair) in ((64, 0), (128, 0), (160, 0), (128, 1), (160, 1), (256, 1))
        assert sk >= 1 and kb // sk >= 1
        if pair:
            assert M > 128 and N % 256 == 0                      # the pair tile is only chosen where it was measured to win
        if sk > 1:
            assert kb // sk >= 8                                 # never split below 8 k-blocks per CTA
        assert estimate_us(M, N, kb, bn, pair, sk) <= estimate_us(M, N, kb, 128 if N > 64 else 64, 0, 1) + 1e-6
    # batched products (attention) never split K and never pair
    bn, pair, sk = choose_config(4096, 4096, 8, batched=True)
    assert pair == 0 and sk == 1
    # the wide VAE convolutions get the 2-CTA tile, the single-wave UNet convolutions do not split
    assert choose_config(65536, 256, 36)[:2] == (256, 1)
    assert choose_config(8192, 320, 45) == (160, 0, 1)


def test_geglu_row_permutation_interleaves_value_and_gate():
    from sdf_b200.sd_engine import geglu_row_permutation
    inner = 64
    perm = geglu_row_permutation(inner, torch.device("cpu"))
    assert sorted(perm.tolist()) == list(range(2 * inner))
    for c in range(2 * inner // 32):
        chunk = perm[32 * c:32 * c + 32]
        assert chunk[:16].tolist() == list(range(16 * c, 16 * c + 16))                      # 16 value rows ...
        assert chunk[16:].tolist() == list(range(inner + 16 * c, inner + 16 * c + 16))      # ... then their 16 gates
    # the fused epilogue's arithmetic on the permuted projection equals GEGLU on the original one
    w, b, x = torch.randn(2 * inner, 24), torch.randn(2 * inner), torch.randn(5, 24)
    y = x @ w.t() + b
    ref = y[:, :inner] * torch.nn.functional.gelu(y[:, inner:])
    yp = (x @ w[perm].t() + b[perm]).view(5, -1, 32)
    fused = (yp[..., :16] * torch.nn.functional.gelu(yp[..., 16:])).reshape(5, inner)
    assert torch.allclose(fused, ref, rtol=1e-5, atol=1e-5)


def test_host_core_detection_is_bounded():
    import bench
    n = bench.host_core_limit()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_default_options_match_the_reference_preset():
    from sdf_b200.options import default_opt
    o = default_opt(h=64, w=64)
    # main.py:58-132 defaults + the -O preset (main.py:155-160)
    assert (o.iters, o.lr, o.bound, o.dt_gamma, o.max_steps, o.update_extra_interval) == (10000, 1e-3, 1, 0, 1024, 16)
    assert (o.latent_iter_ratio, o.albedo_iter_ratio, o.min_ambient_ratio, o.textureless_ratio) == (0.2, 0, 0.1, 0.2)
    assert (o.lambda_entropy, o.lambda_opacity, o.lambda_orient) == (1e-3, 0, 1e-2)
    assert (o.blob_density, o.blob_radius, o.bg_radius, o.density_activation) == (5, 0.2, 1.4, 'exp')
    assert tuple(o.radius_range) == (3.0, 3.5) and tuple(o.fovy_range) == (10, 30) and tuple(o.theta_range) == (45, 105)


def test_x_neighbour_corner_pairs_share_an_aligned_entry_pair():
    """The fused field kernels fetch / reduce the two x-neighbours of a corner pair with ONE 8-byte load / 16-byte reduction when
    (i0 ^ i1) == 1 (csrc/field_common.cuh: pair_issue, fused_field_bwd.cu: scatter_level).  Restate the level index rule of the
    reference (gridencoder/src/gridencoder.cu:46-79) and check the algebra the kernels rely on:
      * (i0 ^ i1) == 1  =>  the two entries are {2m, 2m+1}: adjacent and 8-byte aligned (level offsets are multiples of 8 entries);
      * hashed levels with even x0 ALWAYS satisfy it (x1 = x0 ^ 1 flips only bit 0 of the hash, the level size is a power of two);
      * dense levels satisfy it exactly when the linear index of the x0 corner is even."""
    import numpy as np
    from oracle import oracle as O
    offsets, pls = O.grid_offsets(desired_resolution=2048)
    assert all(int(o) % 8 == 0 for o in offsets)
    rng = np.random.default_rng(0)
    S = np.log2(pls)
    for level in range(16):
        size = int(offsets[level + 1] - offsets[level])
        res = int(np.ceil(np.exp2(level * S) * 16))          # gridencoder.cu:109 (align_corners = False)
        hashed = res ** 3 > size
        p = rng.integers(0, res - 1, size=(20000, 3)).astype(np.uint64)
        x0, y, z = p[:, 0], p[:, 1], p[:, 2]
        x1 = np.minimum(x0 + 1, res - 1)
        if hashed:
            assert size & (size - 1) == 0                    # every hashed level of the -O backbone is a power of two
            h = lambda x: ((x ^ (y * np.uint64(2654435761)) ^ (z * np.uint64(805459861))) & np.uint64(0xFFFFFFFF)) % np.uint64(size)
        else:
            h = lambda x: (x + y * np.uint64(res) + z * np.uint64(res * res)) % np.uint64(size)
        i0, i1 = h(x0), h(x1)
        merged = (i0 ^ i1) == 1
        assert np.all(np.minimum(i0, i1)[merged] % 2 == 0) and np.all(np.abs(i0.astype(np.int64) - i1.astype(np.int64))[merged] == 1)
        if hashed:
            assert np.all(merged[x0 % 2 == 0])               # even x0: always one load
        else:
            assert np.array_equal(merged, i0 % 2 == 0)       # dense: exactly the even linear indices
        assert 0.35 < merged.mean() < 0.65                   # ~half of all corner pairs merge -> ~25 % fewer scattered lanes


def test_default_options_equal_the_reference_parser_output():
    """every option the host mirror uses, against what main.py's own argparse block produces for `--text x -O`
    (tests/golden/options_O.json, written by tests/golden/make_golden_opts.py from the reference source)"""
    import json
    from sdf_b200.options import default_opt
    path = os.path.join(ROOT, "tests", "golden", "options_O.json")
    ref = json.load(open(path))
    ours = vars(default_opt(h=ref["h"], w=ref["w"]))
    shared = [k for k in ours if k in ref]
    assert len(shared) >= 30
    for k in shared:
        a, b = ours[k], ref[k]
        if isinstance(a, (list, tuple)):
            assert list(a) == list(b), k
        else:
            assert a == b, (k, a, b)


def test_dmtet_options_follow_main_py_overrides():
    """main.py:253-274 for a `--dmtet` run: render size x dmtet_reso_scale, Magic3D's fine-stage t_range, no latent / albedo warm-up;
    the mesh-regulariser weights and lattice size are the parser defaults (main.py:47,134-135) recorded in tests/golden/options_O.json"""
    import json
    from sdf_b200.options import dmtet_opt
    o = dmtet_opt()
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "options_O.json")))
    ref = ref.get("opt", ref)
    assert (o.h, o.w) == (int(64 * ref["dmtet_reso_scale"]), int(64 * ref["dmtet_reso_scale"])) == (512, 512)
    assert o.t_range == [0.02, 0.50] and o.latent_iter_ratio == 0 and o.albedo_iter_ratio == 0 and o.dmtet
    for k in ("tet_grid_size", "lambda_mesh_normal", "lambda_mesh_laplacian", "lock_geo", "dmtet_reso_scale"):
        assert getattr(o, k) == ref[k], k


def test_nvdiffrast_dropin_exposes_the_surface_run_dmtet_binds():
    """nerf/renderer.py:12,309-312,895-931 use exactly these names of `nvdiffrast.torch` (no GPU work here: import + signatures)"""
    import inspect
    import nvdiffrast.torch as dr
    assert dr.__file__.startswith(os.path.join(ROOT, "stable-dreamfusion_b200"))
    for name in ("RasterizeCudaContext", "RasterizeGLContext", "rasterize", "interpolate", "antialias"):
        assert hasattr(dr, name), name
    assert list(inspect.signature(dr.rasterize).parameters)[:4] == ["glctx", "pos", "tri", "resolution"]
    assert list(inspect.signature(dr.interpolate).parameters)[:3] == ["attr", "rast", "tri"]
    assert list(inspect.signature(dr.antialias).parameters)[:4] == ["color", "rast", "pos", "tri"]
    dr.RasterizeCudaContext()
    dr.RasterizeGLContext()


def test_bench_shading_cycle_is_the_reference_mix_in_every_window():
    """bench.py's 25-step shading cycle: 20 % latent / 64 % lambertian / 16 % textureless (main.py:150-153 with the -O preset), interleaved so that
    a timed window of any length — the driver chooses --steps — holds the mix to within one step and the resident and end-to-end windows match"""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    c = bench.CYCLE
    assert len(c) == 25 and c.count("latent") == 5 and c.count("lambertian") == 16 and c.count("textureless") == 4
    ring = c + c + c
    for start in range(25):
        for K in (5, 10, 20, 25, 30, 40, 50):
            w = ring[start:start + K]
            assert abs(w.count("latent") - 0.2 * K) <= 1.0, (start, K)
            assert abs(w.count("textureless") - 0.16 * K) <= 1.5, (start, K)
