"""CPU: oracle/nerf_o2.py (the port timed by bench.py's CPU reference arm) replayed against tests/golden/nerf_o2.npz — outputs of the
reference's own `-O2` modules (nerf/network.py, nerf/renderer.py:run) on the same weights, rays and seeds, written by
tests/golden/make_golden_o2.py in the container that has /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import nerf_o2

PATH = os.path.join(os.path.dirname(__file__), "golden", "nerf_o2.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/nerf_o2.npz not generated")


@pytest.mark.parametrize("shading", ["albedo", "lambertian", "textureless", "normal"])
def test_o2_port_matches_reference_outputs(shading):
    g = np.load(PATH)
    port = nerf_o2.VanillaNeRF()
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}
    missing, unexpected = port.load_state_dict(sd, strict=False)
    assert not unexpected
    ro, rd = torch.from_numpy(g["rays_o"])[None], torch.from_numpy(g["rays_d"])[None]
    torch.manual_seed(123)                       # the generator seeds the reference the same way before each run()
    out = port.render(ro, rd, 0.4, shading, None, perturb=True)
    # same PyTorch CPU kernels, same operation order: agreement to float rounding of the library version
    np.testing.assert_allclose(out["image"].detach().numpy(), g[f"{shading}.image"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["weights_sum"].detach().numpy(), g[f"{shading}.weights_sum"], rtol=1e-5, atol=1e-6)
    if f"{shading}.loss_orient" in g.files:
        assert abs(float(out["loss_orient"]) - float(g[f"{shading}.loss_orient"])) <= 1e-6 + 1e-4 * abs(float(g[f"{shading}.loss_orient"]))
    gw = torch.autograd.grad(out["image"].sum() + out["weights_sum"].sum(), port.sigma_net.net[0].dense.weight)[0]
    ref = g[f"{shading}.grad_w0"]
    assert np.abs(gw.numpy() - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("cam", [0, 1])
def test_camera_rays_match_reference_get_rays(cam):
    """sdf_b200.trainer.get_rays_torch and sdf_b200.synth.get_rays against nerf/utils.py:113-176 (pixel centres, unnormalised
    directions, camera-to-world rotation applied as d @ R^T) — square and non-square images"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stable-dreamfusion_b200"))
    from sdf_b200.trainer import get_rays_torch
    g = np.load(PATH)
    if f"cam{cam}.pose" not in g.files:
        pytest.skip("fixture predates the camera vectors")
    H, W, focal = g[f"cam{cam}.hwf"]
    H, W = int(H), int(W)
    ro, rd = get_rays_torch(torch.from_numpy(g[f"cam{cam}.pose"])[None], float(focal), W / 2, H / 2, H, W)
    assert np.array_equal(ro[0].numpy(), g[f"cam{cam}.rays_o"])
    np.testing.assert_allclose(rd[0].numpy(), g[f"cam{cam}.rays_d"], rtol=0, atol=2e-7)      # three FMAs instead of a 3x3 matmul


def test_orbit_poses_match_reference_circle_poses():
    """sdf_b200.synth.circle_pose (look-at orbit camera of the benchmark / trainer) against nerf/provider.py:151-197"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stable-dreamfusion_b200"))
    from sdf_b200 import synth
    g = np.load(PATH)
    if "orbit.params" not in g.files:
        pytest.skip("fixture predates the orbit poses")
    for (r, th, ph), ref in zip(g["orbit.params"], g["orbit.poses"]):
        np.testing.assert_allclose(synth.circle_pose(float(r), float(th), float(ph)), ref, rtol=0, atol=2e-6)
