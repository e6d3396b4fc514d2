"""Pins oracle/nerf_o2.py (the port behind bench.py's CPU reference arm) to the reference's own `-O2` modules.

Run HERE (the container that has /root/reference; CPU only):  python tests/golden/make_golden_o2.py
It imports nerf/network.py + nerf/renderer.py from the reference (stub modules for the packages the -O2 path never touches),
copies the reference network's weights into the port, runs both `run()` / `render()` with identical seeds and inputs, asserts they
agree, and writes tests/golden/nerf_o2.npz (weights + inputs + the REFERENCE's outputs) for tests/test_oracle_o2_golden.py.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        return self


def load_reference():
    sys.path.insert(0, REF)
    # the reference's extension packages JIT-compile at import: never let them load here (the -O2 path does not use them)
    for name in ["gridencoder", "freqencoder", "shencoder", "torch.utils.cpp_extension"]:
        sys.modules.setdefault(name, _Stub(name))
    for name in ["mcubes", "trimesh", "nvdiffrast", "nvdiffrast.torch", "raymarching", "cubvh", "xatlas", "pymeshlab", "tensorboardX", "imageio",
                 "torch_ema", "torchmetrics", "matplotlib", "matplotlib.pyplot", "lpips", "rich", "rich.console"]:
        sys.modules.setdefault(name, _Stub(name))
    import nerf.network as N
    return N


def main():
    N = load_reference()
    sys.path.insert(0, ROOT)
    from oracle import nerf_o2
    opt = types.SimpleNamespace(bound=1, dmtet=False, cuda_ray=False, taichi_ray=False, min_near=0.01, density_thresh=10, h=32, w=32, tet_grid_size=128,
                                density_activation="exp", blob_density=5, blob_radius=0.2, bg_radius=1.4, num_steps=64, upsample_steps=32,
                                lambda_orient=1e-2, lambda_3d_normal_smooth=0, lambda_2d_normal_smooth=0, lambda_normal=0, lock_geo=False,
                                max_steps=1024, dt_gamma=0, lambda_mesh_normal=0, lambda_mesh_laplacian=0, backbone="vanilla", optim="adan", fp16=False)
    torch.manual_seed(0)
    ref = N.NeRFNetwork(opt)
    ref.train()
    port = nerf_o2.VanillaNeRF()
    # same parameters: the port keeps the reference's module / parameter names for sigma_net and bg_net
    sd = {k: v for k, v in ref.state_dict().items() if k.startswith(("sigma_net.", "bg_net."))}
    missing, unexpected = port.load_state_dict(sd, strict=False)
    assert not unexpected and all(not m.startswith(("sigma_net.", "bg_net.")) for m in missing), (missing, unexpected)
    sys.path.insert(0, os.path.join(ROOT, "stable-dreamfusion_b200"))
    from sdf_b200 import synth
    pose = synth.circle_pose(3.2, 70.0, 30.0)
    ro, rd = synth.get_rays(pose, 16, 16, 20.0)
    t_ro, t_rd = torch.from_numpy(ro)[None], torch.from_numpy(rd)[None]
    out = {"rays_o": ro, "rays_d": rd}
    for k, v in sd.items():
        out["w." + k] = v.detach().numpy()
    for shading in ("albedo", "lambertian", "textureless", "normal"):
        torch.manual_seed(123)
        r = ref.run(t_ro, t_rd, light_d=None, ambient_ratio=0.4, shading=shading, bg_color=None, perturb=True)
        torch.manual_seed(123)
        p = port.render(t_ro, t_rd, 0.4, shading, None, perturb=True)
        img_r = r["image"].detach().reshape(-1, 3)
        ws_r = r["weights_sum"].detach().reshape(-1)
        err_i = (img_r - p["image"].detach()).abs().max().item()
        err_w = (ws_r - p["weights_sum"].detach()).abs().max().item()
        print(f"{shading:12s} max |image diff| {err_i:.3e}  max |weights_sum diff| {err_w:.3e}", end="")
        assert err_i < 1e-5 and err_w < 1e-5, "the port no longer matches the reference"
        out[f"{shading}.image"] = img_r.numpy()
        out[f"{shading}.weights_sum"] = ws_r.numpy()
        if "loss_orient" in r:
            lo_r, lo_p = float(r["loss_orient"]), float(p["loss_orient"])
            print(f"  loss_orient ref {lo_r:.6e} port {lo_p:.6e}", end="")
            assert abs(lo_r - lo_p) <= 1e-6 + 1e-4 * abs(lo_r)
            out[f"{shading}.loss_orient"] = np.float32(lo_r)
        # gradient wrt the first layer's weight through the whole render (autograd normals included)
        g_r = torch.autograd.grad(r["image"].sum() + r["weights_sum"].sum(), ref.sigma_net.net[0].dense.weight, retain_graph=False)[0]
        g_p = torch.autograd.grad(p["image"].sum() + p["weights_sum"].sum(), port.sigma_net.net[0].dense.weight)[0]
        err_g = ((g_r - g_p).abs().max() / (g_r.abs().max() + 1e-12)).item()
        print(f"  rel grad diff {err_g:.3e}")
        assert err_g < 1e-4
        out[f"{shading}.grad_w0"] = g_r.numpy()
    # camera rays: nerf/utils.py:113-176 get_rays (N = -1, pixel centres) for two poses / intrinsics
    from nerf.utils import get_rays as ref_get_rays
    rng = np.random.default_rng(5)
    for i, (hh, ww, fov) in enumerate([(16, 16, 20.0), (8, 24, 37.5)]):
        pose_i, _ = synth.rand_pose(rng)
        focal = hh / (2 * np.tan(np.deg2rad(fov) / 2))
        rr = ref_get_rays(torch.from_numpy(pose_i)[None].float(), np.array([focal, focal, ww / 2, hh / 2]), hh, ww, -1)
        out[f"cam{i}.pose"] = pose_i.astype(np.float32)
        out[f"cam{i}.hwf"] = np.array([hh, ww, focal], np.float64)
        out[f"cam{i}.rays_o"] = rr["rays_o"][0].numpy()
        out[f"cam{i}.rays_d"] = rr["rays_d"][0].numpy()
    # orbit camera poses: nerf/provider.py:151-197 circle_poses (the deterministic core of rand_poses with the default
    # uniform_sphere_rate = 0, jitter_pose = False of main.py:72-76)
    import nerf.provider as P
    orbit = np.array([(3.2, 90.0, 0.0), (3.0, 60.0, 45.0), (3.5, 105.0, -170.0), (3.2, 45.0, 200.0), (3.33, 75.0, 90.0)], np.float32)
    ref_poses = []
    for r_, th_, ph_ in orbit:
        res = P.circle_poses("cpu", radius=torch.tensor([r_]), theta=torch.tensor([th_]), phi=torch.tensor([ph_]))
        ref_poses.append((res[0] if isinstance(res, tuple) else res)[0].numpy())
    out["orbit.params"] = orbit
    out["orbit.poses"] = np.stack(ref_poses)
    path = os.path.join(ROOT, "tests", "golden", "nerf_o2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
