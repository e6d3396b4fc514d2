"""Training options of the hot path with the defaults of the reference's main.py (:19-170) and its
-O preset (:172-174: fp16 + cuda_ray).  Only the fields the SDS inner loop reads are present."""
from types import SimpleNamespace


def default_opt(**over):
    o = SimpleNamespace(
        # scene / marching (main.py:102-104, 63, 66)
        bound=1.0, dt_gamma=0.0, min_near=0.01, max_steps=1024, update_extra_interval=16,
        # field (main.py:80-84)
        bg_radius=1.4, density_activation='exp', density_thresh=10.0, blob_density=5.0, blob_radius=0.2,
        # render size / batch (main.py:94-99)
        w=64, h=64, batch_size=1,
        # cameras (main.py:106-114)
        radius_range=[3.0, 3.5], theta_range=[45, 105], phi_range=[-180, 180], fovy_range=[10, 30], default_fovy=20,
        # schedule (main.py:58, 68-71)
        iters=10000, latent_iter_ratio=0.2, albedo_iter_ratio=0.0, min_ambient_ratio=0.1, textureless_ratio=0.2,
        # losses (main.py:128-143)
        lambda_entropy=1e-3, lambda_opacity=0.0, lambda_orient=1e-2, lambda_tv=0.0, lambda_wd=0.0, lambda_guidance=1.0,
        lambda_normal=0.0, lambda_2d_normal_smooth=0.0, lambda_3d_normal_smooth=0.0,
        # guidance (main.py:40, 123)
        guidance_scale=100.0, t_range=[0.02, 0.98],
        # optimiser (main.py:59, 368)
        lr=1e-3,
        # presets
        fp16=True, cuda_ray=True, dmtet=False, taichi_ray=False, progressive_level=False,
        # DMTet stage (main.py:46-50, 98, 137-138)
        tet_grid_size=128, lock_geo=False, dmtet_reso_scale=8, lambda_mesh_normal=0.5, lambda_mesh_laplacian=0.5,
    )
    for k, v in over.items():
        setattr(o, k, v)
    return o


def dmtet_opt(**over):
    """options of a `--dmtet` run after main.py:253-274: render size x dmtet_reso_scale, no latent / albedo warm-up, t_range of Magic3D's
    fine stage"""
    o = default_opt(dmtet=True, latent_iter_ratio=0.0, albedo_iter_ratio=0.0, t_range=[0.02, 0.50], progressive_view=False)
    o.h, o.w = int(o.h * o.dmtet_reso_scale), int(o.w * o.dmtet_reso_scale)
    for k, v in over.items():
        setattr(o, k, v)
    return o
