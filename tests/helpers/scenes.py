"""Seeded test scenes shared by the CPU and GPU tests."""
import numpy as np

from sdf_b200 import synth
from oracle import oracle as O


def make_rays(H=32, W=32, bound=1.0, fovy=20.0, seed=0, default_view=False):
    rng = np.random.default_rng(seed)
    if default_view:
        pose = synth.circle_pose(3.2, 90.0, 0.0)
    else:
        pose, _ = synth.rand_pose(rng)
    ro, rd = synth.get_rays(pose, H, W, fovy)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(ro, rd, aabb, 0.2)
    noises = rng.random(ro.shape[0], dtype=np.float32)
    return ro, rd, aabb, nears, fars, noises


MARCH_CASES = [
    # kind, bound, cascade, dt_gamma, max_steps, contract, fovy
    ("blob", 1.0, 1, 0.0, 1024, False, 20.0),
    ("sparse", 1.0, 1, 0.0, 1024, False, 20.0),
    ("full", 1.0, 1, 0.0, 256, False, 25.0),
    ("empty", 1.0, 1, 0.0, 1024, False, 20.0),
    ("blob", 2.0, 2, 1.0 / 128, 1024, False, 45.0),
    ("sparse", 4.0, 3, 1.0 / 256, 512, True, 60.0),
    ("sparse", 2.0, 2, 0.0, 1024, True, 50.0),
]
