"""CPU: the warp-per-ray marching scheme of csrc/raymarch.cu (32 chain points per iteration + ballot
resolution of the visited set), emulated lane by lane in tests/helpers/warp_march_emu.c, produces
bit-identical counts and sample lists to the sequential restatement of the reference (oracle)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import scenes
from oracle import oracle as O
from sdf_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "helpers", "warp_march_emu.c")
    so = os.path.join(HERE, "helpers", "libwarp_march_emu.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
    lib = C.CDLL(so)
    lib.emu_march_ray.restype = C.c_uint32
    return lib


@pytest.mark.parametrize("case", scenes.MARCH_CASES, ids=[f"{c[0]}-b{c[1]}-g{c[3]:.4f}-c{int(c[5])}" for c in scenes.MARCH_CASES])
def test_warp_scheme_equals_sequential(emu, case):
    kind, bound, cas, dtg, max_steps, contract, fovy = case
    bf = synth.occupancy_bitfield(kind, 128, cas, bound, seed=1)
    ro, rd, aabb, nears, fars, noises = scenes.make_rays(20, 20, bound, fovy, seed=3)
    xyzs, dirs, ts, rays = O.march_rays_train(ro, rd, bound, bf, cas, 128, nears, fars, noises, dtg, max_steps, contract)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    x = np.zeros((max_steps, 3), np.float32); d = np.zeros((max_steps, 3), np.float32); t = np.zeros((max_steps, 2), np.float32)
    for n in range(ro.shape[0]):
        c = emu.emu_march_ray(p(ro[n]), p(rd[n]), p(bf), C.c_float(bound), C.c_int(int(contract)), C.c_float(dtg), C.c_uint32(max_steps),
                              C.c_uint32(cas), C.c_uint32(128), C.c_float(nears[n]), C.c_float(fars[n]), C.c_float(noises[n]),
                              C.c_uint32(max_steps), p(x), p(d), p(t))
        o, cnt = rays[n]
        assert c == cnt
        assert np.array_equal(x[:c], xyzs[o:o + c]) and np.array_equal(t[:c], ts[o:o + c]) and np.array_equal(d[:c], dirs[o:o + c])
    if kind == "empty":
        assert rays[:, 1].sum() == 0
    if kind == "full":
        assert rays[:, 1].max() == max_steps or rays[:, 1].max() > 100


def test_max_emit_truncation(emu):
    """the inference marcher stops after n_step samples (raymarching.cu:760)"""
    bf = synth.occupancy_bitfield("full", 128, 1, 1.0)
    ro, rd, aabb, nears, fars, noises = scenes.make_rays(8, 8, 1.0, 20.0, seed=1)
    alive = np.arange(ro.shape[0], dtype=np.int32)
    xo, do, to = O.march_rays(ro.shape[0], 5, alive, nears.copy(), ro, rd, 1.0, bf, 1, 128, nears, fars)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for n in range(ro.shape[0]):
        x = np.zeros((5, 3), np.float32); d = np.zeros((5, 3), np.float32); t = np.zeros((5, 2), np.float32)
        c = emu.emu_march_ray(p(ro[n]), p(rd[n]), p(bf), C.c_float(1.0), C.c_int(0), C.c_float(0.0), C.c_uint32(1024), C.c_uint32(1),
                              C.c_uint32(128), C.c_float(nears[n]), C.c_float(fars[n]), C.c_float(0.0), C.c_uint32(5), p(x), p(d), p(t))
        assert np.array_equal(x, xo[n * 5:(n + 1) * 5]) and np.array_equal(t, to[n * 5:(n + 1) * 5])
        assert c == int((to[n * 5:(n + 1) * 5, 0] != 0).sum())
