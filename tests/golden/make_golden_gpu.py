"""Generates tests/golden/ref_cuda.npz: outputs of the REFERENCE'S OWN CUDA extensions (oracle/_ref/_*.so, built from the unmodified
sources under /root/reference by oracle/build_ref.py) on seeded inputs, executed on a B200.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden_gpu.py'        (writes gpurun_out/ref_cuda.npz; copy it to tests/golden/)

tests/test_oracle_golden.py then replays the inputs through the CPU oracle (oracle/sdf_oracle.c) anywhere — this is what pins
the oracle restatement to the reference kernels without a GPU or the reference tree."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200"), os.path.join(ROOT, "tests")]
from helpers import ref, scenes          # noqa: E402
from oracle import oracle as O           # noqa: E402
from sdf_b200 import synth               # noqa: E402

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
out = {}

rm = ref.load("_raymarching")
assert rm is not None, "oracle/_ref/_raymarching.so missing"
# ---- marching + compositing, two scenes
for tag, (kind, bound, cas, dtg, ms, contract, fovy) in {"a": ("blob", 1.0, 1, 0.0, 1024, False, 20.0), "b": ("sparse", 2.0, 2, 1.0 / 128, 512, True, 50.0)}.items():
    bf = synth.occupancy_bitfield(kind, 128, cas, bound, seed=1)
    ro, rd, aabb, _, _, noises = scenes.make_rays(16, 16, bound, fovy, seed=11)
    N = ro.shape[0]
    nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
    rm.near_far_from_aabb(T(ro), T(rd), T(aabb), N, 0.2, nears, fars)
    rays = torch.empty(N, 2, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    rm.march_rays_train(T(ro), T(rd), T(bf), bound, contract, dtg, ms, N, cas, 128, nears, fars, None, None, None, rays, cnt, T(noises))
    M = int(cnt.item())
    x = torch.zeros(M, 3, device=dev); d = torch.zeros(M, 3, device=dev); ts = torch.zeros(M, 2, device=dev)
    rm.march_rays_train(T(ro), T(rd), T(bf), bound, contract, dtg, ms, N, cas, 128, nears, fars, x, d, ts, rays, cnt, T(noises))
    torch.cuda.synchronize()
    # canonical (ray-order) packing
    r = rays.cpu().numpy(); c = r[:, 1].astype(np.int64)
    idx = np.concatenate([np.arange(o, o + k) for o, k in zip(r[:, 0], c)]) if M else np.zeros(0, np.int64)
    offs = np.concatenate([[0], np.cumsum(c)[:-1]]).astype(np.int32)
    rays_c = np.stack([offs, c.astype(np.int32)], 1)
    xs, tss = x.cpu().numpy()[idx], ts.cpu().numpy()[idx]
    rng = np.random.default_rng(5)
    sig = np.exp(rng.normal(1.5, 2.0, M)).astype(np.float32); rgb = rng.random((M, 3), dtype=np.float32)
    w = torch.zeros(M, device=dev); ws = torch.empty(N, device=dev); dep = torch.empty(N, device=dev); img = torch.empty(N, 3, device=dev)
    rm.composite_rays_train_forward(T(sig), T(rgb), T(tss), T(rays_c), M, N, 1e-4, False, w, ws, dep, img)
    gw = (rng.normal(size=M) * 0.1).astype(np.float32); gws = rng.normal(size=N).astype(np.float32)
    gd = (rng.normal(size=N) * 0.1).astype(np.float32); gi = rng.normal(size=(N, 3)).astype(np.float32)
    gs = torch.zeros(M, device=dev); gr = torch.zeros(M, 3, device=dev)
    rm.composite_rays_train_backward(T(gw), T(gws), T(gd), T(gi), T(sig), T(rgb), T(tss), T(rays_c), ws, dep, img, M, N, 1e-4, False, gs, gr)
    torch.cuda.synchronize()
    out.update({f"march_{tag}/cfg": np.array([bound, cas, dtg, ms, int(contract), fovy], np.float64), f"march_{tag}/bitfield": bf,
                f"march_{tag}/rays_o": ro, f"march_{tag}/rays_d": rd, f"march_{tag}/noises": noises,
                f"march_{tag}/nears": nears.cpu().numpy(), f"march_{tag}/fars": fars.cpu().numpy(), f"march_{tag}/rays": rays_c,
                f"march_{tag}/xyzs": xs, f"march_{tag}/ts": tss, f"march_{tag}/sig": sig, f"march_{tag}/rgb": rgb,
                f"march_{tag}/weights": w.cpu().numpy(), f"march_{tag}/weights_sum": ws.cpu().numpy(), f"march_{tag}/depth": dep.cpu().numpy(),
                f"march_{tag}/image": img.cpu().numpy(), f"march_{tag}/gw": gw, f"march_{tag}/gws": gws, f"march_{tag}/gd": gd, f"march_{tag}/gi": gi,
                f"march_{tag}/grad_sigmas": gs.cpu().numpy(), f"march_{tag}/grad_rgbs": gr.cpu().numpy()})
# ---- morton / packbits
rng = np.random.default_rng(0)
coords = rng.integers(0, 128, (4096, 3), dtype=np.int32)
ind = torch.empty(4096, dtype=torch.int32, device=dev); rm.morton3D(T(coords), 4096, ind)
grid = rng.random(8 * 1024, dtype=np.float32); bits = torch.empty(1024, dtype=torch.uint8, device=dev); rm.packbits(T(grid.reshape(1, -1)), 1024, 0.4, bits)
torch.cuda.synchronize()
out.update({"morton/coords": coords, "morton/indices": ind.cpu().numpy(), "packbits/grid": grid, "packbits/bits": bits.cpu().numpy()})

# ---- hash grid (small table), fp32 and fp16
ge = ref.load("_gridencoder")
offsets, pls = O.grid_offsets(3, 8, 2, 2.0, 16, 12, 512)
n = int(offsets[-1])
xin = rng.random((600, 3), dtype=np.float32); xin[:8] = rng.random((8, 3), dtype=np.float32) * 1.2 - 0.1
table = (rng.random((n, 2), dtype=np.float32) - 0.5)
S = float(np.log2(pls))
for half in (False, True):
    dt = torch.float16 if half else torch.float32
    o = torch.zeros(8, 600, 2, device=dev, dtype=dt); dd = torch.zeros(600, 8 * 3 * 2, device=dev, dtype=dt)
    ge.grid_encode_forward(T(xin), T(table).to(dt), T(offsets), o, 600, 3, 2, 8, 8, S, 16, dd, 0, False, 1)
    torch.cuda.synchronize()
    out[f"grid/out_{'f16' if half else 'f32'}"] = o.float().cpu().numpy()
    out[f"grid/dydx_{'f16' if half else 'f32'}"] = dd.float().cpu().numpy()
g = rng.normal(size=(8, 600, 2)).astype(np.float32)
gg = torch.zeros(n, 2, device=dev); gin = torch.zeros(600, 3, device=dev)
ge.grid_encode_backward(T(g), T(xin), T(table), T(offsets), gg, 600, 3, 2, 8, 8, S, 16, T(out["grid/dydx_f32"]), gin, 0, False, 1)
torch.cuda.synchronize()
out.update({"grid/offsets": offsets, "grid/pls": np.array([pls]), "grid/x": xin, "grid/table": table, "grid/grad": g,
            "grid/grad_table": gg.cpu().numpy(), "grid/grad_inputs": gin.cpu().numpy()})

# ---- freq / SH
fe = ref.load("_freqencoder"); she = ref.load("_shencoder")
xf = (rng.random((500, 3), dtype=np.float32) * 2 - 1)
yf = torch.empty(500, 39, device=dev); fe.freq_encode_forward(T(xf), 500, 3, 6, 39, yf)
gf = rng.normal(size=(500, 39)).astype(np.float32); gif = torch.zeros(500, 3, device=dev); fe.freq_encode_backward(T(gf), yf, 500, 3, 6, 39, gif)
xs_ = rng.normal(size=(300, 3)).astype(np.float32); xs_ /= np.linalg.norm(xs_, axis=1, keepdims=True)
for deg in (4, 8):
    ys = torch.empty(300, deg * deg, device=dev); dds = torch.empty(300, 3 * deg * deg, device=dev)
    she.sh_encode_forward(T(xs_), ys, 300, 3, deg, dds)
    torch.cuda.synchronize()
    out[f"sh/y{deg}"] = ys.cpu().numpy(); out[f"sh/dydx{deg}"] = dds.cpu().numpy()
torch.cuda.synchronize()
out.update({"freq/x": xf, "freq/y": yf.cpu().numpy(), "freq/g": gf, "freq/gi": gif.cpu().numpy(), "sh/x": xs_})

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "ref_cuda.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB;", torch.cuda.get_device_name(0))
