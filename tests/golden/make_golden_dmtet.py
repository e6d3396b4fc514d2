"""Pins oracle/dmtet_ref.py (and through it csrc/dmtet.cu) to the reference's own DMTet mesh-extraction code.

Run HERE (the container that has /root/reference; CPU only):  python tests/golden/make_golden_dmtet.py
It imports nerf/renderer.py from the reference with stub modules for the packages this code never touches, runs the reference's
`DMTet.__call__` (nerf/renderer.py:94-174), `normal_consistency` (:209-222) and `laplacian_smooth_loss` (:248-254) on this repository's
tetrahedral lattice with seeded signed distances / deformations, asserts that the numpy restatement agrees, and writes
tests/golden/dmtet.npz (inputs + the REFERENCE's outputs) for tests/test_oracle_dmtet_golden.py and the GPU parity tests.
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
    for name in ["gridencoder", "freqencoder", "shencoder", "torch.utils.cpp_extension", "mcubes", "trimesh", "nvdiffrast", "nvdiffrast.torch",
                 "raymarching", "cubvh", "xatlas", "pymeshlab", "tensorboardX", "imageio", "torch_ema", "torchmetrics", "matplotlib",
                 "matplotlib.pyplot", "lpips", "rich", "rich.console"]:
        sys.modules.setdefault(name, _Stub(name))
    import nerf.renderer as R
    return R


def scene(n, seed):
    sys.path.insert(0, os.path.join(ROOT, "stable-dreamfusion_b200"))
    from sdf_b200 import tetgrid
    verts, tets = tetgrid.make_tet_grid(n)
    rng = np.random.default_rng(seed)
    sdf = (0.55 - np.linalg.norm(verts * np.array([1.0, 1.3, 0.8], np.float32), axis=-1) + 0.08 * rng.standard_normal(len(verts))).astype(np.float32)
    deform = (np.tanh(rng.standard_normal(verts.shape)) * (0.5 / n)).astype(np.float32)
    return verts, tets, sdf, deform


def main():
    R = load_reference()
    torch.Tensor.cuda = lambda self, *a, **k: self                 # the loss helpers call .cuda() on index tensors (nerf/renderer.py:196-198)
    sys.path.insert(0, ROOT)
    from oracle import dmtet_ref as O
    out = {}
    for tag, n, seed in (("a", 5, 1), ("b", 11, 2)):
        verts, tets, sdf, deform = scene(n, seed)
        pos = (verts + deform).astype(np.float32)
        model = R.DMTet("cpu")
        v_ref, f_ref = model(torch.from_numpy(pos), torch.from_numpy(sdf.copy()), torch.from_numpy(tets))
        v_ref, f_ref = v_ref.numpy(), f_ref.numpy()
        v_o, f_o = O.marching_tets(pos, sdf, tets)
        assert v_o.shape == v_ref.shape and f_o.shape == f_ref.shape, (v_o.shape, v_ref.shape, f_o.shape, f_ref.shape)
        assert np.array_equal(f_o, f_ref), "face indices differ from the reference's"
        assert np.abs(v_o - v_ref).max() <= 1e-6, np.abs(v_o - v_ref).max()
        fn, vn = O.mesh_normals(v_ref, f_ref)
        nc_ref = float(R.normal_consistency(torch.from_numpy(fn), torch.from_numpy(f_ref)))
        lap_ref = float(R.laplacian_smooth_loss(torch.from_numpy(v_ref), torch.from_numpy(f_ref)))
        assert abs(O.normal_consistency(fn, f_ref) - nc_ref) < 1e-5, (O.normal_consistency(fn, f_ref), nc_ref)
        assert abs(O.laplacian_smooth_loss(v_ref, f_ref) - lap_ref) < 1e-5 * max(1.0, lap_ref), (O.laplacian_smooth_loss(v_ref, f_ref), lap_ref)
        print(f"scene {tag}: grid {n}^3 cells, {len(verts)} lattice vertices, {len(tets)} tetrahedra -> {len(v_ref)} vertices, {len(f_ref)} faces; "
              f"normal_consistency {nc_ref:.6f}, laplacian {lap_ref:.6f}")
        out.update({f"{tag}_n": np.int64(n), f"{tag}_sdf": sdf, f"{tag}_deform": deform, f"{tag}_verts": v_ref, f"{tag}_faces": f_ref.astype(np.int32),
                    f"{tag}_normal_consistency": np.float64(nc_ref), f"{tag}_laplacian": np.float64(lap_ref)})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dmtet.npz"), **out)
    print("wrote tests/golden/dmtet.npz")


if __name__ == "__main__":
    main()
