"""DMTet stage of the hot path (BASELINE config C5; reference: nerf/renderer.py:94-174, 291-310, 818-954; main.py:253-274).

The reference extracts a triangle mesh from a signed-distance lattice with eager PyTorch (two host-synchronising torch.unique calls per
step), rasterises it with nvdiffrast and textures it with the hash-grid network.  Here every piece between the parameters (sdf, deform,
hash table, MLP) and the rendered image is a capacity-sized kernel of csrc/dmtet.cu / csrc/meshrast.cu with device-side counts; the only
library call is one torch.sort of the half-edge keys for the mesh regularisers.

    lattice = TetLattice(tet_grid_size, device)             # sdf_b200/tetgrid.py topology, uploaded once
    mesh = extract_mesh(lattice, sdf, deform)               # marching tetrahedra            (nerf/renderer.py:868-874)
    fn, vn = mesh_normals(mesh)                             #                                 (:877-890)
    xyz, nrm, mask, rast = rasterize(mesh, vn, mvp, H, W)   # dr.rasterize + 2x dr.interpolate + safe_normalize (:893-903)
    nc, lap = mesh_losses(mesh, fn)                         # normal_consistency, laplacian_smooth_loss (:946-950)
All four are autograd Functions; gradients reach sdf / deform through the vertex attributes, the barycentrics and the regularisers.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import _lib, tetgrid

P = _lib.ptr


class TetLattice:
    """static lattice topology on the device + the per-step mesh buffers (capacity-sized; counts stay on the device)"""

    def __init__(self, tet_grid_size, device, n_cells=None, max_faces=None):
        self.tet_grid_size = int(tet_grid_size)
        n = tetgrid.cells_for(tet_grid_size) if n_cells is None else int(n_cells)
        verts, tets = tetgrid.make_tet_grid(n)
        edges, tet_edges = tetgrid.unique_edges(tets)
        self.n_cells, self.N, self.F, self.E = n, len(verts), len(tets), len(edges)
        dev = self.device = device
        self.pos = torch.from_numpy(verts).to(dev).contiguous()                     # [-1, 1]^3 (nerf/renderer.py:293)
        self.tets = torch.from_numpy(tets.astype(np.int32)).to(dev).contiguous()
        self.edges = torch.from_numpy(edges.astype(np.int32)).to(dev).contiguous()  # = self.all_edges of the reference (:305-308)
        self.tet_edges = torch.from_numpy(tet_edges.astype(np.int32)).to(dev).contiguous()
        # vertex / face buffers take the exact worst case (one vertex per lattice edge, two faces per tetrahedron: memory is not the constraint
        # on a 180 GB part); the regularisers sort 3 half-edges per face, so THEIR face budget is bounded at a quarter of the tetrahedra
        # (a closed surface through a 51^3 lattice cuts ~6e4 of its 1.6e6 tetrahedra) — faces beyond it would be left out of the two losses
        self.vcap, self.fcap = self.E, 2 * self.F
        self.reg_faces = int(max_faces or min(2 * self.F, max(8192, self.F // 4)))
        self.scratch = torch.empty(int(_lib.query('sdf_dmtet_scratch_ints', self.E, self.F)), device=dev, dtype=torch.int32)
        self.verts = torch.zeros(self.E, 3, device=dev)                             # extract writes by crossing-edge rank: E rows are always safe
        self.vert_edge = torch.zeros(self.E, device=dev, dtype=torch.int32)
        self.faces = torch.zeros(2 * self.F, 3, device=dev, dtype=torch.int32)
        self.counts = torch.zeros(8, device=dev, dtype=torch.int32)

    def mesh_counts(self):
        """(vertices, faces) of the last extraction — a host read, for logging / tests only"""
        c = self.counts[:2].tolist()
        return int(c[0]), int(c[1])


class Mesh:
    """view of the lattice's current mesh buffers; verts carries the autograd graph"""

    def __init__(self, lattice, verts):
        self.lattice, self.verts = lattice, verts
        self.faces, self.counts = lattice.faces, lattice.counts
        self._topo = None

    def topology(self):
        """sorted half-edge list of the first lattice.reg_faces faces (one library radix sort per step, shared by the regularisers and the
        antialiasing pass): dict(keys int64 [3R] sorted, face_of int32 [3R], face_adj int32 [R, 3])"""
        if self._topo is None:
            lat = self.lattice
            dev = lat.faces.device
            R = lat.reg_faces
            keys = torch.empty(3 * R, device=dev, dtype=torch.int64)
            face_of = torch.empty(3 * R, device=dev, dtype=torch.int32)
            st = _lib.stream()
            _lib.call('sdf_mesh_halfedge_keys', P(lat.faces), P(lat.counts), lat.vcap, R, P(keys), P(face_of), st)
            keys, order = torch.sort(keys)
            order = order.to(torch.int32)
            face_adj = torch.empty(R, 3, device=dev, dtype=torch.int32)
            _lib.call('sdf_mesh_face_adjacency', P(keys), P(order), P(lat.counts), R, P(face_adj), st)
            self._topo = dict(keys=keys, face_of=torch.div(order, 3, rounding_mode='floor').to(torch.int32), face_adj=face_adj)
        return self._topo


class _Extract(Function):
    @staticmethod
    def forward(ctx, sdf, deform, lat):
        sdf_c = sdf.detach().float().contiguous()
        def_c = deform.detach().float().contiguous() if deform is not None else None
        _lib.call('sdf_dmtet_extract', P(lat.pos), P(def_c), float(lat.tet_grid_size), P(sdf_c), P(lat.tets), P(lat.edges), P(lat.tet_edges), lat.N, lat.F,
                  lat.E, P(lat.verts), P(lat.vert_edge), P(lat.faces), P(lat.counts), P(lat.scratch), _lib.stream())
        ctx.lat = lat
        ctx.save_for_backward(sdf_c, def_c if def_c is not None else torch.empty(0))
        ctx.has_deform = def_c is not None
        return lat.verts.view(lat.E, 3)

    @staticmethod
    def backward(ctx, g_verts):
        lat = ctx.lat
        sdf_c, def_c = ctx.saved_tensors
        d_sdf = torch.zeros_like(sdf_c) if ctx.needs_input_grad[0] else None
        d_def = torch.zeros_like(def_c) if (ctx.has_deform and ctx.needs_input_grad[1]) else None
        _lib.call('sdf_dmtet_extract_backward', P(lat.pos), P(def_c) if ctx.has_deform else None, float(lat.tet_grid_size), P(sdf_c), P(lat.edges),
                  P(lat.vert_edge), P(lat.counts), lat.E, P(g_verts.contiguous()), P(d_sdf), P(d_def), _lib.stream())
        return d_sdf, d_def, None


def extract_mesh(lattice, sdf, deform=None):
    """marching tetrahedra of sdf on (lattice.pos + tanh(deform) / tet_grid_size) -> Mesh (vertex / face order = the reference's)"""
    return Mesh(lattice, _Extract.apply(sdf, deform, lattice))


class _Normals(Function):
    @staticmethod
    def forward(ctx, verts, lat):
        face_n = torch.empty(lat.fcap, 3, device=verts.device)
        vn_raw = torch.empty(lat.vcap, 3, device=verts.device)
        vn = torch.empty(lat.vcap, 3, device=verts.device)
        v = verts.detach().contiguous()
        _lib.call('sdf_mesh_normals_forward', P(v), P(lat.faces), P(lat.counts), lat.vcap, lat.fcap, P(face_n), P(vn_raw), P(vn), _lib.stream())
        ctx.lat = lat
        ctx.save_for_backward(v, vn_raw)
        return face_n, vn

    @staticmethod
    def backward(ctx, g_face_n, g_vn):
        lat = ctx.lat
        v, vn_raw = ctx.saved_tensors
        d_verts = torch.zeros_like(v)
        _lib.call('sdf_mesh_normals_backward', P(v), P(lat.faces), P(lat.counts), lat.fcap, P(vn_raw), P(g_vn.contiguous()) if g_vn is not None else None,
                  P(g_face_n.contiguous()) if g_face_n is not None else None, P(d_verts), _lib.stream())
        return d_verts, None


def mesh_normals(mesh):
    """-> face normals [fcap, 3], vertex normals [vcap, 3] (sum of the adjacent face normals, (0, 0, 1) where it vanishes)"""
    return _Normals.apply(mesh.verts, mesh.lattice)


class _Rasterize(Function):
    @staticmethod
    def forward(ctx, verts, vert_n, mvp, lat, H, W):
        dev = verts.device
        v, vn, m = verts.detach().contiguous(), vert_n.detach().contiguous(), mvp.detach().float().contiguous()
        clip = torch.empty(lat.vcap, 4, device=dev)
        zbuf = torch.empty(H * W, device=dev, dtype=torch.int64)
        rast = torch.empty(H, W, 4, device=dev)
        xyz, nrm, mask = torch.empty(H * W, 3, device=dev), torch.empty(H * W, 3, device=dev), torch.empty(H * W, device=dev)
        st = _lib.stream()
        _lib.call('sdf_mesh_clip_transform', P(v), P(lat.counts), lat.vcap, P(m), P(clip), st)
        _lib.call('sdf_mesh_rasterize', P(clip), P(lat.faces), P(lat.counts), lat.fcap, P(v), P(vn), H, W, P(zbuf), P(rast), P(xyz), P(nrm), P(mask), st)
        ctx.lat, ctx.hw = lat, (H, W)
        ctx.save_for_backward(v, vn, m, clip, rast)
        ctx.mark_non_differentiable(mask, rast, clip)
        return xyz, nrm, mask, rast, clip

    @staticmethod
    def backward(ctx, g_xyz, g_nrm, _gm, _gr, _gc):
        lat = ctx.lat
        H, W = ctx.hw
        v, vn, m, clip, rast = ctx.saved_tensors
        d_verts, d_vn = torch.zeros_like(v), torch.zeros_like(vn)
        if g_xyz is not None or g_nrm is not None:
            _lib.call('sdf_mesh_rasterize_backward', P(rast), P(clip), P(lat.faces), P(v), P(vn), P(m), H, W,
                      P(g_xyz.contiguous()) if g_xyz is not None else None, P(g_nrm.contiguous()) if g_nrm is not None else None, P(d_verts), P(d_vn),
                      _lib.stream())
        return d_verts, d_vn, None, None, None, None


def rasterize(mesh, vert_n, mvp, H, W, want_clip=False):
    """one view: mvp [4, 4] (device).  -> xyz [H*W, 3], unit normal [H*W, 3], coverage mask [H*W] in {0, 1}, rast [H, W, 4] = (u, v, z/w, id + 1)
    (+ the clip-space vertices [vcap, 4] with want_clip)"""
    out = _Rasterize.apply(mesh.verts, vert_n, mvp, mesh.lattice, H, W)
    return out if want_clip else out[:4]


class _MeshLosses(Function):
    @staticmethod
    def forward(ctx, verts, face_n, lat, topo):
        dev = verts.device
        v, fn = verts.detach().contiguous(), face_n.detach().contiguous()
        work = torch.empty(3 * lat.vcap + 4, device=dev)
        losses = torch.empty(2, device=dev)
        _lib.call('sdf_mesh_losses_forward', P(topo['keys']), P(topo['face_of']), P(lat.counts), lat.vcap, lat.reg_faces, P(fn), P(v), P(work), P(losses),
                  _lib.stream())
        ctx.lat, ctx.topo = lat, topo
        ctx.save_for_backward(fn, work)
        return losses

    @staticmethod
    def backward(ctx, g):
        lat, topo = ctx.lat, ctx.topo
        fn, work = ctx.saved_tensors
        d_verts = torch.zeros(lat.vcap, 3, device=fn.device)
        d_fn = torch.zeros_like(fn)
        _lib.call('sdf_mesh_losses_backward', P(topo['keys']), P(topo['face_of']), P(lat.counts), lat.vcap, lat.reg_faces, P(fn), P(work),
                  P(g.float().contiguous()), P(d_fn), P(d_verts), _lib.stream())
        return d_verts, d_fn, None, None


def mesh_losses(mesh, face_n):
    """-> tensor [2] = (normal_consistency, laplacian_smooth_loss) of the current mesh"""
    return _MeshLosses.apply(mesh.verts, face_n, mesh.lattice, mesh.topology())


def antialias_context(mesh, rast, clip, mvp):
    """what csrc/meshrast.cu's antialiasing pass needs besides the image: the rasteriser's outputs and the face adjacency"""
    lat = mesh.lattice
    return dict(rast=rast, clip=clip, faces=lat.faces, face_adj=mesh.topology()['face_adj'], adj_faces=lat.reg_faces, mvp=mvp.detach().float().contiguous(),
                vcap=lat.vcap)
