"""Tetrahedral grid of the DMTet stage.

The reference loads `tets/{tet_grid_size}_tets.npz` (nerf/renderer.py:291-294), files produced offline by the Quartet mesher
(tets/generate_tets.py): a body-centred lattice over the cube with ~2 vertices and ~12 tetrahedra per cubic cell (128 -> 277 410
vertices, 1 524 684 tetrahedra; 64 -> 36 562 / 192 492).  Those files are data of the reference repository and are not shipped here;
this module builds a lattice of the same family and the same size class analytically: corners of an n^3 cube grid plus the n^3 cell
centres, every cell split into 12 tetrahedra (centre + the two triangles of each face; the face diagonal starts at the face corner
with the smallest lattice index, so neighbouring cells agree on it and the grid is conforming).

Conventions kept from the reference so that marching tetrahedra (nerf/renderer.py:94-174) produces consistently wound triangles:
vertices cover [-1, 1]^3 (the reference stores [-0.5, 0.5] and multiplies by -2 at load), every tetrahedron has NEGATIVE signed volume
((v1 - v0) x (v2 - v0)) . (v3 - v0) like the reference's grid after that flip, indices are int64 / the edge list is the sorted unique
set the reference computes at nerf/renderer.py:305-308.
"""
import numpy as np


def cells_for(tet_grid_size):
    """cells per axis whose lattice has about as many vertices as the reference's grid of that name (0.4 * size: 128 -> 51 cells,
    273 259 vertices / 1 591 812 tetrahedra)"""
    return max(2, int(round(0.4 * int(tet_grid_size))))


def make_tet_grid(n):
    """-> vertices float32 [(n+1)^3 + n^3, 3] in [-1, 1], tetrahedra int64 [12 n^3, 4]"""
    n = int(n)
    k = np.arange(n + 1)
    ci, cj, ck = np.meshgrid(k, k, k, indexing="ij")
    corners = np.stack([ci, cj, ck], -1).reshape(-1, 3).astype(np.float64)
    c = np.arange(n)
    mi, mj, mk = np.meshgrid(c, c, c, indexing="ij")
    centres = np.stack([mi, mj, mk], -1).reshape(-1, 3).astype(np.float64) + 0.5
    verts = np.concatenate([corners, centres], 0) * (2.0 / n) - 1.0
    n_corner = (n + 1) ** 3

    def cid(i, j, kk):
        return (i * (n + 1) + j) * (n + 1) + kk

    cells = np.stack([mi, mj, mk], -1).reshape(-1, 3)
    centre_id = n_corner + (cells[:, 0] * n + cells[:, 1]) * n + cells[:, 2]
    tets = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in (0, 1):
            def corner(du, dv):
                off = np.zeros((len(cells), 3), dtype=np.int64)
                off[:, axis] = side
                off[:, u] = du
                off[:, v] = dv
                p = cells + off
                return cid(p[:, 0], p[:, 1], p[:, 2])
            q00, q10, q11, q01 = corner(0, 0), corner(1, 0), corner(1, 1), corner(0, 1)        # the face quad, in order around the face
            # diagonal from q00 (the smallest lattice index of the face, on either side of it): triangles (q00, q10, q11) and (q00, q11, q01)
            for tri in ((q00, q10, q11), (q00, q11, q01)):
                tets.append(np.stack([centre_id, tri[0], tri[1], tri[2]], -1))
    tets = np.concatenate(tets, 0).astype(np.int64)
    # orientation: negative signed volume everywhere (swap two vertices where it is positive)
    a, b, cc, d = (verts[tets[:, i]] for i in range(4))
    vol = np.einsum("ij,ij->i", np.cross(b - a, cc - a), d - a)
    flip = vol > 0
    tets[flip, 2], tets[flip, 3] = tets[flip, 3].copy(), tets[flip, 2].copy()
    return verts.astype(np.float32), tets


def unique_edges(tets):
    """sorted unique vertex pairs (a < b, lexicographic) of the six edges of every tetrahedron, and the [F, 6] edge id of each tetrahedron's
    edges in the reference's base order (0-1, 0-2, 0-3, 1-2, 1-3, 2-3; nerf/renderer.py:120,305-308)"""
    base = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3])
    e = tets[:, base].reshape(-1, 2)
    e = np.sort(e, 1)
    nv = int(tets.max()) + 1
    key = e[:, 0].astype(np.int64) * nv + e[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    edges = np.stack([uniq // nv, uniq % nv], -1)
    return edges.astype(np.int64), inv.reshape(-1, 6).astype(np.int64)
