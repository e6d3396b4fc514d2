// dmtet.cu — the mesh-extraction half of the DMTet stage (BASELINE config C5), sm_100a.
//
// The reference extracts the mesh with ~25 eager PyTorch launches per step, two of them torch.unique calls that synchronise with the host to size
// their outputs (nerf/renderer.py:128-174).  Here the lattice topology (sorted unique edges, the six edge ids of every tetrahedron) is built ONCE on
// the host (sdf_b200/tetgrid.py) and a step is five capacity-sized launches with device-side counts:
//   k_mt_edge_flags   one thread per lattice edge: does the signed distance change sign?  block-local exclusive scan of the flags
//   k_mt_tet_flags    one thread per tetrahedron: occupancy code, 0 / 1 / 2 triangles, block-local scans of the one- and two-triangle counts
//   k_mt_scan_blocks  one block: exclusive scan of the three per-block totals, writes (vertices, faces, one-triangle tets, two-triangle tets)
//   k_mt_emit_verts   crossing edges -> interpolated vertices, in sorted-edge order (= the reference's torch.unique order)
//   k_mt_emit_faces   triangle table -> faces, one-triangle tetrahedra first, then the two-triangle ones (the reference's torch.cat order)
// so that vertex and face arrays are index-for-index the reference's.  The backward scatters d(vertices) to d(sdf) and d(deform) with atomics.
// Also here: face / vertex normals (nerf/renderer.py:877-890) and the two mesh regularisers (normal consistency :209-222, uniform Laplacian
// :225-254) over a sorted half-edge list, forward and backward.  All HBM-bound streaming kernels (a few dozen bytes per element).
#include "common.cuh"

namespace {

__constant__ signed char c_tri_table[16][6] = {
    {-1, -1, -1, -1, -1, -1}, {1, 0, 2, -1, -1, -1}, {4, 0, 3, -1, -1, -1}, {1, 4, 2, 1, 3, 4},
    {3, 1, 5, -1, -1, -1},    {2, 3, 0, 2, 5, 3},    {1, 4, 0, 1, 5, 4},    {4, 2, 5, -1, -1, -1},
    {4, 5, 2, -1, -1, -1},    {4, 1, 0, 4, 5, 1},    {3, 2, 0, 3, 5, 2},    {1, 3, 5, -1, -1, -1},
    {4, 1, 2, 4, 3, 1},       {3, 0, 4, -1, -1, -1}, {2, 0, 1, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}};      // nerf/renderer.py:97-114
__constant__ signed char c_num_tri[16] = {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0};                  // nerf/renderer.py:115

constexpr int kScanThreads = 1024;

// exclusive scan of one int per thread over a 1024-thread block; *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += n; }
        warp_tot[lane] = winc - w;                       // exclusive prefix of the warp totals
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const int out = warp_tot[warp] + inc - v;
    __syncthreads();                                     // warp_tot is reused by the next call
    return out;
}

__global__ void __launch_bounds__(kScanThreads) k_mt_edge_flags(const float* __restrict__ sdf, const int* __restrict__ edges, int E,
                                                                int* __restrict__ edge_pref, int* __restrict__ block_sums) {
    __shared__ int tot;
    const int e = blockIdx.x * kScanThreads + threadIdx.x;
    int flag = 0;
    if (e < E) { const int2 ab = reinterpret_cast<const int2*>(edges)[e]; flag = (sdf[ab.x] > 0.f) != (sdf[ab.y] > 0.f); }
    const int pref = block_exclusive_scan(flag, &tot);
    if (e < E) edge_pref[e] = flag ? pref : -1;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) k_mt_tet_flags(const float* __restrict__ sdf, const int* __restrict__ tets, int F,
                                                               unsigned char* __restrict__ tet_code, int* __restrict__ tet_pref,
                                                               int* __restrict__ bs1, int* __restrict__ bs2) {
    __shared__ int tot;
    const int f = blockIdx.x * kScanThreads + threadIdx.x;
    int packed = 0, code = 0;
    if (f < F) {
        const int4 t = reinterpret_cast<const int4*>(tets)[f];
        code = (sdf[t.x] > 0.f ? 1 : 0) | (sdf[t.y] > 0.f ? 2 : 0) | (sdf[t.z] > 0.f ? 4 : 0) | (sdf[t.w] > 0.f ? 8 : 0);      // nerf/renderer.py:157-158
        const int nt = c_num_tri[code];
        packed = (nt == 1 ? 1 : 0) | (nt == 2 ? (1 << 16) : 0);                  // both counts in one scan (<= 1024 each per block)
    }
    const int pref = block_exclusive_scan(packed, &tot);
    if (f < F) { tet_code[f] = (unsigned char)code; tet_pref[f] = pref; }
    if (threadIdx.x == 0) { bs1[blockIdx.x] = tot & 0xffff; bs2[blockIdx.x] = tot >> 16; }
}

// in-place exclusive scans of the three per-block total arrays; counts = (vertices, faces, one-triangle tets, two-triangle tets)
__global__ void __launch_bounds__(kScanThreads) k_mt_scan_blocks(int* __restrict__ bs_e, int nbe, int* __restrict__ bs1, int* __restrict__ bs2, int nbt,
                                                                 int* __restrict__ counts) {
    __shared__ int tot;
    int totals[3];
    for (int which = 0; which < 3; which++) {
        int* a = which == 0 ? bs_e : (which == 1 ? bs1 : bs2);
        const int n = which == 0 ? nbe : nbt;
        int carry = 0;
        for (int base = 0; base < n; base += kScanThreads) {
            const int i = base + threadIdx.x;
            const int v = i < n ? a[i] : 0;
            const int p = block_exclusive_scan(v, &tot);
            if (i < n) a[i] = carry + p;
            carry += tot;
            __syncthreads();
        }
        totals[which] = carry;
    }
    if (threadIdx.x == 0) { counts[0] = totals[0]; counts[1] = totals[1] + 2 * totals[2]; counts[2] = totals[1]; counts[3] = totals[2]; }
}

// lattice vertex position: pos + tanh(deform) / tet_grid_size   (nerf/renderer.py:872,874)
__device__ __forceinline__ void lattice_pos(float p[3], const float* __restrict__ pos, const float* __restrict__ deform, float inv_scale_div, int i) {
#pragma unroll
    for (int d = 0; d < 3; d++) p[d] = deform ? __fadd_rn(pos[3 * i + d], __fdiv_rn(tanhf(deform[3 * i + d]), inv_scale_div)) : pos[3 * i + d];
}

__global__ void __launch_bounds__(kScanThreads) k_mt_emit_verts(const float* __restrict__ pos, const float* __restrict__ deform, float tet_grid_size,
                                                                const float* __restrict__ sdf, const int* __restrict__ edges, int E,
                                                                const int* __restrict__ edge_pref, const int* __restrict__ bs_e,
                                                                float* __restrict__ verts, int* __restrict__ vert_edge, int* __restrict__ edge_vid) {
    const int e = blockIdx.x * kScanThreads + threadIdx.x;
    if (e >= E) return;
    const int pref = edge_pref[e];
    if (pref < 0) { edge_vid[e] = -1; return; }
    const int vid = bs_e[blockIdx.x] + pref;
    const int2 ab = reinterpret_cast<const int2*>(edges)[e];
    float pa[3], pb[3];
    lattice_pos(pa, pos, deform, tet_grid_size, ab.x);
    lattice_pos(pb, pos, deform, tet_grid_size, ab.y);
    // nerf/renderer.py:146-153: weights (-s_b, s_a) / (s_a - s_b); products and the sum rounded separately like the eager ops
    const float sa = sdf[ab.x], nsb = -sdf[ab.y];
    const float den = __fadd_rn(sa, nsb);
    const float wa = __fdiv_rn(nsb, den), wb = __fdiv_rn(sa, den);
#pragma unroll
    for (int d = 0; d < 3; d++) verts[3 * (size_t)vid + d] = __fadd_rn(__fmul_rn(pa[d], wa), __fmul_rn(pb[d], wb));
    vert_edge[vid] = e;
    edge_vid[e] = vid;
}

__global__ void __launch_bounds__(kScanThreads) k_mt_emit_faces(const int* __restrict__ tet_edges, const unsigned char* __restrict__ tet_code,
                                                                const int* __restrict__ tet_pref, const int* __restrict__ bs1, const int* __restrict__ bs2,
                                                                const int* __restrict__ counts, const int* __restrict__ edge_vid, int F,
                                                                int* __restrict__ faces) {
    const int f = blockIdx.x * kScanThreads + threadIdx.x;
    if (f >= F) return;
    const int code = tet_code[f], nt = c_num_tri[code];
    if (nt == 0) return;
    const int pref = tet_pref[f];
    const int first = nt == 1 ? bs1[blockIdx.x] + (pref & 0xffff) : counts[2] + 2 * (bs2[blockIdx.x] + (pref >> 16));
    const int* te = tet_edges + 6 * (size_t)f;
    for (int k = 0; k < 3 * nt; k++) faces[3 * (size_t)first + k] = edge_vid[te[c_tri_table[code][k]]];
}

// d(verts) -> d(sdf), d(deform): v = p_a w_a + p_b w_b, w_a = -s_b / (s_a - s_b), w_b = s_a / (s_a - s_b)
__global__ void k_mt_backward(const float* __restrict__ pos, const float* __restrict__ deform, float tet_grid_size, const float* __restrict__ sdf,
                              const int* __restrict__ edges, const int* __restrict__ vert_edge, const int* __restrict__ counts,
                              const float* __restrict__ d_verts, float* __restrict__ d_sdf, float* __restrict__ d_deform) {
    const int nv = counts[0];
    for (int vid = blockIdx.x * blockDim.x + threadIdx.x; vid < nv; vid += gridDim.x * blockDim.x) {
        const int2 ab = reinterpret_cast<const int2*>(edges)[vert_edge[vid]];
        float pa[3], pb[3];
        lattice_pos(pa, pos, deform, tet_grid_size, ab.x);
        lattice_pos(pb, pos, deform, tet_grid_size, ab.y);
        const float sa = sdf[ab.x], sb = sdf[ab.y];
        const float den = sa - sb, inv = 1.f / den, inv2 = inv * inv;
        const float wa = -sb * inv, wb = sa * inv;
        float gsa = 0.f, gsb = 0.f;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float g = d_verts[3 * (size_t)vid + d];
            const float diff = pa[d] - pb[d];
            gsa += g * diff * sb * inv2;                 // dv/ds_a = (p_a - p_b) s_b / den^2
            gsb -= g * diff * sa * inv2;                 // dv/ds_b = (p_b - p_a) s_a / den^2
            if (d_deform) {
                const float ta = tanhf(deform[3 * ab.x + d]), tb = tanhf(deform[3 * ab.y + d]);
                atomicAdd(&d_deform[3 * ab.x + d], g * wa * (1.f - ta * ta) / tet_grid_size);
                atomicAdd(&d_deform[3 * ab.y + d], g * wb * (1.f - tb * tb) / tet_grid_size);
            }
        }
        if (d_sdf) { atomicAdd(&d_sdf[ab.x], gsa); atomicAdd(&d_sdf[ab.y], gsb); }
    }
}

// ------------------------------------------------------------------ normals
__device__ __forceinline__ void load3(float o[3], const float* p, int i) { o[0] = p[3 * (size_t)i]; o[1] = p[3 * (size_t)i + 1]; o[2] = p[3 * (size_t)i + 2]; }
__device__ __forceinline__ void cross3(float o[3], const float a[3], const float b[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// face normal = safe_normalize(cross(v1 - v0, v2 - v0)); vertex normal sums by atomics (nerf/renderer.py:877-888)
__global__ void k_face_normals(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ counts,
                               float* __restrict__ face_n, float* __restrict__ vert_n) {
    const int nf = counts[1];
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
        const int i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
        float v0[3], v1[3], v2[3], e1[3], e2[3], c[3];
        load3(v0, verts, i0); load3(v1, verts, i1); load3(v2, verts, i2);
#pragma unroll
        for (int d = 0; d < 3; d++) { e1[d] = v1[d] - v0[d]; e2[d] = v2[d] - v0[d]; }
        cross3(c, e1, e2);
        const float inv = rsqrtf(fmaxf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2], 1e-20f));
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float n = c[d] * inv;
            face_n[3 * (size_t)f + d] = n;
            atomicAdd(&vert_n[3 * (size_t)i0 + d], n); atomicAdd(&vert_n[3 * (size_t)i1 + d], n); atomicAdd(&vert_n[3 * (size_t)i2 + d], n);
        }
    }
}

// vn = where(|vn|^2 > 1e-20, vn, (0, 0, 1))  (nerf/renderer.py:890); `vert_n_raw` keeps the sums for the backward's mask
__global__ void k_vert_normals_finish(const float* __restrict__ vert_n_raw, const int* __restrict__ counts, float* __restrict__ vert_n) {
    const int nv = counts[0];
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
        float n[3];
        load3(n, vert_n_raw, v);
        const bool ok = n[0] * n[0] + n[1] * n[1] + n[2] * n[2] > 1e-20f;
        vert_n[3 * (size_t)v] = ok ? n[0] : 0.f; vert_n[3 * (size_t)v + 1] = ok ? n[1] : 0.f; vert_n[3 * (size_t)v + 2] = ok ? n[2] : 1.f;
    }
}

__global__ void k_face_normals_bwd(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ counts,
                                   const float* __restrict__ vert_n_raw, const float* __restrict__ d_vert_n, const float* __restrict__ d_face_n,
                                   float* __restrict__ d_verts) {
    const int nf = counts[1];
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
        const int idx[3] = {faces[3 * (size_t)f], faces[3 * (size_t)f + 1], faces[3 * (size_t)f + 2]};
        float g[3] = {0.f, 0.f, 0.f};
        if (d_face_n) load3(g, d_face_n, f);
        if (d_vert_n) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float r[3], dv[3];
                load3(r, vert_n_raw, idx[k]);
                if (r[0] * r[0] + r[1] * r[1] + r[2] * r[2] > 1e-20f) { load3(dv, d_vert_n, idx[k]); g[0] += dv[0]; g[1] += dv[1]; g[2] += dv[2]; }
            }
        }
        float v0[3], v1[3], v2[3], e1[3], e2[3], c[3];
        load3(v0, verts, idx[0]); load3(v1, verts, idx[1]); load3(v2, verts, idx[2]);
#pragma unroll
        for (int d = 0; d < 3; d++) { e1[d] = v1[d] - v0[d]; e2[d] = v2[d] - v0[d]; }
        cross3(c, e1, e2);
        const float len2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        const float inv = rsqrtf(fmaxf(len2, 1e-20f));
        float dc[3];
        if (len2 > 1e-20f) {
            const float n[3] = {c[0] * inv, c[1] * inv, c[2] * inv};
            const float ng = n[0] * g[0] + n[1] * g[1] + n[2] * g[2];
#pragma unroll
            for (int d = 0; d < 3; d++) dc[d] = (g[d] - n[d] * ng) * inv;
        } else {
#pragma unroll
            for (int d = 0; d < 3; d++) dc[d] = g[d] * inv;              // the clamped denominator is a constant
        }
        float de1[3], de2[3];
        cross3(de1, e2, dc);                                           // c = e1 x e2: d e1 = e2 x dc, d e2 = dc x e1
        cross3(de2, dc, e1);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            atomicAdd(&d_verts[3 * (size_t)idx[0] + d], -(de1[d] + de2[d]));
            atomicAdd(&d_verts[3 * (size_t)idx[1] + d], de1[d]);
            atomicAdd(&d_verts[3 * (size_t)idx[2] + d], de2[d]);
        }
    }
}

// ------------------------------------------------------------------ mesh regularisers over the sorted half-edge list
// half-edge key of face f, corner k (edge i_k -> i_{k+1}): ((min * cap + max) << 1) | (i_k > i_{k+1}); unused slots get the maximum key
__global__ void k_halfedge_keys(const int* __restrict__ faces, const int* __restrict__ counts, long long vcap, int fcap,
                                long long* __restrict__ keys, int* __restrict__ payload) {
    const int nf = min(counts[1], fcap);
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < 3 * fcap; h += gridDim.x * blockDim.x) {
        const int f = h / 3, k = h % 3;
        long long key = 0x7fffffffffffffffll;
        if (f < nf) {
            const int a = faces[3 * (size_t)f + k], b = faces[3 * (size_t)f + (k + 1) % 3];
            const long long lo = a < b ? a : b, hi = a < b ? b : a;
            key = ((lo * vcap + hi) << 1) | (a > b ? 1 : 0);
        }
        keys[h] = key;
        payload[h] = f;
    }
}

// One thread per sorted half-edge that starts a run of equal undirected edges.  Forward: sums of (1 - clamp(n_t0 . n_t1)) and the number of unique
// edges (normal consistency), Laplacian rows acc[i] += v_i - v_j, acc[j] += v_j - v_i.  The faces paired on an edge are the one holding it in
// ascending direction (t0) and the one holding it descending (t1); a missing side pairs with face 0 like the reference's zero-initialised table
// (nerf/renderer.py:198-205).
__global__ void k_mesh_edges_fwd(const long long* __restrict__ keys, const int* __restrict__ face_of, const int* __restrict__ counts, int fcap, long long vcap,
                                 const float* __restrict__ face_n, const float* __restrict__ verts, float* __restrict__ lap_acc,
                                 float* __restrict__ sums /* [0] sum of terms, [1] unique edges */) {
    const int nh = 3 * min(counts[1], fcap);
    float term_sum = 0.f, edge_cnt = 0.f;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x) {
        const long long e = keys[h] >> 1;
        if (h > 0 && (keys[h - 1] >> 1) == e) continue;
        int t0 = 0, t1 = 0;
        for (int j = h; j < nh && (keys[j] >> 1) == e; j++) { if (keys[j] & 1) t1 = face_of[j]; else t0 = face_of[j]; }
        float n0[3], n1[3];
        load3(n0, face_n, t0); load3(n1, face_n, t1);
        const float d = fminf(fmaxf(n0[0] * n1[0] + n0[1] * n1[1] + n0[2] * n1[2], -1.f), 1.f);
        term_sum += fabsf(1.f - d);
        edge_cnt += 1.f;
        const int i = (int)(e / vcap), j2 = (int)(e % vcap);
        float vi[3], vj[3];
        load3(vi, verts, i); load3(vj, verts, j2);
#pragma unroll
        for (int dd = 0; dd < 3; dd++) { atomicAdd(&lap_acc[3 * (size_t)i + dd], vi[dd] - vj[dd]); atomicAdd(&lap_acc[3 * (size_t)j2 + dd], vj[dd] - vi[dd]); }
    }
    term_sum = warp_sum(term_sum); edge_cnt = warp_sum(edge_cnt);
    if ((threadIdx.x & 31) == 0 && edge_cnt > 0.f) { atomicAdd(&sums[0], term_sum); atomicAdd(&sums[1], edge_cnt); }
}

// losses[0] = normal consistency = sums[0] / sums[1]; losses[1] = Laplacian = mean_v |acc_v|
__global__ void k_mesh_losses_finish(const float* __restrict__ lap_acc, const int* __restrict__ counts, const float* __restrict__ sums, float* __restrict__ lap_sum,
                                     float* __restrict__ losses, unsigned int* __restrict__ ticket) {
    const int nv = counts[0];
    float s = 0.f;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
        float a[3];
        load3(a, lap_acc, v);
        s += sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(lap_sum, s);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            losses[0] = sums[1] > 0.f ? sums[0] / sums[1] : 0.f;
            losses[1] = nv > 0 ? *reinterpret_cast<volatile float*>(lap_sum) / (float)nv : 0.f;
            *ticket = 0;
        }
    }
}

// backward of both losses: g[2] = upstream gradients of (normal consistency, Laplacian) on the device
__global__ void k_mesh_edges_bwd(const long long* __restrict__ keys, const int* __restrict__ face_of, const int* __restrict__ counts, int fcap, long long vcap,
                                 const float* __restrict__ face_n, const float* __restrict__ lap_acc, const float* __restrict__ sums, const float* __restrict__ g,
                                 float* __restrict__ d_face_n, float* __restrict__ d_verts) {
    const int nh = 3 * min(counts[1], fcap), nv = counts[0];
    const float g_nc = g[0], g_lap = g[1];
    const float gn = sums[1] > 0.f ? g_nc / sums[1] : 0.f, gl = nv > 0 ? g_lap / (float)nv : 0.f;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x) {
        const long long e = keys[h] >> 1;
        if (h > 0 && (keys[h - 1] >> 1) == e) continue;
        int t0 = 0, t1 = 0;
        for (int j = h; j < nh && (keys[j] >> 1) == e; j++) { if (keys[j] & 1) t1 = face_of[j]; else t0 = face_of[j]; }
        if (gn != 0.f) {
            float n0[3], n1[3];
            load3(n0, face_n, t0); load3(n1, face_n, t1);
            const float d = n0[0] * n1[0] + n0[1] * n1[1] + n0[2] * n1[2];
            if (d >= -1.f && d <= 1.f) {             // clamp passes the gradient on the closed interval; |1 - d| = 1 - d there
#pragma unroll
                for (int dd = 0; dd < 3; dd++) { atomicAdd(&d_face_n[3 * (size_t)t0 + dd], -gn * n1[dd]); atomicAdd(&d_face_n[3 * (size_t)t1 + dd], -gn * n0[dd]); }
            }
        }
        if (gl != 0.f) {
            const int i = (int)(e / vcap), j2 = (int)(e % vcap);
            float ai[3], aj[3];
            load3(ai, lap_acc, i); load3(aj, lap_acc, j2);
            const float li = sqrtf(ai[0] * ai[0] + ai[1] * ai[1] + ai[2] * ai[2]), lj = sqrtf(aj[0] * aj[0] + aj[1] * aj[1] + aj[2] * aj[2]);
#pragma unroll
            for (int dd = 0; dd < 3; dd++) {
                const float gi = li > 0.f ? gl * ai[dd] / li : 0.f, gj = lj > 0.f ? gl * aj[dd] / lj : 0.f;
                atomicAdd(&d_verts[3 * (size_t)i + dd], gi - gj);
                atomicAdd(&d_verts[3 * (size_t)j2 + dd], gj - gi);
            }
        }
    }
}

inline int grid_for(long long n, int threads) { return (int)max(1ll, min((n + threads - 1) / threads, (long long)sdf_num_sms() * 8)); }

}  // namespace

// ------------------------------------------------------------------ C ABI
// scratch ints needed by sdf_dmtet_extract for a lattice with E edges and F tetrahedra
SDF_API long long sdf_dmtet_scratch_ints(int E, int F) {
    const long long nbe = (E + kScanThreads - 1) / kScanThreads, nbt = (F + kScanThreads - 1) / kScanThreads;
    return (long long)E /*edge_pref*/ + E /*edge_vid*/ + F /*tet_pref*/ + (F + 3) / 4 /*tet_code bytes*/ + nbe + 2 * nbt + 16;
}

// Marching tetrahedra (nerf/renderer.py:128-174) on a fixed lattice.  pos [N,3] lattice vertices, deform [N,3] raw parameter or NULL
// (position = pos + tanh(deform) / tet_grid_size), sdf [N]; tets [F,4], edges [E,2] sorted unique (a < b), tet_edges [F,6] (sdf_b200/tetgrid.py).
// Outputs (capacity-sized, device-side counts): verts [E,3], vert_edge [E] (lattice edge of each vertex, for the backward), faces [2F,3],
// counts[4] = (vertices, faces, one-triangle tets, two-triangle tets).
SDF_API int sdf_dmtet_extract(const float* pos, const float* deform, float tet_grid_size, const float* sdf, const int* tets, const int* edges,
                              const int* tet_edges, int N, int F, int E, float* verts, int* vert_edge, int* faces, int* counts, int* scratch, void* stream) {
    SDF_CHECK_ARG(pos && sdf && tets && edges && tet_edges && verts && vert_edge && faces && counts && scratch && N > 0 && F > 0 && E > 0,
                  "dmtet_extract: null pointer or empty lattice");
    cudaStream_t st = (cudaStream_t)stream;
    const int nbe = (E + kScanThreads - 1) / kScanThreads, nbt = (F + kScanThreads - 1) / kScanThreads;
    int* edge_pref = scratch;
    int* edge_vid = edge_pref + E;
    int* tet_pref = edge_vid + E;
    unsigned char* tet_code = reinterpret_cast<unsigned char*>(tet_pref + F);
    int* bs_e = tet_pref + F + (F + 3) / 4;
    int* bs1 = bs_e + nbe;
    int* bs2 = bs1 + nbt;
    k_mt_edge_flags<<<nbe, kScanThreads, 0, st>>>(sdf, edges, E, edge_pref, bs_e);
    k_mt_tet_flags<<<nbt, kScanThreads, 0, st>>>(sdf, tets, F, tet_code, tet_pref, bs1, bs2);
    k_mt_scan_blocks<<<1, kScanThreads, 0, st>>>(bs_e, nbe, bs1, bs2, nbt, counts);
    k_mt_emit_verts<<<nbe, kScanThreads, 0, st>>>(pos, deform, tet_grid_size, sdf, edges, E, edge_pref, bs_e, verts, vert_edge, edge_vid);
    k_mt_emit_faces<<<nbt, kScanThreads, 0, st>>>(tet_edges, tet_code, tet_pref, bs1, bs2, counts, edge_vid, F, faces);
    SDF_CHECK_LAUNCH("dmtet_extract");
    return SDF_OK;
}

// d_verts [E,3] (rows >= counts[0] ignored) -> d_sdf [N] and d_deform [N,3] (either may be NULL), ACCUMULATED
SDF_API int sdf_dmtet_extract_backward(const float* pos, const float* deform, float tet_grid_size, const float* sdf, const int* edges, const int* vert_edge,
                                       const int* counts, int E, const float* d_verts, float* d_sdf, float* d_deform, void* stream) {
    SDF_CHECK_ARG(pos && sdf && edges && vert_edge && counts && d_verts && (!d_deform || deform), "dmtet_extract_backward: null pointer");
    k_mt_backward<<<grid_for(E / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(pos, deform, tet_grid_size, sdf, edges, vert_edge, counts, d_verts, d_sdf, d_deform);
    SDF_CHECK_LAUNCH("dmtet_extract_backward");
    return SDF_OK;
}

// face normals [fcap,3], raw vertex-normal sums [vcap,3] and the vertex normals the renderer interpolates [vcap,3] (nerf/renderer.py:877-890)
SDF_API int sdf_mesh_normals_forward(const float* verts, const int* faces, const int* counts, int vcap, int fcap, float* face_n, float* vert_n_raw, float* vert_n,
                                     void* stream) {
    SDF_CHECK_ARG(verts && faces && counts && face_n && vert_n_raw && vert_n, "mesh_normals_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SDF_CHECK_CUDA(cudaMemsetAsync(vert_n_raw, 0, sizeof(float) * 3 * (size_t)vcap, st));
    k_face_normals<<<grid_for(fcap / 4 + 1, 256), 256, 0, st>>>(verts, faces, counts, face_n, vert_n_raw);
    k_vert_normals_finish<<<grid_for(vcap / 4 + 1, 256), 256, 0, st>>>(vert_n_raw, counts, vert_n);
    SDF_CHECK_LAUNCH("mesh_normals_forward");
    return SDF_OK;
}

// d_vert_n [vcap,3] and / or d_face_n [fcap,3] -> d_verts [vcap,3], ACCUMULATED
SDF_API int sdf_mesh_normals_backward(const float* verts, const int* faces, const int* counts, int fcap, const float* vert_n_raw, const float* d_vert_n,
                                      const float* d_face_n, float* d_verts, void* stream) {
    SDF_CHECK_ARG(verts && faces && counts && vert_n_raw && d_verts && (d_vert_n || d_face_n), "mesh_normals_backward: null pointer");
    k_face_normals_bwd<<<grid_for(fcap / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(verts, faces, counts, vert_n_raw, d_vert_n, d_face_n, d_verts);
    SDF_CHECK_LAUNCH("mesh_normals_backward");
    return SDF_OK;
}

// half-edge keys [3*fcap] int64 + owning face [3*fcap] of the first min(faces, fcap) faces (fcap here = the regulariser's face budget, which may be
// smaller than the face buffer: it sizes the sort); the caller sorts (keys, payload) by key and passes them to the two calls below
SDF_API int sdf_mesh_halfedge_keys(const int* faces, const int* counts, int vcap, int fcap, long long* keys, int* face_of, void* stream) {
    SDF_CHECK_ARG(faces && counts && keys && face_of, "mesh_halfedge_keys: null pointer");
    k_halfedge_keys<<<grid_for(3ll * fcap, 256), 256, 0, (cudaStream_t)stream>>>(faces, counts, (long long)vcap, fcap, keys, face_of);
    SDF_CHECK_LAUNCH("mesh_halfedge_keys");
    return SDF_OK;
}

// losses[0] = normal_consistency (nerf/renderer.py:209-222), losses[1] = laplacian_smooth_loss (:248-254).  work: floats [3*vcap + 4] (Laplacian rows,
// sums) + one uint ticket, all zeroed here.
SDF_API int sdf_mesh_losses_forward(const long long* sorted_keys, const int* sorted_face_of, const int* counts, int vcap, int fcap, const float* face_n,
                                    const float* verts, float* work, float* losses, void* stream) {
    SDF_CHECK_ARG(sorted_keys && sorted_face_of && counts && face_n && verts && work && losses, "mesh_losses_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SDF_CHECK_CUDA(cudaMemsetAsync(work, 0, sizeof(float) * (3 * (size_t)vcap + 4), st));
    float* lap_acc = work;
    float* sums = work + 3 * (size_t)vcap;
    k_mesh_edges_fwd<<<grid_for(3ll * fcap / 4 + 1, 256), 256, 0, st>>>(sorted_keys, sorted_face_of, counts, fcap, (long long)vcap, face_n, verts, lap_acc, sums);
    k_mesh_losses_finish<<<grid_for(vcap / 4 + 1, 256), 256, 0, st>>>(lap_acc, counts, sums, sums + 2, losses, reinterpret_cast<unsigned int*>(sums + 3));
    SDF_CHECK_LAUNCH("mesh_losses_forward");
    return SDF_OK;
}

// gradients of g[0] * losses[0] + g[1] * losses[1] (g: 2 floats on the device): d_face_n [fcap,3] and d_verts [vcap,3], ACCUMULATED; `work` as left by the forward
SDF_API int sdf_mesh_losses_backward(const long long* sorted_keys, const int* sorted_face_of, const int* counts, int vcap, int fcap, const float* face_n,
                                     const float* work, const float* g, float* d_face_n, float* d_verts, void* stream) {
    SDF_CHECK_ARG(sorted_keys && sorted_face_of && counts && face_n && work && g && d_face_n && d_verts, "mesh_losses_backward: null pointer");
    k_mesh_edges_bwd<<<grid_for(3ll * fcap / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(sorted_keys, sorted_face_of, counts, fcap, (long long)vcap, face_n, work,
                                                                                       work + 3 * (size_t)vcap, g, d_face_n, d_verts);
    SDF_CHECK_LAUNCH("mesh_losses_backward");
    return SDF_OK;
}
