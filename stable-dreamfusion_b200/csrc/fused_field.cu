// fused_field.cu — fused radiance-field evaluation for the -O backbone, sm_100a.
//
// One kernel replaces, per point, the reference chain (nerf/network_grid.py:68-130):
//   GridEncoder (16 level-kernels + fp16 table cast + permute) -> 3 cuBLAS GEMMs + 2 ReLU
//   -> trunc_exp / sigmoid / density_blob -> 6 more of the same at x +- eps (finite-difference
//   normal, network_grid.py:81-102) -> safe_normalize / nan_to_num -> Lambertian shading.
// A warp owns 16 samples at a time.  Each lane gathers the hash-grid corners of (2 rows x 4
// levels) so that the interpolated features land directly in mma.sync A-fragment registers;
// the 32-64-64-4 MLP runs on tensor cores (m16n8k16, fp16 in / fp32 accumulate) with the
// activations chained through registers; the 7 stencil densities of a sample end up in the
// same lane, which finishes normal + shading.  Nothing but xyz in and (sigma, rgb, normal)
// out touches HBM; the 24 MB fp16 table is served from L2.  One hashed/dense branch per level and branch-free
// corner arithmetic keep all 16 gathers of a level pair in flight; aligned x-neighbour corner pairs come in one
// 8-byte load; the stencil loop is rolled and the +-eps points re-use the centre cell's coarse-level corners.
//
// Algorithmic bytes: forward 540 B per point-eval (12 B xyz + 128 half2
// corner reads + 16 B out).  Roofline: HBM.
#include "field_common.cuh"
#include <cstdlib>

using namespace field;

namespace {

constexpr float kFdEps = 1e-2f;     // finite_difference_normal epsilon (network_grid.py:81)

constexpr int kAuxStride = 10;      // per-sample forward stash: sigma at the 7 stencil points + albedo

enum Shading { kAlbedo = 0, kLambertian = 1, kTextureless = 2, kNormal = 3 };

__device__ __forceinline__ float nan_to_num(float v) {
    if (isnan(v)) return 0.f;
    if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return v;
}

// stencil point p of a sample at x (clamped to the scene box, network_grid.py:83-88)
__device__ __forceinline__ void stencil_point(float out[3], const float x[3], int p, float bound) {
    // p = 0: the sample itself; p = 1..6: +eps / -eps along x, y, z (nerf/network_grid.py:92-100).  Branch-free so that a
    // rolled stencil loop keeps the point in registers.
    const int axis = p > 0 ? (p - 1) >> 1 : -1;
    const float e = ((p - 1) & 1) ? -kFdEps : kFdEps;
#pragma unroll
    for (int d = 0; d < 3; d++) out[d] = (d == axis) ? fminf(fmaxf(x[d] + e, -bound), bound) : x[d];
}

__device__ __forceinline__ bool to_unit(float u[3], const float x[3], float bound) {
    const float inv = 1.f / (2.f * bound);
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; d++) { u[d] = (x[d] + bound) * inv; ok &= (u[d] >= 0.f && u[d] <= 1.f); }
    return ok;
}

template <int SHADING, int OCC>
__global__ void __launch_bounds__(256, OCC)
k_field_forward(FieldParams p, const float* __restrict__ xyzs, const float* __restrict__ light_d, int light_per_sample,
                float ratio, uint32_t M_cap, const int* __restrict__ m_dev,
                float* __restrict__ sigmas, float* __restrict__ colors, float* __restrict__ normals, float* __restrict__ aux,
                uint4* __restrict__ feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WeightsSmem& s = *reinterpret_cast<WeightsSmem*>(smem_raw);
    load_weights(s, p);
    __syncthreads();

    const uint32_t M = m_dev ? min((uint32_t)max(*m_dev, 0), M_cap) : M_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NP = (SHADING == kAlbedo) ? 1 : 7;
    const uint32_t n_groups = (M + 15) / 16;

    for (uint32_t grp = blockIdx.x * (blockDim.x >> 5) + warp; grp < n_groups; grp += gridDim.x * (blockDim.x >> 5)) {
        const uint32_t sa = grp * 16 + g, sb = sa + 8;
        const bool ina = sa < M, inb = sb < M;
        float xa[3] = {0.f, 0.f, 0.f}, xb[3] = {0.f, 0.f, 0.f};
        if (ina) { xa[0] = xyzs[(size_t)sa * 3]; xa[1] = xyzs[(size_t)sa * 3 + 1]; xa[2] = xyzs[(size_t)sa * 3 + 2]; }
        if (inb) { xb[0] = xyzs[(size_t)sb * 3]; xb[1] = xyzs[(size_t)sb * 3 + 1]; xb[2] = xyzs[(size_t)sb * 3 + 2]; }

        // The stencil loop stays rolled (7 unrolled copies of gather + MLP are ~400 KB of SASS: instruction-cache misses showed
        // up as the third stall reason): densities go straight to the stash, the finite differences accumulate as they come.
        float sig0_a = 0.f, sig0_b = 0.f;
        float nda[3] = {0.f, 0.f, 0.f}, ndb[3] = {0.f, 0.f, 0.f};      // sigma(+eps) - sigma(-eps) per axis
        float alb_a[3] = {0.f, 0.f, 0.f}, alb_b[3] = {0.f, 0.f, 0.f};
        CornerCache cache;
#pragma unroll 1
        for (int sp = 0; sp < NP; sp++) {
            float pa[3], pb[3], ua[3], ub[3];
            stencil_point(pa, xa, sp, p.bound);
            stencil_point(pb, xb, sp, p.bound);
            const bool va = to_unit(ua, pa, p.bound) && ina, vb = to_unit(ub, pb, p.bound) && inb;
            uint32_t a0[2][4];
            if (NP > 1) encode_rows_cached(a0, s, p, lane, ua, va, ub, vb, cache, sp == 0);
            else encode_rows(a0, s, p, lane, ua, va, ub, vb);
            if (feat) {
                // feature stash for the backward: the lane's A fragments (2 rows x 4 levels x 2 features, fp16) as two 16-byte stores;
                // a warp writes 1 KB contiguous per stencil point.  Re-gathering them in the backward costs ~1 ms of scattered L1/L2
                // traffic at 432 k samples; streaming them through HBM costs ~0.03 ms each way.
                uint4* f = feat + ((size_t)grp * NP + sp) * 64 + lane * 2;
                f[0] = make_uint4(a0[0][0], a0[0][1], a0[0][2], a0[0][3]);
                f[1] = make_uint4(a0[1][0], a0[1][1], a0[1][2], a0[1][3]);
            }
            float h[4];
            mlp_forward<false>(h, a0, s, lane, nullptr, nullptr);
            // lanes t==0: h[0],h[1] = logits 0,1 of row g ; h[2],h[3] = of row g+8.  lanes t==1: logits 2,3.
            const float sga = __expf(round_h(h[0]) + blob(p, pa));
            const float sgb = __expf(round_h(h[2]) + blob(p, pb));
            if (aux && t == 0) {      // saved for the backward pass: the stencil densities
                if (ina) aux[(size_t)sa * kAuxStride + sp] = sga;
                if (inb) aux[(size_t)sb * kAuxStride + sp] = sgb;
            }
            if (sp == 0) {
                sig0_a = sga; sig0_b = sgb;
                const float h2a = __shfl_down_sync(0xffffffffu, h[0], 1), h3a = __shfl_down_sync(0xffffffffu, h[1], 1);
                const float h2b = __shfl_down_sync(0xffffffffu, h[2], 1), h3b = __shfl_down_sync(0xffffffffu, h[3], 1);
                alb_a[0] = round_h(1.f / (1.f + __expf(-round_h(h[1])))); alb_a[1] = round_h(1.f / (1.f + __expf(-round_h(h2a)))); alb_a[2] = round_h(1.f / (1.f + __expf(-round_h(h3a))));
                alb_b[0] = round_h(1.f / (1.f + __expf(-round_h(h[3])))); alb_b[1] = round_h(1.f / (1.f + __expf(-round_h(h2b)))); alb_b[2] = round_h(1.f / (1.f + __expf(-round_h(h3b))));
            } else {
                const int axis = (sp - 1) >> 1;
                const bool plus = (sp & 1) != 0;                 // sp = 1, 3, 5 are the +eps points
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    if (d == axis) { nda[d] = plus ? sga : nda[d] - sga; ndb[d] = plus ? sgb : ndb[d] - sgb; }
                }
            }
        }
        if (t != 0) continue;
        // --- epilogue on the 8 lanes that hold the logits: two samples each
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint32_t si = r ? sb : sa;
            if (!(r ? inb : ina)) continue;
            const float* nd = r ? ndb : nda;
            const float* al = r ? alb_b : alb_a;
            sigmas[si] = r ? sig0_b : sig0_a;
            if (aux) {      // the rest of the stash: unused stencil slots (albedo mode) and the albedo
                float* ax = aux + (size_t)si * kAuxStride;
#pragma unroll
                for (int q = NP; q < 7; q++) ax[q] = 0.f;
                ax[7] = al[0]; ax[8] = al[1]; ax[9] = al[2];
            }
            float col[3] = {al[0], al[1], al[2]};
            if (SHADING != kAlbedo) {
                float n[3];
                n[0] = -(0.5f * nd[0] / kFdEps);
                n[1] = -(0.5f * nd[1] / kFdEps);
                n[2] = -(0.5f * nd[2] / kFdEps);
                const float inv = 1.f / sqrtf(fmaxf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2], 1e-20f));   // safe_normalize (nerf/utils.py:110)
                n[0] = nan_to_num(n[0] * inv); n[1] = nan_to_num(n[1] * inv); n[2] = nan_to_num(n[2] * inv);
                const float* l = light_d + (light_per_sample ? (size_t)si * 3 : 0);
                const float lam = ratio + (1.f - ratio) * fmaxf(n[0] * l[0] + n[1] * l[1] + n[2] * l[2], 0.f);
                if (SHADING == kTextureless) { col[0] = col[1] = col[2] = lam; }
                else if (SHADING == kNormal) { col[0] = (n[0] + 1.f) * 0.5f; col[1] = (n[1] + 1.f) * 0.5f; col[2] = (n[2] + 1.f) * 0.5f; }
                else { col[0] *= lam; col[1] *= lam; col[2] *= lam; }
                if (normals) { normals[(size_t)si * 3] = n[0]; normals[(size_t)si * 3 + 1] = n[1]; normals[(size_t)si * 3 + 2] = n[2]; }
            }
            if (colors) { colors[(size_t)si * 3] = col[0]; colors[(size_t)si * 3 + 1] = col[1]; colors[(size_t)si * 3 + 2] = col[2]; }
        }
    }
}

}  // namespace

// Fused NeRFNetwork.forward / .density (nerf/network_grid.py:104-142) for the hashgrid(L16,C2,smoothstep) + MLP(32-64-64-4)
// backbone.  xyzs [M,3] fp32 in [-bound,bound]; table: fp16 [n_entries,2]; w*/b*: fp32 nn.Linear parameters of sigma_net.
// shading: 0 albedo, 1 lambertian, 2 textureless, 3 normal.  light_d: [3] (light_per_sample=0) or [M,3].
// m_dev (optional): device int32 holding the live sample count (<= M); lets a caller that never learned M on the host
// launch with M = capacity.  colors / normals may be NULL (density-only query).  aux (optional, [M,10] fp32) receives the
// per-sample stash sdf_field_backward needs.  feat (optional, sdf_field_feat_bytes(M, shading) bytes, 16-byte aligned) receives the
// interpolated hash-grid features of every stencil point so that the backward does not have to gather them again.
SDF_API int sdf_field_forward(const float* xyzs, uint32_t M, const int* m_dev, const void* table_fp16, const int* offsets,
                              uint32_t n_levels, uint32_t n_levels_active, float per_level_scale_log2, uint32_t base_resolution,
                              int interp_smoothstep, const float* w1, const float* b1, const float* w2, const float* b2,
                              const float* w3, const float* b3, float bound, float blob_density, float blob_radius,
                              int shading, const float* light_d, int light_per_sample, float ambient_ratio,
                              float* sigmas, float* colors, float* normals, float* aux, void* feat, void* stream) {
    if (M == 0) return SDF_OK;
    SDF_CHECK_ARG(xyzs && table_fp16 && offsets && w1 && b1 && w2 && b2 && w3 && b3 && sigmas, "field_forward: null pointer");
    SDF_CHECK_ARG(n_levels == (uint32_t)kLevels, "field_forward: this build fuses the 16-level / 2-feature grid of the -O backbone");
    SDF_CHECK_ARG(n_levels_active >= 1 && n_levels_active <= n_levels, "field_forward: bad active level count");
    SDF_CHECK_ARG(shading >= 0 && shading <= 3, "field_forward: shading must be 0..3");
    SDF_CHECK_ARG(shading == 0 || light_d, "field_forward: light_d required for shaded modes");
    cudaStream_t st = (cudaStream_t)stream;
    LevelParams* lp;
    int rc = sdf_get_level_params(offsets, n_levels, per_level_scale_log2, base_resolution, st, &lp);
    if (rc) return rc;
    FieldParams p;
    p.table = reinterpret_cast<const __half2*>(table_fp16);
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3;
    p.lp = lp; p.bound = bound; p.n_levels_active = n_levels_active;
    p.blob_density = blob_density; p.blob_radius = blob_radius; p.interp_smoothstep = interp_smoothstep;
    const size_t smem = sizeof(WeightsSmem);
    static_assert(sizeof(WeightsSmem) <= 48 * 1024, "forward weights must fit the default dynamic smem window");
    const uint32_t groups = (M + 15) / 16;
    // resident CTAs per SM: 2 (128 registers; measured 1.18 ms at 432k shaded samples) or 3 (80 registers; 1.34 ms); SDF_FIELD_OCC=3 selects the latter
    static const int occ = [] { const char* e = getenv("SDF_FIELD_OCC"); return (e && e[0] == '3') ? 3 : 2; }();
    const uint32_t blocks = min((uint32_t)(sdf_num_sms() * occ), (groups + 7) / 8);
#define LAUNCH(SH)                                                                                                   \
    do {                                                                                                             \
        if (occ == 2) k_field_forward<SH, 2><<<blocks, 256, smem, st>>>(p, xyzs, light_d, light_per_sample, ambient_ratio, M, m_dev, sigmas, colors, normals, aux, (uint4*)feat); \
        else k_field_forward<SH, 3><<<blocks, 256, smem, st>>>(p, xyzs, light_d, light_per_sample, ambient_ratio, M, m_dev, sigmas, colors, normals, aux, (uint4*)feat); \
    } while (0)
    switch (shading) {
        case 0: LAUNCH(kAlbedo); break;
        case 1: LAUNCH(kLambertian); break;
        case 2: LAUNCH(kTextureless); break;
        default: LAUNCH(kNormal); break;
    }
#undef LAUNCH
    SDF_CHECK_LAUNCH("field_forward");
    return SDF_OK;
}

// bytes of the optional feature stash: ceil(M / 16) groups x (1 | 7) stencil points x 1 KB
SDF_API long long sdf_field_feat_bytes(uint32_t M, int shading) {
    return (long long)((M + 15) / 16) * (shading == 0 ? 1 : 7) * 1024;
}
