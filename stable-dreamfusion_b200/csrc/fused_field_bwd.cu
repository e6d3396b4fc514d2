// fused_field_bwd.cu — backward of the fused radiance field (fused_field.cu), sm_100a.
//
// Replaces, per training step, the reference's autograd walk through 7 x (3 cuBLAS dgrad + 3 wgrad
// GEMMs + ReLU/exp/sigmoid backward kernels + kernel_grid_backward with fp16 atomics + a 12.2 M-entry
// zero-fill and cast-back), nerf/network_grid.py:68-130 and gridencoder/grid.py:72-96.
//
// One CTA (16 warps) owns 256 point-evals per round: each warp re-gathers the hash-grid features of its
// 16 rows (recompute instead of saving 3 M x 160 activations; aligned x-neighbour corner pairs come in one
// 8-byte load, the centre cell's coarse-level corners are re-used by the +-eps stencil points), re-runs the
// MLP keeping the two hidden activations in registers, back-propagates through the MLP on tensor cores
// (the forward-orientation weights in shared memory are read through ldmatrix.trans: no transposed copies),
// scatters d(enc) into the fp32 table gradient with float2 reductions (one float4 reduction for an aligned
// x-neighbour pair), and stages activations / deltas transposed (movmatrix) in shared memory so the CTA can
// form the weight gradients as [features x 256 rows] x [256 rows x features] tensor-core products (64 tile
// pairs, 4 per warp) whose accumulators live in registers for the whole kernel (flushed once with atomics).
// Bias gradients ride along as an extra all-ones activation row (layer 3's is summed per lane instead).
// Warps whose 16 samples carry no upstream gradient skip everything but the barriers.
//
// Algorithmic bytes: 1 052 B per point-eval (12 B xyz + 16 B upstream + 128 float2/half2
// RMWs counted 8 B each).  Roofline: HBM (L2 atomics in practice).
#include "field_common.cuh"
#include <cstdlib>

using namespace field;

namespace {

constexpr float kFdEps = 1e-2f;
constexpr int kAuxStride = 10;
enum Shading { kAlbedo = 0, kLambertian = 1, kTextureless = 2, kNormal = 3 };

// WARPS = 16: one CTA per SM, 256 point-evals per round.  WARPS = 8: two CTAs per SM (~105 KB each), 128 per round — the
// gather / scatter (LSU-bound) phase of one CTA overlaps the weight-gradient (tensor-bound) phase and the barriers of the other.
template <int WARPS>
struct BwdSmemT {
    static constexpr int kRows = WARPS * 16;        // rows per CTA round
    static constexpr int kTStride = kRows + 8;      // halfs; row stride of the transposed staging buffers
    WeightsSmem w;                                  // forward-orientation weights; the data-gradient products read them through ldmatrix.trans
    // staging for the weight gradients, all [feature][row]
    __half enct[kEncDim + 8][kTStride];     // + ones row (bias) + zero rows up to a full n-tile
    __half a1t[kHidden + 8][kTStride];
    __half a2t[kHidden + 8][kTStride];
    __half dh1t[kHidden][kTStride];
    __half dh2t[kHidden][kTStride];
    __half dh3t[16][kTStride];              // 4 logits, zero padded to one m-tile
};

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

// Per-sample upstream -> gradients wrt the 7 stencil densities and the 3 albedo logits.
template <int SHADING>
__device__ __forceinline__ void sample_grads(float gsig[7], float glogit[3], const float* __restrict__ ax,
                                             float g_sigma, const float gcol[3], const float gnrm[3],
                                             const float* __restrict__ l, float ratio) {
    constexpr int NP = (SHADING == kAlbedo) ? 1 : 7;
    const float alb[3] = {ax[7], ax[8], ax[9]};
    float galb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 7; q++) gsig[q] = 0.f;
    gsig[0] = g_sigma;
    if (SHADING == kAlbedo) {
        galb[0] = gcol[0]; galb[1] = gcol[1]; galb[2] = gcol[2];
    } else {
        float nr[3];
        nr[0] = -(0.5f * (ax[1] - ax[2]) / kFdEps);
        nr[1] = -(0.5f * (ax[3] - ax[4]) / kFdEps);
        nr[2] = -(0.5f * (ax[5] - ax[6]) / kFdEps);
        const float q2 = nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2];
        const float inv = 1.f / sqrtf(fmaxf(q2, 1e-20f));
        const float n[3] = {nr[0] * inv, nr[1] * inv, nr[2] * inv};
        const float ndl = n[0] * l[0] + n[1] * l[1] + n[2] * l[2];
        const float lam = ratio + (1.f - ratio) * fmaxf(ndl, 0.f);
        float gn[3] = {gnrm[0], gnrm[1], gnrm[2]};
        float glam = 0.f;
        if (SHADING == kLambertian) {
            glam = gcol[0] * alb[0] + gcol[1] * alb[1] + gcol[2] * alb[2];
            galb[0] = gcol[0] * lam; galb[1] = gcol[1] * lam; galb[2] = gcol[2] * lam;
        } else if (SHADING == kTextureless) {
            glam = gcol[0] + gcol[1] + gcol[2];
        } else {
            gn[0] += 0.5f * gcol[0]; gn[1] += 0.5f * gcol[1]; gn[2] += 0.5f * gcol[2];
        }
        if (ndl > 0.f) {
            const float k = glam * (1.f - ratio);
            gn[0] += k * l[0]; gn[1] += k * l[1]; gn[2] += k * l[2];
        }
        float gnr[3];
        if (q2 > 1e-20f) {
            const float d = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
            gnr[0] = (gn[0] - n[0] * d) * inv; gnr[1] = (gn[1] - n[1] * d) * inv; gnr[2] = (gn[2] - n[2] * d) * inv;
        } else {
            gnr[0] = gn[0] * inv; gnr[1] = gn[1] * inv; gnr[2] = gn[2] * inv;
        }
        const float k = 0.5f / kFdEps;
        gsig[1] = -k * gnr[0]; gsig[2] = k * gnr[0];
        gsig[3] = -k * gnr[1]; gsig[4] = k * gnr[1];
        gsig[5] = -k * gnr[2]; gsig[6] = k * gnr[2];
    }
    // trunc_exp backward (activation.py:14-18): g * exp(min(z, 15)); sigma_p = exp(z_p) is in the stash
#pragma unroll
    for (int q = 0; q < NP; q++) gsig[q] *= fminf(ax[q], 3269017.3724721107f);
#pragma unroll
    for (int c = 0; c < 3; c++) glogit[c] = galb[c] * alb[c] * (1.f - alb[c]);
}

// Scatter d(enc) of one (row, level) into the fp32 table gradient (same corner geometry as the forward gather).
// The kernel is bound by the LSU's scattered-reduction rate (~1.3 cycles per active lane per SM), so the two x-neighbours of
// a corner pair share ONE 16-byte reduction whenever they sit in the same aligned pair of table entries: always for hashed
// levels when x0 is even (x1 = x0 ^ 1 flips only bit 0 of the hash) and for dense levels when the linear index is even.
__device__ __forceinline__ void scatter_level(float* __restrict__ grad_table, const LevelSmem& lv, float x, float y, float z,
                                              bool smooth, float g0, float g1) {
    Corners c;
    level_corners(c, lv, x, y, z, smooth);
    float2* t = reinterpret_cast<float2*>(grad_table) + lv.offset;       // level offsets are multiples of 8 entries: pairs stay 16-byte aligned
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i0 = c.idx[2 * j], i1 = c.idx[2 * j + 1];
        const float w0 = c.w[2 * j], w1 = c.w[2 * j + 1];
        if ((i0 ^ i1) == 1u) {
            const bool lo = (i0 & 1u) == 0u;                                 // which corner owns the even slot
            const float wa = lo ? w0 : w1, wb = lo ? w1 : w0;
            atomicAdd(reinterpret_cast<float4*>(t + (i0 & ~1u)), make_float4(wa * g0, wa * g1, wb * g0, wb * g1));
        } else {
            atomicAdd(t + i0, make_float2(w0 * g0, w0 * g1));
            atomicAdd(t + i1, make_float2(w1 * g0, w1 * g1));
        }
    }
}

// scatter_level for the COARSE levels, executed by the whole warp: consecutive samples of a ray sit in the same cell of a coarse level, so
// the 8 lanes that handle one level (lanes t, t+4, ..., t+28: rows g = 0..7) would fire 8 x 8 reductions at the SAME 8 table entries — and
// same-address reductions serialise in L2 (measured: 2.46 ms on marched samples vs 1.49 ms on uniformly random points of the same count).
// The 16 values of a lane (8 corners x 2 features) are pre-summed across aligned pairs / quads / the octet of rows that share a cell with a
// halving butterfly (8 + 4 + 2 shuffles); a block that shares its cell then issues ONE reduction per corner instead of one per row.
__device__ __forceinline__ uint32_t pick8(const uint32_t (&a)[8], int k) {
    uint32_t r = a[0];
#pragma unroll
    for (int i = 1; i < 8; i++) r = (k == i) ? a[i] : r;
    return r;
}
__device__ __forceinline__ void scatter_level_agg(float* __restrict__ grad_table, const LevelSmem& lv, float x, float y, float z,
                                                  bool smooth, float g0, float g1, int lane) {
    Corners c;
    const uint32_t key = level_corners(c, lv, x, y, z, smooth);
    float v[16];
#pragma unroll
    for (int k = 0; k < 8; k++) { v[2 * k] = c.w[k] * g0; v[2 * k + 1] = c.w[k] * g1; }
    // which aligned blocks of rows share this lane's cell (rows g = lane >> 2; partner rows at lane ^ 4, ^ 8, ^ 16)
    const uint32_t full = 0xffffffffu;
    const uint32_t eq4 = __ballot_sync(full, __shfl_xor_sync(full, key, 4) == key);
    const uint32_t eq8 = __ballot_sync(full, __shfl_xor_sync(full, key, 8) == key);
    const uint32_t eq16 = __ballot_sync(full, __shfl_xor_sync(full, key, 16) == key);
    const bool pair_same = (eq4 >> lane) & 1u;
    const bool quad_same = pair_same && ((eq8 >> lane) & 1u) && ((eq4 >> (lane ^ 8)) & 1u);
    const uint32_t quad_mask = __ballot_sync(full, quad_same);
    const bool oct_same = quad_same && ((eq16 >> lane) & 1u) && ((quad_mask >> (lane ^ 16)) & 1u);
    // halving butterfly: after step A a lane holds 4 corners summed over its pair, after B 2 corners over its quad, after C 1 corner over the octet
    float a8[8], b4[4], c2[2];
    const bool up4 = lane & 4, up8 = lane & 8, up16 = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float send = up4 ? v[i] : v[i + 8], keep = up4 ? v[i + 8] : v[i];
        a8[i] = keep + __shfl_xor_sync(full, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float send = up8 ? a8[i] : a8[i + 4], keep = up8 ? a8[i + 4] : a8[i];
        b4[i] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = up16 ? b4[i] : b4[i + 2], keep = up16 ? b4[i + 2] : b4[i];
        c2[i] = keep + __shfl_xor_sync(full, send, 16);
    }
    float2* t = reinterpret_cast<float2*>(grad_table) + lv.offset;
    const int k4 = up4 ? 4 : 0, k2 = up8 ? 2 : 0, k1 = up16 ? 1 : 0;
    if (oct_same) {
        if (c2[0] != 0.f || c2[1] != 0.f) atomicAdd(t + pick8(c.idx, k4 + k2 + k1), make_float2(c2[0], c2[1]));
    } else if (quad_same) {
#pragma unroll
        for (int j = 0; j < 2; j++)
            if (b4[2 * j] != 0.f || b4[2 * j + 1] != 0.f) atomicAdd(t + pick8(c.idx, k4 + k2 + j), make_float2(b4[2 * j], b4[2 * j + 1]));
    } else if (pair_same) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (a8[2 * j] != 0.f || a8[2 * j + 1] != 0.f) atomicAdd(t + pick8(c.idx, k4 + j), make_float2(a8[2 * j], a8[2 * j + 1]));
    } else if (g0 != 0.f || g1 != 0.f) {
#pragma unroll
        for (int j = 0; j < 4; j++) {        // this row alone: aligned x-neighbour pairs still share one 16-byte reduction
            const uint32_t i0 = c.idx[2 * j], i1 = c.idx[2 * j + 1];
            if ((i0 ^ i1) == 1u) {
                const bool lo = (i0 & 1u) == 0u;
                const float w0 = lo ? c.w[2 * j] : c.w[2 * j + 1], w1 = lo ? c.w[2 * j + 1] : c.w[2 * j];
                atomicAdd(reinterpret_cast<float4*>(t + (i0 & ~1u)), make_float4(w0 * g0, w0 * g1, w1 * g0, w1 * g1));
            } else {
                atomicAdd(t + i0, make_float2(v[4 * j], v[4 * j + 1]));
                atomicAdd(t + i1, make_float2(v[4 * j + 2], v[4 * j + 3]));
            }
        }
    }
}

// Staging for the weight-gradient products is [feature][row].  A register in A-fragment block layout (lane (g, t) holds
// X[row0 + g][f0 + 2t .. 2t+1]) is transposed across the warp with movmatrix, after which lane (g, t) holds
// X[row0 + 2t .. 2t+1][f0 + g]: one conflict-free 32-bit store per register instead of two 16-bit ones.
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ void stage_block(__half* buf, int stride, int f0, int row0, uint32_t reg, int g, int t) {
    *reinterpret_cast<uint32_t*>(buf + (f0 + g) * stride + row0 + 2 * t) = movm_trans(reg);
}
// all four blocks of one A-fragment k-tile (16 rows x 16 features)
__device__ __forceinline__ void stage_frag(__half* buf, int stride, int f0, int row0, const uint32_t a[4], int g, int t) {
    stage_block(buf, stride, f0, row0, a[0], g, t);
    stage_block(buf, stride, f0, row0 + 8, a[1], g, t);
    stage_block(buf, stride, f0 + 8, row0, a[2], g, t);
    stage_block(buf, stride, f0 + 8, row0 + 8, a[3], g, t);
}
__device__ __forceinline__ void clear_rows(__half* buf, int stride, int n_feat, int row0, int lane) {
    // 16 staged rows (32 bytes) of every feature: 8 x 32-bit per feature
    for (int i = lane; i < n_feat * 8; i += 32) *reinterpret_cast<uint32_t*>(buf + (i >> 3) * stride + row0 + (i & 7) * 2) = 0u;
}
// B fragments of one k-step (k0 .. k0+15) for the TWO n-tiles at n0 and n0 + 8 of an operand stored [k][n] (n contiguous):
// r[0], r[1] = (b0, b1) of n-tile n0; r[2], r[3] = of n-tile n0 + 8.  This is how the data-gradient products read the
// forward-orientation weights W[out = k][in = n] without a transposed copy.
__device__ __forceinline__ void ldsm_bt2(uint32_t r[4], const __half* base, int stride, int k0, int n0, int lane) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(base + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * stride + n0 + (lane >> 4) * 8);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// b0 of FOUR consecutive n-tiles (n0, n0+8, n0+16, n0+24) for the k rows k0 .. k0+7 of a [k][n] operand.
__device__ __forceinline__ void ldsm_bt_k8(uint32_t r[4], const __half* base, int stride, int k0, int n0, int lane) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(base + (k0 + (lane & 7)) * stride + n0 + (lane >> 3) * 8);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <int SHADING, bool PREFETCH, int WARPS, bool AGG>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 16 ? 1 : 2)
k_field_backward(FieldParams p, const float* __restrict__ xyzs, const float* __restrict__ light_d, int light_per_sample,
                 float ratio, uint32_t M_cap, const int* __restrict__ m_dev, const float* __restrict__ aux,
                 const float* __restrict__ g_sigmas, const float* __restrict__ g_colors, const float* __restrict__ g_normals,
                 float* __restrict__ grad_table, float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ gw2,
                 float* __restrict__ gb2, float* __restrict__ gw3, float* __restrict__ gb3, const uint4* __restrict__ feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = BwdSmemT<WARPS>;
    constexpr int kWarps = WARPS, kRows = Smem::kRows, kTStride = Smem::kTStride;
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    load_weights(s.w, p);
    // constant rows of the staging buffers
    for (int i = threadIdx.x; i < 8 * kTStride; i += blockDim.x) {
        const __half v = __float2half_rn((i / kTStride) == 0 ? 1.f : 0.f);
        (&s.enct[kEncDim][0])[i] = v;
        (&s.a1t[kHidden][0])[i] = v;
        (&s.a2t[kHidden][0])[i] = v;
    }
    for (int i = threadIdx.x; i < 16 * kTStride; i += blockDim.x) (&s.dh3t[0][0])[i] = __float2half_rn(0.f);
    // rows of a warp that skips a round keep whatever was staged before: make sure that is never NaN garbage
    for (int i = threadIdx.x; i < kEncDim * kTStride; i += blockDim.x) (&s.enct[0][0])[i] = __float2half_rn(0.f);
    for (int i = threadIdx.x; i < kHidden * kTStride; i += blockDim.x) {
        (&s.a1t[0][0])[i] = __float2half_rn(0.f); (&s.a2t[0][0])[i] = __float2half_rn(0.f);
        (&s.dh1t[0][0])[i] = __float2half_rn(0.f); (&s.dh2t[0][0])[i] = __float2half_rn(0.f);
    }
    __syncthreads();

    const uint32_t M = m_dev ? min((uint32_t)max(*m_dev, 0), M_cap) : M_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NP = (SHADING == kAlbedo) ? 1 : 7;
    const bool smooth = p.interp_smoothstep != 0;
    const uint32_t n_groups = (M + 15) / 16;
    const uint32_t n_super = (n_groups + kWarps - 1) / kWarps;     // CTA rounds of 16 groups

    // ---- weight-gradient tile assignment: 64 (m-tile, n-tile) pairs over the CTA's warps
    //   q in [0,36): layer 2  (4 m-tiles x 9 n-tiles; n-tile 8 = ones row -> bias)
    //   q in [36,56): layer 1 (4 x 5)           q in [56,64): layer 3 (1 x 8; its bias gradient is summed in registers below,
    //   which leaves exactly 4 (16 warps) or 8 (8 warps) pairs per warp)
    constexpr int kPairs = 64, kMaxPerWarp = kPairs / kWarps;
    float gb3_acc[2] = {0.f, 0.f};      // lanes t == 0: logits 0, 1; t == 1: logits 2, 3
    float wacc[kMaxPerWarp][4];
#pragma unroll
    for (int j = 0; j < kMaxPerWarp; j++) wacc[j][0] = wacc[j][1] = wacc[j][2] = wacc[j][3] = 0.f;

    for (uint32_t sup = blockIdx.x; sup < n_super; sup += gridDim.x) {
        const uint32_t grp = sup * kWarps + warp;
        const uint32_t sa = grp * 16 + g, sb = sa + 8;
        const bool ina = sa < M, inb = sb < M;
        float xa[3] = {0.f, 0.f, 0.f}, xb[3] = {0.f, 0.f, 0.f};
        if (ina) { xa[0] = xyzs[(size_t)sa * 3]; xa[1] = xyzs[(size_t)sa * 3 + 1]; xa[2] = xyzs[(size_t)sa * 3 + 2]; }
        if (inb) { xb[0] = xyzs[(size_t)sb * 3]; xb[1] = xyzs[(size_t)sb * 3 + 1]; xb[2] = xyzs[(size_t)sb * 3 + 2]; }

        // per-sample gradient wrt the stencil densities / albedo logits (only the t==0 lanes need them)
        float gsa[7], gsb[7], gla[3], glb[3];
#pragma unroll
        for (int q = 0; q < 7; q++) gsa[q] = gsb[q] = 0.f;
        gla[0] = gla[1] = gla[2] = glb[0] = glb[1] = glb[2] = 0.f;
        if (t == 0) {
            const float zero3[3] = {0.f, 0.f, 0.f};
            if (ina) {
                const float gc[3] = {g_colors ? g_colors[(size_t)sa * 3] : 0.f, g_colors ? g_colors[(size_t)sa * 3 + 1] : 0.f, g_colors ? g_colors[(size_t)sa * 3 + 2] : 0.f};
                const float gn[3] = {g_normals ? g_normals[(size_t)sa * 3] : 0.f, g_normals ? g_normals[(size_t)sa * 3 + 1] : 0.f, g_normals ? g_normals[(size_t)sa * 3 + 2] : 0.f};
                sample_grads<SHADING>(gsa, gla, aux + (size_t)sa * kAuxStride, g_sigmas ? g_sigmas[sa] : 0.f, gc, gn,
                                      light_d ? light_d + (light_per_sample ? (size_t)sa * 3 : 0) : zero3, ratio);
            }
            if (inb) {
                const float gc[3] = {g_colors ? g_colors[(size_t)sb * 3] : 0.f, g_colors ? g_colors[(size_t)sb * 3 + 1] : 0.f, g_colors ? g_colors[(size_t)sb * 3 + 2] : 0.f};
                const float gn[3] = {g_normals ? g_normals[(size_t)sb * 3] : 0.f, g_normals ? g_normals[(size_t)sb * 3 + 1] : 0.f, g_normals ? g_normals[(size_t)sb * 3 + 2] : 0.f};
                sample_grads<SHADING>(gsb, glb, aux + (size_t)sb * kAuxStride, g_sigmas ? g_sigmas[sb] : 0.f, gc, gn,
                                      light_d ? light_d + (light_per_sample ? (size_t)sb * 3 : 0) : zero3, ratio);
            }
        }

        // a warp whose 16 samples carry no upstream gradient (early-terminated ray tails, rays that missed) only has to
        // clear its delta rows; it still meets the CTA barriers because the weight-gradient products span all 256 rows
        bool mine = false;
#pragma unroll
        for (int q = 0; q < 7; q++) mine |= (gsa[q] != 0.f) | (gsb[q] != 0.f);
        mine |= (gla[0] != 0.f) | (gla[1] != 0.f) | (gla[2] != 0.f) | (glb[0] != 0.f) | (glb[1] != 0.f) | (glb[2] != 0.f);
        const bool warp_active = __any_sync(0xffffffffu, mine);
        bool rows_cleared = false;

        // software pipeline over the stencil points: the 32 fine-level gathers of point sp+1 are issued before the CTA barrier and the
        // weight-gradient products of point sp, so their latency is hidden behind tensor-core work instead of being exposed
        __half2 raw[2][16];
        CornerCache cache;       // centre-cell corner values of the two coarse level slots, reused by the +-eps stencil points
        if (PREFETCH && !feat && warp_active) {
            float pa[3], pb[3], ua[3], ub[3];
            stencil_point(pa, xa, 0, p.bound);
            stencil_point(pb, xb, 0, p.bound);
            to_unit(ua, pa, p.bound); to_unit(ub, pb, p.bound);
            gather_issue(raw, s.w, p, lane, ua, ub);
        }

#pragma unroll 1
        for (int sp = 0; sp < NP; sp++) {
            if (!warp_active) {
                if (!rows_cleared) {
                    clear_rows(&s.dh3t[0][0], kTStride, 8, warp * 16, lane);
                    clear_rows(&s.dh2t[0][0], kTStride, kHidden, warp * 16, lane);
                    clear_rows(&s.dh1t[0][0], kTStride, kHidden, warp * 16, lane);
                    rows_cleared = true;
                }
            } else {
            // ---- delta at the logits as an A fragment: k = logit index (0..3), zero beyond
            float d0a = 0.f, d1a = 0.f, d0b = 0.f, d1b = 0.f;     // (k=2t, 2t+1) for rows g and g+8
            {
                float za = 0.f, zb = 0.f;
#pragma unroll
                for (int q = 0; q < NP; q++) if (q == sp) { za = gsa[q]; zb = gsb[q]; }
                const float l1a = sp == 0 ? gla[0] : 0.f, l2a = sp == 0 ? gla[1] : 0.f, l3a = sp == 0 ? gla[2] : 0.f;
                const float l1b = sp == 0 ? glb[0] : 0.f, l2b = sp == 0 ? glb[1] : 0.f, l3b = sp == 0 ? glb[2] : 0.f;
                // lanes t==0 own the values; lane t==1 needs logits 2,3
                const float r2a = __shfl_up_sync(0xffffffffu, l2a, 1), r3a = __shfl_up_sync(0xffffffffu, l3a, 1);
                const float r2b = __shfl_up_sync(0xffffffffu, l2b, 1), r3b = __shfl_up_sync(0xffffffffu, l3b, 1);
                if (t == 0) { d0a = za; d1a = l1a; d0b = zb; d1b = l1b; }
                else if (t == 1) { d0a = r2a; d1a = r3a; d0b = r2b; d1b = r3b; }
            }
            // ---- forward recompute
            float pa[3], pb[3], ua[3], ub[3];
            stencil_point(pa, xa, sp, p.bound);
            stencil_point(pb, xb, sp, p.bound);
            const bool va = to_unit(ua, pa, p.bound) && ina, vb = to_unit(ub, pb, p.bound) && inb;
            uint32_t a0[2][4], a1[4][4], a2[4][4];
            if (feat) {
                // features stashed by the forward (coalesced 2 x 16 B per lane) instead of 64 scattered gathers per lane
                const uint4* f = feat + ((size_t)grp * NP + sp) * 64 + lane * 2;
                const uint4 f0 = __ldg(f), f1 = __ldg(f + 1);
                a0[0][0] = f0.x; a0[0][1] = f0.y; a0[0][2] = f0.z; a0[0][3] = f0.w;
                a0[1][0] = f1.x; a0[1][1] = f1.y; a0[1][2] = f1.z; a0[1][3] = f1.w;
            } else if (PREFETCH) gather_finish(a0, raw, s.w, p, lane, ua, va, ub, vb);
            else if (NP > 1) encode_rows_cached(a0, s.w, p, lane, ua, va, ub, vb, cache, sp == 0);
            else encode_rows(a0, s.w, p, lane, ua, va, ub, vb);
            float hdummy[4];
            mlp_forward<true>(hdummy, a0, s.w, lane, a1, a2);

            const int row0 = warp * 16;
            gb3_acc[0] += d0a + d0b; gb3_acc[1] += d1a + d1b;
            // stage dh3^T (logit deltas), the layer inputs a2^T, a1^T and enc^T
            uint32_t d3frag[4] = {pack_half2(d0a, d1a), pack_half2(d0b, d1b), 0u, 0u};
            stage_block(&s.dh3t[0][0], kTStride, 0, row0, d3frag[0], g, t);
            stage_block(&s.dh3t[0][0], kTStride, 0, row0 + 8, d3frag[1], g, t);
#pragma unroll
            for (int kt = 0; kt < 4; kt++) {
                stage_frag(&s.a2t[0][0], kTStride, kt * 16, row0, a2[kt], g, t);
                stage_frag(&s.a1t[0][0], kTStride, kt * 16, row0, a1[kt], g, t);
            }
#pragma unroll
            for (int kt = 0; kt < 2; kt++) stage_frag(&s.enct[0][0], kTStride, kt * 16, row0, a0[kt], g, t);

            // ---- dh2 = (dh3 . W3) * relu'(a2)    A = dh3 [16 x 16(k: 4 valid)], B[k = logit][n = hidden] = W3[k][n] (rows 4..7 zero)
            uint32_t dh2[4][4];
            uint32_t w3b[8];       // b0 of the 8 n-tiles (k = logit 0..7; b1 = 0)
            ldsm_bt_k8(w3b, &s.w.w3[0][0], kW2Stride, 0, 0, lane);
            ldsm_bt_k8(w3b + 4, &s.w.w3[0][0], kW2Stride, 0, 32, lane);
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                mma16816(c, d3frag, w3b[nt], 0u);
                const int kt2 = nt >> 1, hi = (nt & 1) * 2;
                const float2 ma = unpack_half2(a2[kt2][hi + 0]), mb = unpack_half2(a2[kt2][hi + 1]);
                c[0] = ma.x > 0.f ? c[0] : 0.f; c[1] = ma.y > 0.f ? c[1] : 0.f;
                c[2] = mb.x > 0.f ? c[2] : 0.f; c[3] = mb.y > 0.f ? c[3] : 0.f;
                dh2[kt2][hi + 0] = pack_half2(c[0], c[1]);
                dh2[kt2][hi + 1] = pack_half2(c[2], c[3]);
                stage_block(&s.dh2t[0][0], kTStride, nt * 8, row0, dh2[kt2][hi + 0], g, t);
                stage_block(&s.dh2t[0][0], kTStride, nt * 8, row0 + 8, dh2[kt2][hi + 1], g, t);
            }
            // ---- dh1 = (dh2 . W2) * relu'(a1)    B[k = out][n = in] = W2[k][n]: forward weights through ldmatrix.trans,
            //      two n-tiles per load
            uint32_t dh1[4][4];
#pragma unroll
            for (int np = 0; np < 4; np++) {
                float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int kt = 0; kt < 4; kt++) {
                    uint32_t wb[4];
                    ldsm_bt2(wb, &s.w.w2[0][0], kW2Stride, kt * 16, np * 16, lane);
                    mma16816(c[0], dh2[kt], wb[0], wb[1]);
                    mma16816(c[1], dh2[kt], wb[2], wb[3]);
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int nt = np * 2 + e;                      // kt2 = np, hi = 2 e
                    const float2 ma = unpack_half2(a1[np][2 * e + 0]), mb = unpack_half2(a1[np][2 * e + 1]);
                    c[e][0] = ma.x > 0.f ? c[e][0] : 0.f; c[e][1] = ma.y > 0.f ? c[e][1] : 0.f;
                    c[e][2] = mb.x > 0.f ? c[e][2] : 0.f; c[e][3] = mb.y > 0.f ? c[e][3] : 0.f;
                    dh1[np][2 * e + 0] = pack_half2(c[e][0], c[e][1]);
                    dh1[np][2 * e + 1] = pack_half2(c[e][2], c[e][3]);
                    stage_block(&s.dh1t[0][0], kTStride, nt * 8, row0, dh1[np][2 * e + 0], g, t);
                    stage_block(&s.dh1t[0][0], kTStride, nt * 8, row0 + 8, dh1[np][2 * e + 1], g, t);
                }
            }
            // ---- d(enc) = dh1 . W1 ; n-tile nt covers levels 4nt..4nt+3: lane gets (row g / g+8, level 4nt + t)
#pragma unroll
            for (int np = 0; np < 2; np++) {
                float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int kt = 0; kt < 4; kt++) {
                    uint32_t wb[4];
                    ldsm_bt2(wb, &s.w.w1[0][0], kW1Stride, kt * 16, np * 16, lane);
                    mma16816(c[0], dh1[kt], wb[0], wb[1]);
                    mma16816(c[1], dh1[kt], wb[2], wb[3]);
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t level = (np * 2 + e) * 4 + t;
                    if (AGG && np == 0) {
                        // levels 0..7: warp-wide pre-summation of rows that share a cell (every lane takes part; inactive rows add zeros)
                        const bool act = level < p.n_levels_active;
                        const LevelSmem lv = s.w.lv[act ? level : 0];
                        scatter_level_agg(grad_table, lv, ua[0], ua[1], ua[2], smooth, (act && va) ? c[e][0] : 0.f, (act && va) ? c[e][1] : 0.f, lane);
                        scatter_level_agg(grad_table, lv, ub[0], ub[1], ub[2], smooth, (act && vb) ? c[e][2] : 0.f, (act && vb) ? c[e][3] : 0.f, lane);
                    } else if (level < p.n_levels_active) {
                        const LevelSmem lv = s.w.lv[level];
                        if (va && (c[e][0] != 0.f || c[e][1] != 0.f)) scatter_level(grad_table, lv, ua[0], ua[1], ua[2], smooth, c[e][0], c[e][1]);
                        if (vb && (c[e][2] != 0.f || c[e][3] != 0.f)) scatter_level(grad_table, lv, ub[0], ub[1], ub[2], smooth, c[e][2], c[e][3]);
                    }
                }
            }
            if (PREFETCH && !feat && sp + 1 < NP) {      // next stencil point's gathers fly during the barrier + weight-gradient phase
                float qa[3], qb[3], wa[3], wb[3];
                stencil_point(qa, xa, sp + 1, p.bound);
                stencil_point(qb, xb, sp + 1, p.bound);
                to_unit(wa, qa, p.bound); to_unit(wb, qb, p.bound);
                gather_issue(raw, s.w, p, lane, wa, wb);
            }
            }   // warp_active
            __syncthreads();

            // ---- weight gradients over the staged rows: k (row) step outer, the warp's tile pairs inner, so the kMaxPerWarp
            //      accumulator chains are independent back-to-back MMAs
            {
                const __half* Aj[kMaxPerWarp]; const __half* Bj[kMaxPerWarp];
#pragma unroll
                for (int j = 0; j < kMaxPerWarp; j++) {
                    const int q = warp + j * kWarps;
                    int mt, nt; const __half* A; const __half* B;
                    if (q < 36) { mt = q / 9; nt = q % 9; A = &s.dh2t[0][0]; B = &s.a1t[0][0]; }
                    else if (q < 56) { mt = (q - 36) / 5; nt = (q - 36) % 5; A = &s.dh1t[0][0]; B = &s.enct[0][0]; }
                    else { mt = 0; nt = q - 56; A = &s.dh3t[0][0]; B = &s.a2t[0][0]; }
                    Aj[j] = A + (mt * 16) * kTStride;
                    Bj[j] = B + (nt * 8) * kTStride;
                }
#pragma unroll 2
                for (int kp = 0; kp < kRows / 32; kp++) {          // two k-steps (32 staged rows) per trip: 3 ldmatrix.x4 for 2 MMAs
#pragma unroll
                    for (int j = 0; j < kMaxPerWarp; j++) {
                        uint32_t af0[4], af1[4], bf[4];
                        ldsm_a(af0, Aj[j], kTStride, kp * 32, lane);
                        ldsm_a(af1, Aj[j], kTStride, kp * 32 + 16, lane);
                        ldsm_b2(bf, Bj[j], kTStride, kp * 32, lane);
                        mma16816(wacc[j], af0, bf[0], bf[1]);
                        mma16816(wacc[j], af1, bf[2], bf[3]);
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- layer-3 bias gradient: sum the per-lane partials over the 8 row groups of the warp
#pragma unroll
    for (int e = 0; e < 2; e++) {
        float v = gb3_acc[e];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0 && t < 2 && v != 0.f) atomicAdd(gb3 + 2 * t + e, v);
    }
    // ---- flush the weight-gradient accumulators: wacc[j] = (out = 16mt + g (+8), in = 8nt + 2t (+1))
#pragma unroll
    for (int j = 0; j < kMaxPerWarp; j++) {
        const int q = warp + j * kWarps;
        if (q >= kPairs) continue;
        float* gw; float* gb; int mt, nt, in_dim, out_dim;
        if (q < 36) { mt = q / 9; nt = q % 9; gw = gw2; gb = gb2; in_dim = kHidden; out_dim = kHidden; }
        else if (q < 56) { mt = (q - 36) / 5; nt = (q - 36) % 5; gw = gw1; gb = gb1; in_dim = kEncDim; out_dim = kHidden; }
        else { mt = 0; nt = q - 56; gw = gw3; gb = gb3; in_dim = kHidden; out_dim = kOut; }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int out = mt * 16 + g + ((e & 2) ? 8 : 0);
            const int in = nt * 8 + 2 * t + (e & 1);
            const float v = wacc[j][e];
            if (out < out_dim && v != 0.f) {
                if (in < in_dim) atomicAdd(gw + out * in_dim + in, v);
                else if (in == in_dim) atomicAdd(gb + out, v);
            }
        }
    }
}

}  // namespace

// Backward of sdf_field_forward.  Upstream gradients: g_sigmas [M] (may be NULL), g_colors [M,3] (may be NULL),
// g_normals [M,3] (may be NULL).  aux: the [M,10] stash written by the forward.  All outputs are ACCUMULATED into
// (caller zero-fills): grad_table fp32 [n_entries,2]; gw1 [64,32], gb1 [64], gw2 [64,64], gb2 [64], gw3 [4,64], gb3 [4].
// feat (optional): the feature stash the forward wrote for the SAME (xyzs, M, shading); NULL = re-gather from the table.
SDF_API int sdf_field_backward(const float* xyzs, uint32_t M, const int* m_dev, const void* table_fp16, const int* offsets,
                               uint32_t n_levels, uint32_t n_levels_active, float per_level_scale_log2, uint32_t base_resolution,
                               int interp_smoothstep, const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* w3, const float* b3, float bound, float blob_density, float blob_radius,
                               int shading, const float* light_d, int light_per_sample, float ambient_ratio, const float* aux,
                               const float* g_sigmas, const float* g_colors, const float* g_normals,
                               float* grad_table, float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3,
                               const void* feat, void* stream) {
    if (M == 0) return SDF_OK;
    SDF_CHECK_ARG(xyzs && table_fp16 && offsets && w1 && b1 && w2 && b2 && w3 && b3 && aux, "field_backward: null pointer");
    SDF_CHECK_ARG(grad_table && gw1 && gb1 && gw2 && gb2 && gw3 && gb3, "field_backward: null gradient output");
    SDF_CHECK_ARG(n_levels == (uint32_t)kLevels && n_levels_active >= 1 && n_levels_active <= n_levels, "field_backward: bad level count");
    SDF_CHECK_ARG(shading >= 0 && shading <= 3, "field_backward: shading must be 0..3");
    SDF_CHECK_ARG(shading == 0 || light_d, "field_backward: light_d required for shaded modes");
    cudaStream_t st = (cudaStream_t)stream;
    LevelParams* lp;
    int rc = sdf_get_level_params(offsets, n_levels, per_level_scale_log2, base_resolution, st, &lp);
    if (rc) return rc;
    FieldParams p;
    p.table = reinterpret_cast<const __half2*>(table_fp16);
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3;
    p.lp = lp; p.bound = bound; p.n_levels_active = n_levels_active;
    p.blob_density = blob_density; p.blob_radius = blob_radius; p.interp_smoothstep = interp_smoothstep;
    // Measured at 432 k shaded samples (tools/bench_field.py): 16 warps 3.32 ms, 16 warps + prefetch 3.35 ms, 2 x 8 warps 4.03 ms,
    // 2 x 8 warps + prefetch 4.31 ms.  The default is the first; SDF_FIELD_BWD_WARPS=8 / SDF_FIELD_BWD_PREFETCH=1 select the others.
    static const bool prefetch = [] { const char* e = getenv("SDF_FIELD_BWD_PREFETCH"); return e && e[0] == '1'; }();
    static const int warps = [] { const char* e = getenv("SDF_FIELD_BWD_WARPS"); return (e && atoi(e) == 8) ? 8 : 16; }();
    // warp-wide pre-summation of the coarse levels' reductions (scatter_level_agg); SDF_FIELD_BWD_AGG=0 selects the per-row scatter
    static const bool agg = [] { const char* e = getenv("SDF_FIELD_BWD_AGG"); return !(e && e[0] == '0'); }();
    const uint32_t n_groups = (M + 15) / 16;
#define LAUNCH_V(SH, PF, WP, AG)                                                                                              \
    do {                                                                                                                  \
        const int smem = (int)sizeof(BwdSmemT<WP>);                                                                       \
        const uint32_t n_super = (n_groups + WP - 1) / WP;                                                                \
        const uint32_t blocks = min((uint32_t)(sdf_num_sms() * (WP == 16 ? 1 : 2)), n_super);                                   \
        static bool attr_set[64] = {false};                                                                               \
        int dev = 0; cudaGetDevice(&dev);                                                                                 \
        if (dev < 64 && !attr_set[dev]) {                                                                                 \
            SDF_CHECK_CUDA(cudaFuncSetAttribute(k_field_backward<SH, PF, WP, AG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            attr_set[dev] = true;                                                                                         \
        }                                                                                                                 \
        k_field_backward<SH, PF, WP, AG><<<blocks, WP * 32, smem, st>>>(p, xyzs, light_d, light_per_sample, ambient_ratio, M, m_dev, aux, \
                                                                   g_sigmas, g_colors, g_normals, grad_table, gw1, gb1, gw2, gb2, gw3, gb3, (const uint4*)feat); \
    } while (0)
#define LAUNCH(SH)                                                                           \
    do {                                                                                     \
        if (warps == 16) { if (prefetch) LAUNCH_V(SH, true, 16, false); else if (agg) LAUNCH_V(SH, false, 16, true); else LAUNCH_V(SH, false, 16, false); } \
        else { if (prefetch) LAUNCH_V(SH, true, 8, false); else LAUNCH_V(SH, false, 8, false); }           \
    } while (0)
    switch (shading) {
        case 0: LAUNCH(kAlbedo); break;
        case 1: LAUNCH(kLambertian); break;
        case 2: LAUNCH(kTextureless); break;
        default: LAUNCH(kNormal); break;
    }
#undef LAUNCH
#undef LAUNCH_V
    SDF_CHECK_LAUNCH("field_backward");
    return SDF_OK;
}
