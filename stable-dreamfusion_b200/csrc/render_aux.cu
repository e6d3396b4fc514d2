// render_aux.cu — the small kernels that close the training render around march / field / composite so that one SDS step has
// no eager-PyTorch arithmetic and no host synchronisation between the pose upload and the optimiser step:
//   * background: freq_encode(rays_d, 6) -> Linear 39-32 + ReLU -> Linear 32-3 -> sigmoid  (nerf/network_grid.py:141-147,
//     freqencoder/src/freqencoder.cu:30-60) fused with the background mix  image + (1 - weights_sum) * bg  (nerf/renderer.py:796-808)
//     and the NCHW layout the guidance wants (nerf/utils.py:545-549); its backward produces the gradient of the composited colour /
//     opacity and accumulates the four bg_net gradients;
//   * regularisers: entropy of the sample weights and the orientation loss (nerf/utils.py:690-704, nerf/renderer.py:741-743) as
//     means over the DEVICE-side sample count, plus their closed-form gradients;
//   * occupancy refresh: jittered cell centres in Morton order, decayed max-update of the density grid with its running mean, and
//     bit packing against min(mean, density_thresh) read from device memory (nerf/renderer.py:1103-1149; raymarching.cu:268-289).
// All HBM-bound streaming kernels with a few bytes per element; none appears above 0.5 % of a step.
#include "common.cuh"

namespace {

constexpr int kBgIn = 39, kBgHid = 32;      // 3 + 3 * 2 * 6 frequency features, hidden width (nerf/network_grid.py:60-66)

__device__ __forceinline__ float rh(float v, int on) { return on ? __half2float(__float2half_rn(v)) : v; }

// the 39 frequency features of one direction (freqencoder.cu:46-56: cos is a phase-shifted fast sine)
__device__ __forceinline__ void freq39(const float d[3], float enc[kBgIn]) {
#pragma unroll
    for (int c = 0; c < kBgIn; c++) {
        if (c < 3) { enc[c] = d[c]; continue; }
        const int col = c / 3 - 1, dd = c % 3, freq = col / 2;
        const float phase = (float)(col % 2) * (3.141592653589793f / 2);
        enc[c] = __sinf(scalbnf(d[dd], freq) + phase);
    }
}

struct BgWeights { float w1[kBgHid * kBgIn]; float b1[kBgHid]; float w2[3 * kBgHid]; float b2[3]; };

__device__ __forceinline__ void load_bg_weights(BgWeights& s, const float* w1, const float* b1, const float* w2, const float* b2, int hr) {
    for (int i = threadIdx.x; i < kBgHid * kBgIn; i += blockDim.x) s.w1[i] = rh(w1[i], hr);
    for (int i = threadIdx.x; i < kBgHid; i += blockDim.x) s.b1[i] = rh(b1[i], hr);
    for (int i = threadIdx.x; i < 3 * kBgHid; i += blockDim.x) s.w2[i] = rh(w2[i], hr);
    for (int i = threadIdx.x; i < 3; i += blockDim.x) s.b2[i] = rh(b2[i], hr);
}

// hidden activations (post-ReLU) and sigmoid outputs of one ray; hr: round where fp16 autocast rounds
__device__ __forceinline__ void bg_mlp(const BgWeights& s, const float enc[kBgIn], float hid[kBgHid], float out[3], int hr) {
#pragma unroll
    for (int j = 0; j < kBgHid; j++) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < kBgIn; i++) a = fmaf(rh(enc[i], hr), s.w1[j * kBgIn + i], a);
        hid[j] = fmaxf(rh(a + s.b1[j], hr), 0.f);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < kBgHid; j++) a = fmaf(hid[j], s.w2[c * kBgHid + j], a);
        out[c] = rh(1.f / (1.f + __expf(-rh(a + s.b2[c], hr))), hr);
    }
}

__global__ void __launch_bounds__(128) k_background_fwd(const float* __restrict__ rays_d, uint32_t N, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                        const float* __restrict__ bg_const, int hr, const float* __restrict__ image_c,
                                                        const float* __restrict__ ws, float* __restrict__ bg_out, float* __restrict__ image,
                                                        float* __restrict__ pred, uint32_t HW, uint32_t C) {
    __shared__ BgWeights s;
    const bool net = w1 != nullptr;
    if (net) load_bg_weights(s, w1, b1, w2, b2, hr);
    __syncthreads();
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float bg[3];
    if (net) {
        const float d[3] = {rays_d[(size_t)n * 3], rays_d[(size_t)n * 3 + 1], rays_d[(size_t)n * 3 + 2]};
        float enc[kBgIn], hid[kBgHid];
        freq39(d, enc);
        bg_mlp(s, enc, hid, bg, hr);
    } else {
        bg[0] = bg_const[0]; bg[1] = bg_const[1]; bg[2] = bg_const[2];
    }
    const float w = ws[n], T = 1.f - w;
    float px[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        px[c] = image_c[(size_t)n * 3 + c] + T * bg[c];
        if (bg_out) bg_out[(size_t)n * 3 + c] = bg[c];
        if (image) image[(size_t)n * 3 + c] = px[c];
    }
    if (pred) {                                     // [B, C, HW] planar: channels 0-2 colour, channel 3 (latent mode) opacity
        const uint32_t b = n / HW, pix = n - b * HW;
        float* o = pred + (size_t)b * C * HW + pix;
        o[0] = px[0]; o[HW] = px[1]; o[2 * (size_t)HW] = px[2];
        if (C > 3) o[3 * (size_t)HW] = w;
    }
}

// one block = 64 rays: per-ray recompute + input gradients, then the block's weight gradients as [feat x 64] x [64 x feat] sums
__global__ void __launch_bounds__(64) k_background_bwd(const float* __restrict__ g_image, const float* __restrict__ g_pred, uint32_t HW, uint32_t C,
                                                        const float* __restrict__ rays_d, uint32_t N, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                        const float* __restrict__ bg_const, int hr, const float* __restrict__ ws,
                                                        float* __restrict__ g_image_c, float* __restrict__ g_ws, float* __restrict__ gw1,
                                                        float* __restrict__ gb1, float* __restrict__ gw2, float* __restrict__ gb2) {
    __shared__ BgWeights s;
    __shared__ float s_enc[64][kBgIn + 1];
    __shared__ float s_dh[64][kBgHid + 1];
    __shared__ float s_hid[64][kBgHid + 1];
    __shared__ float s_do[64][4];
    const bool net = w1 != nullptr;
    if (net) load_bg_weights(s, w1, b1, w2, b2, hr);
    __syncthreads();
    const int tid = threadIdx.x;
    // weight-gradient partial sums stay in registers across the block's chunks of 64 rays and are flushed once: at 512x512 (DMTet stage) a flush per
    // chunk meant 4 096 blocks x 1 379 same-address global atomics
    constexpr int kW1PerThread = (kBgHid * kBgIn + 63) / 64, kW2PerThread = (3 * kBgHid + 63) / 64;
    float acc_w1[kW1PerThread], acc_w2[kW2PerThread], acc_b1 = 0.f, acc_b2 = 0.f;
#pragma unroll
    for (int q = 0; q < kW1PerThread; q++) acc_w1[q] = 0.f;
#pragma unroll
    for (int q = 0; q < kW2PerThread; q++) acc_w2[q] = 0.f;
    for (uint32_t chunk = blockIdx.x; (size_t)chunk * 64 < N; chunk += gridDim.x) {
        const uint32_t n = chunk * 64 + tid;
        const bool in = n < N;
        float g[3] = {0.f, 0.f, 0.f}, g_w_extra = 0.f;
        if (in) {
            if (g_pred) {
                const uint32_t b = n / HW, pix = n - b * HW;
                const float* gp = g_pred + (size_t)b * C * HW + pix;
                g[0] = gp[0]; g[1] = gp[HW]; g[2] = gp[2 * (size_t)HW];
                if (C > 3) g_w_extra = gp[3 * (size_t)HW];
            }
            if (g_image) { g[0] += g_image[(size_t)n * 3]; g[1] += g_image[(size_t)n * 3 + 1]; g[2] += g_image[(size_t)n * 3 + 2]; }
        }
        float bg[3] = {0.f, 0.f, 0.f}, hid[kBgHid], enc[kBgIn];
        if (net) {
            float d[3] = {0.f, 0.f, 0.f};
            if (in) { d[0] = rays_d[(size_t)n * 3]; d[1] = rays_d[(size_t)n * 3 + 1]; d[2] = rays_d[(size_t)n * 3 + 2]; }
            freq39(d, enc);
            bg_mlp(s, enc, hid, bg, hr);
        } else if (in) {
            bg[0] = bg_const[0]; bg[1] = bg_const[1]; bg[2] = bg_const[2];
        }
        const float T = in ? 1.f - ws[n] : 0.f;
        if (in) {
            g_image_c[(size_t)n * 3] = g[0]; g_image_c[(size_t)n * 3 + 1] = g[1]; g_image_c[(size_t)n * 3 + 2] = g[2];
            g_ws[n] = g_w_extra - (g[0] * bg[0] + g[1] * bg[1] + g[2] * bg[2]);
        }
        if (!net) continue;
        // d bg -> pre-sigmoid -> hidden
        float dz[3];
#pragma unroll
        for (int c = 0; c < 3; c++) dz[c] = in ? T * g[c] * bg[c] * (1.f - bg[c]) : 0.f;
#pragma unroll
        for (int j = 0; j < kBgHid; j++) {
            const float dh = dz[0] * s.w2[j] + dz[1] * s.w2[kBgHid + j] + dz[2] * s.w2[2 * kBgHid + j];
            s_dh[tid][j] = hid[j] > 0.f ? dh : 0.f;
            s_hid[tid][j] = hid[j];
        }
#pragma unroll
        for (int i = 0; i < kBgIn; i++) s_enc[tid][i] = rh(enc[i], hr);
        s_do[tid][0] = dz[0]; s_do[tid][1] = dz[1]; s_do[tid][2] = dz[2];
        __syncthreads();
        // weight gradients of this chunk's 64 rays
#pragma unroll
        for (int q = 0; q < kW1PerThread; q++) {
            const int e = tid + 64 * q;
            if (e < kBgHid * kBgIn) {
                const int j = e / kBgIn, i = e - j * kBgIn;
                float a = 0.f;
                for (int r = 0; r < 64; r++) a = fmaf(s_dh[r][j], s_enc[r][i], a);
                acc_w1[q] += a;
            }
        }
        if (tid < kBgHid) {
            float a = 0.f;
            for (int r = 0; r < 64; r++) a += s_dh[r][tid];
            acc_b1 += a;
        }
#pragma unroll
        for (int q = 0; q < kW2PerThread; q++) {
            const int e = tid + 64 * q;
            if (e < 3 * kBgHid) {
                const int c = e / kBgHid, j = e - c * kBgHid;
                float a = 0.f;
                for (int r = 0; r < 64; r++) a = fmaf(s_do[r][c], s_hid[r][j], a);
                acc_w2[q] += a;
            }
        }
        if (tid >= 32 && tid < 35) {
            const int c = tid - 32;
            float a = 0.f;
            for (int r = 0; r < 64; r++) a += s_do[r][c];
            acc_b2 += a;
        }
        __syncthreads();                     // the staging arrays are rewritten by the next chunk
    }
    if (!net) return;
#pragma unroll
    for (int q = 0; q < kW1PerThread; q++) { const int e = tid + 64 * q; if (e < kBgHid * kBgIn) atomicAdd(&gw1[e], acc_w1[q]); }
    if (tid < kBgHid) atomicAdd(&gb1[tid], acc_b1);
#pragma unroll
    for (int q = 0; q < kW2PerThread; q++) { const int e = tid + 64 * q; if (e < 3 * kBgHid) atomicAdd(&gw2[e], acc_w2[q]); }
    if (tid >= 32 && tid < 35) atomicAdd(&gb2[tid - 32], acc_b2);
}

// ---------------------------------------------------------------- regularisers
__device__ __forceinline__ float entropy_term(float w, float* dw) {
    const float a = fminf(fmaxf(w, 1e-5f), 1.f - 1e-5f);
    const float la = log2f(a), lb = log2f(1.f - a);
    if (dw) *dw = (w >= 1e-5f && w <= 1.f - 1e-5f) ? (lb - la) : 0.f;         // d/da [-a log2 a - (1-a) log2(1-a)] = log2((1-a)/a)
    return -a * la - (1.f - a) * lb;
}

__device__ __forceinline__ float orient_dot(const float* nrm, const float* dir, size_t m, float dh[3]) {
    const float dx = dir[m * 3], dy = dir[m * 3 + 1], dz = dir[m * 3 + 2];
    const float inv = 1.f / sqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-20f));       // safe_normalize (nerf/renderer.py:733)
    dh[0] = dx * inv; dh[1] = dy * inv; dh[2] = dz * inv;
    return fmaxf(nrm[m * 3] * dh[0] + nrm[m * 3 + 1] * dh[1] + nrm[m * 3 + 2] * dh[2], 0.f);
}

// out[0] = mean entropy, out[1] = mean orientation loss over the m live samples; scratch: float[2] sums + uint32 ticket (zeroed by the caller)
__global__ void __launch_bounds__(256) k_regularizers_fwd(const float* __restrict__ weights, const float* __restrict__ normals,
                                                          const float* __restrict__ dirs, const int* __restrict__ m_dev, uint32_t M_cap,
                                                          float* __restrict__ scratch, float* __restrict__ out) {
    const uint32_t M = m_dev ? min((uint32_t)max(*m_dev, 0), M_cap) : M_cap;
    float se = 0.f, so = 0.f;
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const float w = weights[m];
        se += entropy_term(w, nullptr);
        if (normals) { float dh[3]; const float c = orient_dot(normals, dirs, m, dh); so += w * c * c; }
    }
    se = warp_sum(se); so = warp_sum(so);
    __shared__ float r0[8], r1[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = se; r1[threadIdx.x >> 5] = so; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; i++) { a += r0[i]; b += r1[i]; }
        atomicAdd(&scratch[0], a); atomicAdd(&scratch[1], b);
        __threadfence();
        const uint32_t ticket = atomicAdd(reinterpret_cast<uint32_t*>(scratch + 2), 1u);
        last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        const float inv = M > 0 ? 1.f / (float)M : 0.f;
        out[0] = atomicAdd(&scratch[0], 0.f) * inv;
        out[1] = atomicAdd(&scratch[1], 0.f) * inv;
    }
}

// g_out[2] = upstream gradients of the two means; g_weights[m] = g_e * dH/dw / m ; g_normals[m] = g_o * w * 2 max(n.d, 0) d / m
// (weights are detached in the orientation loss)
__global__ void __launch_bounds__(256) k_regularizers_bwd(const float* __restrict__ g_out, float lambda_entropy, float lambda_orient,
                                                          const float* __restrict__ weights, const float* __restrict__ normals,
                                                          const float* __restrict__ dirs, const int* __restrict__ m_dev, uint32_t M_cap,
                                                          float* __restrict__ g_weights, float* __restrict__ g_normals) {
    const uint32_t M = m_dev ? min((uint32_t)max(*m_dev, 0), M_cap) : M_cap;
    const float inv = M > 0 ? 1.f / (float)M : 0.f;
    const float ge = g_out[0] * lambda_entropy * inv, go = g_out[1] * lambda_orient * inv;
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const float w = weights[m];
        float dw;
        entropy_term(w, &dw);
        g_weights[m] = ge * dw;
        if (normals && g_normals) {
            float dh[3];
            const float c = orient_dot(normals, dirs, m, dh);
            const float k = go * w * 2.f * c;
            g_normals[(size_t)m * 3] = k * dh[0]; g_normals[(size_t)m * 3 + 1] = k * dh[1]; g_normals[(size_t)m * 3 + 2] = k * dh[2];
        }
    }
}

// ---------------------------------------------------------------- occupancy refresh
__device__ __forceinline__ uint32_t compact_bits(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// cell `idx` (Morton order, as density_grid is indexed): centre in [-(b - h), b - h] plus a uniform jitter of +-h, h = b / G
__global__ void k_occupancy_points(const float* __restrict__ noise, uint32_t n, uint32_t G, float bound_cas, float* __restrict__ xyzs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float half = bound_cas / (float)G, span = bound_cas - half;
    const uint32_t c[3] = {compact_bits(i), compact_bits(i >> 1), compact_bits(i >> 2)};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float x = 2.f * (float)c[d] / (float)(G - 1) - 1.f;
        xyzs[(size_t)i * 3 + d] = x * span + (noise[(size_t)i * 3 + d] * 2.f - 1.f) * half;
    }
}

// grid <- max(grid * decay, sigma) on cells with grid >= 0; acc[0] += sum of the updated valid cells, acc[1] += their count
__global__ void __launch_bounds__(256) k_occupancy_update(float* __restrict__ grid, const float* __restrict__ sigmas, uint32_t n, float decay,
                                                          float* __restrict__ acc) {
    float s = 0.f, c = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        if (g >= 0.f) { g = fmaxf(g * decay, sigmas[i]); grid[i] = g; s += g; c += 1.f; }
    }
    s = warp_sum(s); c = warp_sum(c);
    __shared__ float r0[8], r1[8];
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s; r1[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; i++) { a += r0[i]; b += r1[i]; }
        atomicAdd(&acc[0], a); atomicAdd(&acc[1], b);
    }
}

__global__ void k_packbits_mean(const float4* __restrict__ grid, uint32_t N, const float* __restrict__ acc, float density_thresh,
                                uint8_t* __restrict__ bitfield, float* __restrict__ mean_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const float mean = acc[1] > 0.f ? acc[0] / acc[1] : 0.f;
    if (n == 0 && mean_out) *mean_out = mean;
    if (n >= N) return;
    const float thresh = fminf(mean, density_thresh);
    const float4 a = grid[(size_t)n * 2], b = grid[(size_t)n * 2 + 1];
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// pinhole rays of B cameras (nerf/utils.py:113-176 with N = -1): pixel centres, z = -1, unnormalised directions R * (x, y, z)
__global__ void k_get_rays(const float* __restrict__ poses, uint32_t B, uint32_t H, uint32_t W, float focal, float cx, float cy,
                           uint32_t first, uint32_t stride, uint32_t per_view, float* __restrict__ rays_o, float* __restrict__ rays_d) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * per_view) return;
    const uint32_t b = t / per_view, k = t - b * per_view, pix = first + k * stride;
    const float i = (float)(pix % W) + 0.5f, j = (float)(pix / W) + 0.5f;
    const float zs = -1.f, xs = -(i - cx) / focal * zs, ys = (j - cy) / focal * zs;
    const float* P = poses + (size_t)b * 16;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        rays_d[(size_t)t * 3 + r] = xs * P[r * 4] + ys * P[r * 4 + 1] + zs * P[r * 4 + 2];
        rays_o[(size_t)t * 3 + r] = P[r * 4 + 3];
    }
}

// out[offset .. offset + count) of every ray <- the ray's 3-vector (warp per ray, coalesced): per-view light directions for the samples
__global__ void k_expand_ray_vec3(const float* __restrict__ values, const int* __restrict__ rays, uint32_t N, uint32_t cap, float* __restrict__ out) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[n * 2], cnt = (uint32_t)rays[n * 2 + 1];
    if (off + cnt > cap) return;
    const float v0 = values[(size_t)n * 3], v1 = values[(size_t)n * 3 + 1], v2 = values[(size_t)n * 3 + 2];
    for (uint32_t i = lane; i < cnt * 3; i += 32) out[(size_t)off * 3 + i] = (i % 3 == 0) ? v0 : ((i % 3 == 1) ? v1 : v2);
}

}  // namespace

/* values [N,3] per ray -> out [cap,3] per sample through rays[N,2] = (offset, count); rows no ray owns are left untouched */
SDF_API int sdf_expand_ray_vec3(const float* values, const int* rays, uint32_t N, uint32_t cap, float* out, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(values && rays && out, "expand_ray_vec3: null pointer");
    k_expand_ray_vec3<<<cdiv(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(values, rays, N, cap, out);
    SDF_CHECK_LAUNCH("expand_ray_vec3");
    return SDF_OK;
}

SDF_API int sdf_background_forward(const float* rays_d, uint32_t N, const float* w1, const float* b1, const float* w2, const float* b2,
                                   const float* bg_const, int half_round, const float* image_c, const float* weights_sum,
                                   float* bg, float* image, float* pred, uint32_t HW, uint32_t C, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(image_c && weights_sum && (image || pred), "background_forward: null pointer");
    SDF_CHECK_ARG((w1 && b1 && w2 && b2 && rays_d) || bg_const, "background_forward: either the bg_net parameters + rays_d or a constant colour");
    SDF_CHECK_ARG(!pred || (HW > 0 && N % HW == 0 && (C == 3 || C == 4)), "background_forward: pred needs N = B * HW and C in {3, 4}");
    k_background_fwd<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_d, N, w1, b1, w2, b2, bg_const, half_round, image_c, weights_sum, bg, image,
                                                                     pred, HW, C);
    SDF_CHECK_LAUNCH("background_forward");
    return SDF_OK;
}

SDF_API int sdf_background_backward(const float* g_image, const float* g_pred, uint32_t HW, uint32_t C, const float* rays_d, uint32_t N,
                                    const float* w1, const float* b1, const float* w2, const float* b2, const float* bg_const, int half_round,
                                    const float* weights_sum, float* g_image_c, float* g_weights_sum, float* gw1, float* gb1, float* gw2,
                                    float* gb2, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG((g_image || g_pred) && weights_sum && g_image_c && g_weights_sum, "background_backward: null pointer");
    SDF_CHECK_ARG(!w1 || (b1 && w2 && b2 && rays_d && gw1 && gb1 && gw2 && gb2), "background_backward: bg_net needs all parameters and gradient buffers");
    SDF_CHECK_ARG(w1 || bg_const, "background_backward: either the bg_net parameters or a constant colour");
    SDF_CHECK_ARG(!g_pred || (HW > 0 && N % HW == 0 && (C == 3 || C == 4)), "background_backward: g_pred needs N = B * HW and C in {3, 4}");
    k_background_bwd<<<min(cdiv(N, 64), (uint32_t)(6 * sdf_num_sms())), 64, 0, (cudaStream_t)stream>>>(g_image, g_pred, HW, C, rays_d, N, w1, b1, w2, b2, bg_const, half_round, weights_sum,
                                                                     g_image_c, g_weights_sum, gw1, gb1, gw2, gb2);
    SDF_CHECK_LAUNCH("background_backward");
    return SDF_OK;
}

SDF_API int sdf_render_regularizers_forward(const float* weights, const float* normals, const float* dirs, const int* m_dev, uint32_t M_cap,
                                            float* scratch, float* out, void* stream) {
    SDF_CHECK_ARG(weights && scratch && out && (!normals || dirs), "render_regularizers_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SDF_CHECK_CUDA(cudaMemsetAsync(scratch, 0, 3 * sizeof(float), st));
    const uint32_t blocks = M_cap == 0 ? 1u : min((uint32_t)(sdf_num_sms() * 4), cdiv(M_cap, 256));
    k_regularizers_fwd<<<blocks, 256, 0, st>>>(weights, normals, dirs, m_dev, M_cap, scratch, out);
    SDF_CHECK_LAUNCH("render_regularizers_forward");
    return SDF_OK;
}

SDF_API int sdf_render_regularizers_backward(const float* g_out, float lambda_entropy, float lambda_orient, const float* weights,
                                             const float* normals, const float* dirs, const int* m_dev, uint32_t M_cap, float* g_weights,
                                             float* g_normals, void* stream) {
    if (M_cap == 0) return SDF_OK;
    SDF_CHECK_ARG(g_out && weights && g_weights && (!normals || dirs), "render_regularizers_backward: null pointer");
    const uint32_t blocks = min((uint32_t)(sdf_num_sms() * 4), cdiv(M_cap, 256));
    k_regularizers_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(g_out, lambda_entropy, lambda_orient, weights, normals, dirs, m_dev, M_cap, g_weights,
                                                                 g_normals);
    SDF_CHECK_LAUNCH("render_regularizers_backward");
    return SDF_OK;
}

SDF_API int sdf_occupancy_points(const float* noise, uint32_t n, uint32_t grid_size, float bound_cas, float* xyzs, void* stream) {
    if (n == 0) return SDF_OK;
    SDF_CHECK_ARG(noise && xyzs && grid_size > 1 && n <= grid_size * grid_size * grid_size, "occupancy_points: bad arguments");
    k_occupancy_points<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(noise, n, grid_size, bound_cas, xyzs);
    SDF_CHECK_LAUNCH("occupancy_points");
    return SDF_OK;
}

SDF_API int sdf_occupancy_update(float* grid, const float* sigmas, uint32_t n, float decay, float* acc, void* stream) {
    if (n == 0) return SDF_OK;
    SDF_CHECK_ARG(grid && sigmas && acc, "occupancy_update: null pointer");
    k_occupancy_update<<<min((uint32_t)(sdf_num_sms() * 8), cdiv(n, 256)), 256, 0, (cudaStream_t)stream>>>(grid, sigmas, n, decay, acc);
    SDF_CHECK_LAUNCH("occupancy_update");
    return SDF_OK;
}

SDF_API int sdf_packbits_mean(const float* grid, uint32_t N, const float* acc, float density_thresh, uint8_t* bitfield, float* mean_out,
                              void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(grid && acc && bitfield && ((uintptr_t)grid & 15) == 0, "packbits_mean: bad arguments");
    k_packbits_mean<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(grid), N, acc, density_thresh, bitfield, mean_out);
    SDF_CHECK_LAUNCH("packbits_mean");
    return SDF_OK;
}

SDF_API int sdf_get_rays(const float* poses, uint32_t B, uint32_t H, uint32_t W, float focal, float cx, float cy, uint32_t first, uint32_t stride,
                         float* rays_o, float* rays_d, void* stream) {
    SDF_CHECK_ARG(poses && rays_o && rays_d && stride >= 1 && first < stride, "get_rays: bad arguments");
    const uint32_t per_view = (H * W - first + stride - 1) / stride;
    if (B * per_view == 0) return SDF_OK;
    k_get_rays<<<cdiv(B * per_view, 256), 256, 0, (cudaStream_t)stream>>>(poses, B, H, W, focal, cx, cy, first, stride, per_view, rays_o, rays_d);
    SDF_CHECK_LAUNCH("get_rays");
    return SDF_OK;
}
