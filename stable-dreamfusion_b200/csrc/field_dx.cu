// field_dx.cu — d(albedo) / d(position) of the texture lookup in the DMTet stage.
//
// The reference textures the rasterised mesh with albedo = sigmoid(sigma_net(grid_encode(x))[1:]) at the interpolated surface points
// (nerf/renderer.py:905-912, nerf/network_grid.py:68-79) and, because those points carry the mesh's autograd graph, back-propagates the colour
// gradient through the MLP and the hash encoding INTO the positions (gridencoder/grid.py:77-100: grad_inputs from dy_dx).  The fused field
// kernels of the volume stage have no position gradient (sample positions are constants there), so this path adds it on the side:
// sdf_grid_encode_forward (csrc/gridenc.cu, the drop-in encoder kernel) supplies the features and dy_dx of every pixel's point, and this
// kernel runs the 32-64-64-4 MLP forward + the data-gradient back to the 32 features per point and contracts it with dy_dx.
// One thread per pixel, weights in shared memory rounded to fp16 like the autocast reference, fp32 accumulation; ~25 k FMAs per covered pixel,
// 262 144 pixels at 512x512: FMA-pipe work of a few hundred microseconds, far below the guidance's 10 ms.
#include "common.cuh"

namespace {

constexpr int kIn = 32, kHid = 64, kOut = 4;

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }

__global__ void __launch_bounds__(128) k_albedo_input_grad(const __half* __restrict__ feat, const __half* __restrict__ dy_dx, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                           const float* __restrict__ w3, const float* __restrict__ b3, const float* __restrict__ g_albedo,
                                                           const float* __restrict__ mask, int P, int L, float inv_2bound, float* __restrict__ d_xyz) {
    __shared__ float s_w1[kHid * kIn], s_w2[kHid * kHid], s_w3[kOut * kHid], s_b1[kHid], s_b2[kHid], s_b3[kOut];
    for (int i = threadIdx.x; i < kHid * kIn; i += blockDim.x) s_w1[i] = rh(w1[i]);
    for (int i = threadIdx.x; i < kHid * kHid; i += blockDim.x) s_w2[i] = rh(w2[i]);
    for (int i = threadIdx.x; i < kOut * kHid; i += blockDim.x) s_w3[i] = rh(w3[i]);
    for (int i = threadIdx.x; i < kHid; i += blockDim.x) { s_b1[i] = rh(b1[i]); s_b2[i] = rh(b2[i]); }
    if (threadIdx.x < kOut) s_b3[threadIdx.x] = rh(b3[threadIdx.x]);
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float out[3] = {0.f, 0.f, 0.f};
    const float g[3] = {g_albedo[3 * (size_t)p], g_albedo[3 * (size_t)p + 1], g_albedo[3 * (size_t)p + 2]};
    if (mask[p] > 0.f && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f)) {
        float f[kIn], h1[kHid], h2[kHid];
#pragma unroll
        for (int i = 0; i < kIn; i++) f[i] = i < 2 * L ? __half2float(feat[(size_t)p * 2 * L + i]) : 0.f;
        for (int j = 0; j < kHid; j++) {
            float a = s_b1[j];
#pragma unroll
            for (int i = 0; i < kIn; i++) a = fmaf(s_w1[j * kIn + i], f[i], a);
            h1[j] = rh(fmaxf(a, 0.f));
        }
        for (int j = 0; j < kHid; j++) {
            float a = s_b2[j];
#pragma unroll
            for (int i = 0; i < kHid; i++) a = fmaf(s_w2[j * kHid + i], h1[i], a);
            h2[j] = rh(fmaxf(a, 0.f));
        }
        float go[kOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float a = s_b3[1 + c];
#pragma unroll
            for (int i = 0; i < kHid; i++) a = fmaf(s_w3[(1 + c) * kHid + i], h2[i], a);
            const float sg = 1.f / (1.f + __expf(-rh(a)));
            go[1 + c] = g[c] * sg * (1.f - sg);
        }
        // back through the two hidden layers: gh2 overwrites h2, gh1 overwrites h1 (the ReLU masks are read first)
        float gh1[kHid];
#pragma unroll
        for (int i = 0; i < kHid; i++) {
            const float v = s_w3[1 * kHid + i] * go[1] + s_w3[2 * kHid + i] * go[2] + s_w3[3 * kHid + i] * go[3];
            h2[i] = h2[i] > 0.f ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kHid; i++) gh1[i] = 0.f;
        for (int j = 0; j < kHid; j++) {
            const float v = h2[j];
            if (v == 0.f) continue;
#pragma unroll
            for (int i = 0; i < kHid; i++) gh1[i] = fmaf(s_w2[j * kHid + i], v, gh1[i]);
        }
#pragma unroll
        for (int i = 0; i < kHid; i++) gh1[i] = h1[i] > 0.f ? gh1[i] : 0.f;
        float gf[kIn];
#pragma unroll
        for (int i = 0; i < kIn; i++) gf[i] = 0.f;
        for (int j = 0; j < kHid; j++) {
            const float v = gh1[j];
            if (v == 0.f) continue;
#pragma unroll
            for (int i = 0; i < kIn; i++) gf[i] = fmaf(s_w1[j * kIn + i], v, gf[i]);
        }
        // grad_inputs[d] = sum_{l, c} gf[2 l + c] * dy_dx[p, l, d, c]     (gridencoder.cu:353-378)
        const __half* dd = dy_dx + (size_t)p * L * 6;
        for (int l = 0; l < L; l++) {
#pragma unroll
            for (int d = 0; d < 3; d++)
                out[d] += gf[2 * l] * __half2float(dd[l * 6 + d * 2]) + gf[2 * l + 1] * __half2float(dd[l * 6 + d * 2 + 1]);
        }
#pragma unroll
        for (int d = 0; d < 3; d++) out[d] *= inv_2bound;          // the encoder sees (x + bound) / (2 bound)
    }
    d_xyz[3 * (size_t)p] = out[0]; d_xyz[3 * (size_t)p + 1] = out[1]; d_xyz[3 * (size_t)p + 2] = out[2];
}

}  // namespace

// d_xyz [P,3] = d(sum g_albedo . albedo(x)) / dx for the points whose mask is set (zeros elsewhere).  feat fp16 [P, 2L] and dy_dx fp16 [P, L, 3, 2] come from
// sdf_grid_encode_forward at (x + bound) / (2 bound); w1 [64,32] b1 [64] w2 [64,64] b2 [64] w3 [4,64] b3 [4] fp32 = sigma_net (nerf/network_grid.py:57).
SDF_API int sdf_field_albedo_input_grad(const void* feat, const void* dy_dx, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                        const float* b3, const float* g_albedo, const float* mask, int P, int L, float bound, float* d_xyz, void* stream) {
    SDF_CHECK_ARG(feat && dy_dx && w1 && b1 && w2 && b2 && w3 && b3 && g_albedo && mask && d_xyz && L >= 1 && 2 * L <= kIn && bound > 0.f,
                  "field_albedo_input_grad: bad arguments");
    if (P > 0) k_albedo_input_grad<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const __half*)feat, (const __half*)dy_dx, w1, b1, w2, b2, w3, b3, g_albedo, mask,
                                                                                       P, L, 1.f / (2.f * bound), d_xyz);
    SDF_CHECK_LAUNCH("field_albedo_input_grad");
    return SDF_OK;
}
