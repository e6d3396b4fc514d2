// adan.cu — fused Adan step for the NeRF parameters (reference: optimizer.py:102-258, the foreach=False path selected by
// main.py:368, plus the GradScaler protocol around it, nerf/utils.py:1063-1067).
//
// The reference runs ~20 elementwise kernels over 12.2 M parameters x 4 state tensors (~1 GB of traffic) and two host syncs
// (.item() on the clip factor, GradScaler's inf check).  Here:
//   pass 1  sdf_adan_grad_norm : sum of squares of the (unscaled) gradients of every parameter tensor -> device scalar,
//                                plus a non-finite flag (the GradScaler inf check);
//   pass 2  sdf_adan_step      : clip factor from the device scalar, moments, parameter update, -g stash for the next
//                                step, optional fp16 mirror of the updated parameter (the hash table's working copy) and
//                                optional gradient zeroing — one read/write of each stream (p, g, m, d, n, pre: ~0.34 GB).
// Nothing is read back to the host; a step with non-finite gradients leaves parameters and state untouched.
// Roofline: HBM.  Algorithmic bytes per parameter: 4 B x (read p,g,m,d,n,pre + write p,m,d,n,pre) = 44 B (+2 B fp16 mirror, +4 B zeroing).
#include "common.cuh"

namespace {

// acc[0] += sum(g^2) * inv_scale^2 ; acc[1] = 1 if any non-finite
__global__ void __launch_bounds__(256) k_grad_norm(const float* __restrict__ g, long long n, float inv_scale, float* __restrict__ acc) {
    float s = 0.f;
    bool bad = false;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const float4 v = *reinterpret_cast<const float4*>(g + i);
            const float a = v.x * inv_scale, b = v.y * inv_scale, c = v.z * inv_scale, d = v.w * inv_scale;
            s += a * a + b * b + c * c + d * d;
            bad |= !isfinite(v.x) | !isfinite(v.y) | !isfinite(v.z) | !isfinite(v.w);
        } else {
            for (long long j = i; j < n; j++) { const float a = g[j] * inv_scale; s += a * a; bad |= !isfinite(g[j]); }
        }
    }
    s = warp_sum(s);
    __shared__ float red[8];
    __shared__ int anybad;
    if (threadIdx.x == 0) anybad = 0;
    __syncthreads();
    if (bad) anybad = 1;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        atomicAdd(&acc[0], t);
        if (anybad) acc[1] = 1.f;
    }
}

struct AdanHyper {
    float beta1, beta2, beta3, bc1, bc2, bc3_sqrt, lr, weight_decay, eps, max_grad_norm, inv_scale;
    int first_step, no_prox;
};

__global__ void __launch_bounds__(256) k_adan_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ d,
                                                   float* __restrict__ nn, float* __restrict__ pre, long long n, AdanHyper h,
                                                   const float* __restrict__ acc, __half* __restrict__ p_half, int zero_grad) {
    if (acc[1] != 0.f) return;                                      // non-finite gradients: skip the step (GradScaler semantics)
    float clip = 1.f;
    if (h.max_grad_norm > 0.f) clip = fminf(h.max_grad_norm / (sqrtf(acc[0]) + h.eps), 1.0f);
    const float gs = clip * h.inv_scale;
    const float step_size = h.lr / h.bc1, step_size_diff = h.lr * h.beta2 / h.bc2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * gs;
    const float prev = h.first_step ? -gi : pre[i];                 // neg_pre_grad initialised to -g on the first step
    const float diff = prev + gi;
    const float mi = h.beta1 * m[i] + (1.f - h.beta1) * gi;
    const float di = h.beta2 * d[i] + (1.f - h.beta2) * diff;
    const float u = h.beta2 * diff + gi;
    const float ni = h.beta3 * nn[i] + (1.f - h.beta3) * u * u;
    const float denom = sqrtf(ni) / h.bc3_sqrt + h.eps;
    float pi = p[i];
    if (h.no_prox) { pi *= 1.f - h.lr * h.weight_decay; pi -= step_size * mi / denom; pi -= step_size_diff * di / denom; }
    else { pi -= step_size * mi / denom; pi -= step_size_diff * di / denom; pi /= 1.f + h.lr * h.weight_decay; }
    p[i] = pi; m[i] = mi; d[i] = di; nn[i] = ni; pre[i] = -gi;
    if (p_half) p_half[i] = __float2half_rn(pi);
    if (zero_grad) g[i] = 0.f;
}

}  // namespace

// acc: device float[2], zeroed once per optimiser step by the caller (sdf_adan_begin), accumulated over all tensors
SDF_API int sdf_adan_begin(float* acc, void* stream) {
    SDF_CHECK_ARG(acc, "adan_begin: null pointer");
    SDF_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(float), (cudaStream_t)stream));
    return SDF_OK;
}

SDF_API int sdf_adan_grad_norm(const float* grad, long long n, float inv_scale, float* acc, void* stream) {
    if (n == 0) return SDF_OK;
    SDF_CHECK_ARG(grad && acc && ((uintptr_t)grad & 15) == 0, "adan_grad_norm: bad arguments");
    const long long blocks = (n / 4 + 255) / 256;
    const int grid = (int)(blocks < kNumSMs * 8 ? (blocks > 0 ? blocks : 1) : kNumSMs * 8);
    k_grad_norm<<<grid, 256, 0, (cudaStream_t)stream>>>(grad, n, inv_scale, acc);
    SDF_CHECK_LAUNCH("adan_grad_norm");
    return SDF_OK;
}

// One parameter tensor.  step: 1-based optimiser step (bias corrections 1 - beta^step).  p_half (optional): fp16 mirror of the
// updated parameter.  zero_grad != 0 clears grad after use.  acc: from sdf_adan_begin + sdf_adan_grad_norm over ALL tensors.
SDF_API int sdf_adan_step(float* param, float* grad, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq, float* neg_pre_grad, long long n,
                          float beta1, float beta2, float beta3, int step, float lr, float weight_decay, float eps, float max_grad_norm,
                          int no_prox, float inv_scale, const float* acc, void* param_half, int zero_grad, void* stream) {
    if (n == 0) return SDF_OK;
    SDF_CHECK_ARG(param && grad && exp_avg && exp_avg_diff && exp_avg_sq && neg_pre_grad && acc && step >= 1, "adan_step: bad arguments");
    AdanHyper h;
    h.beta1 = beta1; h.beta2 = beta2; h.beta3 = beta3;
    h.bc1 = 1.f - powf(beta1, (float)step); h.bc2 = 1.f - powf(beta2, (float)step); h.bc3_sqrt = sqrtf(1.f - powf(beta3, (float)step));
    h.lr = lr; h.weight_decay = weight_decay; h.eps = eps; h.max_grad_norm = max_grad_norm; h.inv_scale = inv_scale;
    h.first_step = step == 1; h.no_prox = no_prox;
    k_adan_step<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_diff, exp_avg_sq, neg_pre_grad, n, h, acc,
                                                                              (__half*)param_half, zero_grad);
    SDF_CHECK_LAUNCH("adan_step");
    return SDF_OK;
}
