// adan.cu — fused Adan step for the NeRF parameters (reference: optimizer.py:102-258, the foreach=False path selected by
// main.py:368, plus the GradScaler protocol around it, nerf/utils.py:1063-1067, and the torch_ema shadow update of
// nerf/utils.py:282-283,1090-1091).
//
// The reference runs ~20 elementwise kernels over 12.2 M parameters x 4 state tensors (~1 GB of traffic) and two host syncs
// (.item() on the clip factor, GradScaler's inf check).  Here:
//   pass 1  sdf_adan_grad_norm : sum of squares of the (unscaled) gradients of every parameter tensor -> device scalar,
//                                plus a non-finite flag (the GradScaler inf check);
//   pass 1b sdf_adan_advance   : one thread: if the flag is clear, every group's EXECUTED-step counter advances (device int32) —
//                                a skipped step does not touch bias corrections or the first-step initialisation of
//                                neg_pre_grad, exactly as GradScaler.step() not calling optimizer.step();
//   pass 2  sdf_adan_step      : clip factor from the device scalar, moments, parameter update, -g stash for the next
//                                step, optional fp16 mirror of the updated parameter (the hash table's working copy), optional
//                                EMA shadow update and optional gradient zeroing — one read/write of each stream, 16-byte
//                                vectors, grid-stride (p, g, m, d, n, pre: ~0.34 GB per step).
// Nothing is read back to the host; a step with non-finite gradients leaves parameters and state untouched but still clears the
// gradients when asked to (the reference calls optimizer.zero_grad() every iteration, nerf/utils.py:1043).
// Roofline: HBM.  Algorithmic bytes per parameter: 4 B x (read p,g,m,d,n,pre + write p,m,d,n,pre) = 44 B (+2 B fp16 mirror, +4 B zeroing, +8 B EMA).
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

__global__ void k_advance(const float* __restrict__ acc, int* __restrict__ steps, int n_groups) {
    if (acc[1] != 0.f) return;
    for (int i = threadIdx.x; i < n_groups; i += blockDim.x) steps[i] += 1;
}

struct AdanHyper {
    float beta1, beta2, beta3, lr, weight_decay, eps, max_grad_norm, inv_scale, ema_omd;
    int step, no_prox;
};

struct AdanCoef { float gs, b1, b2, b3, step_size, step_size_diff, inv_bc3, eps, wd_mul, wd_div, ema_omd; int first, no_prox; };

__device__ __forceinline__ void adan_one(float& p, float g_in, float& m, float& d, float& nn, float& pre, const AdanCoef& c) {
    const float gi = g_in * c.gs;
    // neg_pre_grad = -g the first time THIS tensor is stepped or on the group's first executed step (optimizer.py:164); the host
    // allocates the buffer as NaN, so 'never stepped' needs no flag and survives skipped steps
    const float prev = (c.first || isnan(pre)) ? -gi : pre;
    const float diff = prev + gi;
    m = c.b1 * m + (1.f - c.b1) * gi;
    d = c.b2 * d + (1.f - c.b2) * diff;
    const float u = c.b2 * diff + gi;
    nn = c.b3 * nn + (1.f - c.b3) * u * u;
    const float denom = sqrtf(nn) * c.inv_bc3 + c.eps;
    if (c.no_prox) { p *= c.wd_mul; p -= c.step_size * m / denom; p -= c.step_size_diff * d / denom; }
    else { p -= c.step_size * m / denom; p -= c.step_size_diff * d / denom; p /= c.wd_div; }
    pre = -gi;
}

__global__ void __launch_bounds__(256) k_adan_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ d,
                                                   float* __restrict__ nn, float* __restrict__ pre, long long n, AdanHyper h,
                                                   const float* __restrict__ acc, const int* __restrict__ step_dev, __half* __restrict__ p_half,
                                                   float* __restrict__ ema, int zero_grad) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const long long n4 = n >> 2;
    if (acc[1] != 0.f) {                                            // non-finite gradients: skip the step (GradScaler semantics) ...
        if (zero_grad) {                                            // ... but the gradients are cleared all the same
            for (long long i = tid; i < n4; i += nthr) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long long i = (n4 << 2) + tid; i < n; i += nthr) g[i] = 0.f;
        }
        return;
    }
    const int step = step_dev ? *step_dev : h.step;
    AdanCoef c;
    float clip = 1.f;
    if (h.max_grad_norm > 0.f) clip = fminf(h.max_grad_norm / (sqrtf(acc[0]) + h.eps), 1.0f);
    c.gs = clip * h.inv_scale;
    c.b1 = h.beta1; c.b2 = h.beta2; c.b3 = h.beta3;
    c.step_size = h.lr / (1.f - powf(h.beta1, (float)step));
    c.step_size_diff = h.lr * h.beta2 / (1.f - powf(h.beta2, (float)step));
    c.inv_bc3 = 1.f / sqrtf(1.f - powf(h.beta3, (float)step));
    c.eps = h.eps; c.wd_mul = 1.f - h.lr * h.weight_decay; c.wd_div = 1.f + h.lr * h.weight_decay;
    c.first = step == 1; c.no_prox = h.no_prox; c.ema_omd = h.ema_omd;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = tid; i < n4; i += nthr) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i], D = reinterpret_cast<float4*>(d)[i], N = reinterpret_cast<float4*>(nn)[i];
        float4 R = reinterpret_cast<float4*>(pre)[i];
        adan_one(P.x, G.x, M.x, D.x, N.x, R.x, c);
        adan_one(P.y, G.y, M.y, D.y, N.y, R.y, c);
        adan_one(P.z, G.z, M.z, D.z, N.z, R.z, c);
        adan_one(P.w, G.w, M.w, D.w, N.w, R.w, c);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(d)[i] = D;
        reinterpret_cast<float4*>(nn)[i] = N; reinterpret_cast<float4*>(pre)[i] = R;
        if (p_half) {
            const __half2 lo = __floats2half2_rn(P.x, P.y), hi = __floats2half2_rn(P.z, P.w);
            uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            reinterpret_cast<uint2*>(p_half)[i] = pk;
        }
        if (ema) {
            float4 E = reinterpret_cast<float4*>(ema)[i];
            E.x -= c.ema_omd * (E.x - P.x); E.y -= c.ema_omd * (E.y - P.y); E.z -= c.ema_omd * (E.z - P.z); E.w -= c.ema_omd * (E.w - P.w);
            reinterpret_cast<float4*>(ema)[i] = E;
        }
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = z4;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nthr) {          // ragged tail (< 4 elements)
        float P = p[i], M = m[i], D = d[i], N = nn[i], R = pre[i];
        adan_one(P, g[i], M, D, N, R, c);
        p[i] = P; m[i] = M; d[i] = D; nn[i] = N; pre[i] = R;
        if (p_half) p_half[i] = __float2half_rn(P);
        if (ema) ema[i] -= c.ema_omd * (ema[i] - P);
        if (zero_grad) g[i] = 0.f;
    }
}

__global__ void __launch_bounds__(256) k_ema(float* __restrict__ ema, const float* __restrict__ p, long long n, float omd) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x, n4 = n >> 2;
    for (long long i = tid; i < n4; i += nthr) {
        float4 E = reinterpret_cast<float4*>(ema)[i];
        const float4 P = reinterpret_cast<const float4*>(p)[i];
        E.x -= omd * (E.x - P.x); E.y -= omd * (E.y - P.y); E.z -= omd * (E.z - P.z); E.w -= omd * (E.w - P.w);
        reinterpret_cast<float4*>(ema)[i] = E;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nthr) ema[i] -= omd * (ema[i] - p[i]);
}

inline int stream_grid(long long n) {
    const long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)sdf_num_sms() * 8;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
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
    k_grad_norm<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(grad, n, inv_scale, acc);
    SDF_CHECK_LAUNCH("adan_grad_norm");
    return SDF_OK;
}

// after every tensor's sdf_adan_grad_norm: steps[0..n_groups) += 1 unless the non-finite flag is set
SDF_API int sdf_adan_advance(const float* acc, int* steps, int n_groups, void* stream) {
    SDF_CHECK_ARG(acc && steps && n_groups >= 0, "adan_advance: bad arguments");
    if (n_groups == 0) return SDF_OK;
    k_advance<<<1, 32, 0, (cudaStream_t)stream>>>(acc, steps, n_groups);
    SDF_CHECK_LAUNCH("adan_advance");
    return SDF_OK;
}

// One parameter tensor.  step: 1-based optimiser step (bias corrections 1 - beta^step), read from step_dev when that is not NULL.
// p_half (optional): fp16 mirror of the updated parameter.  ema (optional): shadow -= ema_one_minus_decay * (shadow - param_new).
// zero_grad != 0 clears grad after use (also on a skipped step).  acc: from sdf_adan_begin + sdf_adan_grad_norm over ALL tensors.
SDF_API int sdf_adan_step(float* param, float* grad, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq, float* neg_pre_grad, long long n,
                          float beta1, float beta2, float beta3, int step, const int* step_dev, float lr, float weight_decay, float eps,
                          float max_grad_norm, int no_prox, float inv_scale, const float* acc, void* param_half, float* ema,
                          float ema_one_minus_decay, int zero_grad, void* stream) {
    if (n == 0) return SDF_OK;
    SDF_CHECK_ARG(param && grad && exp_avg && exp_avg_diff && exp_avg_sq && neg_pre_grad && acc && (step >= 1 || step_dev), "adan_step: bad arguments");
    const uintptr_t al = (uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_diff | (uintptr_t)exp_avg_sq |
                         (uintptr_t)neg_pre_grad | (uintptr_t)ema | (param_half ? ((uintptr_t)param_half << 1) : 0);
    SDF_CHECK_ARG((al & 15) == 0, "adan_step: tensors must be 16-byte aligned (fp16 mirror: 8-byte)");
    AdanHyper h;
    h.beta1 = beta1; h.beta2 = beta2; h.beta3 = beta3;
    h.lr = lr; h.weight_decay = weight_decay; h.eps = eps; h.max_grad_norm = max_grad_norm; h.inv_scale = inv_scale;
    h.ema_omd = ema_one_minus_decay; h.step = step; h.no_prox = no_prox;
    k_adan_step<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_diff, exp_avg_sq, neg_pre_grad, n, h, acc, step_dev,
                                                                  (__half*)param_half, ema, zero_grad);
    SDF_CHECK_LAUNCH("adan_step");
    return SDF_OK;
}

// torch_ema.ExponentialMovingAverage.update for one tensor (nerf/utils.py:1090-1091): shadow -= one_minus_decay * (shadow - param)
SDF_API int sdf_ema_update(float* shadow, const float* param, long long n, float one_minus_decay, void* stream) {
    if (n == 0) return SDF_OK;
    SDF_CHECK_ARG(shadow && param && (((uintptr_t)shadow | (uintptr_t)param) & 15) == 0, "ema_update: bad arguments");
    k_ema<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(shadow, param, n, one_minus_decay);
    SDF_CHECK_LAUNCH("ema_update");
    return SDF_OK;
}
