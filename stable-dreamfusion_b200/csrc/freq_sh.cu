// freq_sh.cu — frequency (sin/cos) and spherical-harmonics encoders, sm_100a.
//
// Operator-parity replacements for freqencoder/src/freqencoder.cu:30-94 and
// shencoder/src/shencoder.cu:28-383 of the reference.
//
// The reference's SH kernel is a 64-entry table of closed-form polynomials plus
// 192 hand-expanded partial derivatives.  They are the real spherical harmonics
//     Y_l^m = (-1)^m N_lm Q_l^|m|(z) * {Re,Im}(x+iy)^|m| * (m ? sqrt2 : 1),
//     Q_l^m = d^m/dz^m P_l(z),  N_lm = sqrt((2l+1)/(4pi) (l-m)!/(l+m)!),
// laid out at index l*l+l+m, treated as polynomials in independent x,y,z.  Here
// the Legendre-derivative coefficients are generated on the host in double and
// kept in __constant__ memory; a thread evaluates one point with Horner + the
// (x+iy)^m recurrence, which also yields the exact partials
//     d/dx = N Q m A_{m-1} ... , d/dz = N Q_l^{m+1} A_m.
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace {

// ---------------------------------------------------------------- frequency
__global__ void k_freq_fwd(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t C, float* __restrict__ outputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * C) return;
    const uint32_t b = t / C, c = t - b * C;
    if (c < D) { outputs[t] = inputs[(size_t)b * D + c]; return; }
    const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
    const float phase = (float)(col % 2) * (3.141592653589793f / 2);
    outputs[t] = __sinf(scalbnf(inputs[(size_t)b * D + d], (int)freq) + phase);     // cos as phase-shifted fast sine, as the reference
}

__global__ void k_freq_bwd(const float* __restrict__ grad, const float* __restrict__ outputs, uint32_t B, uint32_t D, uint32_t deg,
                           uint32_t C, float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float result = g[d];
    g += D; o += D;
    for (uint32_t f = 0; f < deg; f++) {
        result += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
        g += 2 * D; o += 2 * D;
    }
    grad_inputs[t] = result;
}

// ---------------------------------------------------------------- spherical harmonics
constexpr int kMaxDeg = 8;
struct ShTable {
    float q[kMaxDeg][kMaxDeg + 1][kMaxDeg];   // q[l][m][k]: coefficient of z^k in Q_l^m (m = l+1 -> all zero)
    float n[kMaxDeg][kMaxDeg];                // signed normalisation incl. sqrt2 and (-1)^m
};
__constant__ ShTable c_sh;

void build_sh_table(ShTable* t) {
    memset(t, 0, sizeof *t);
    double P[kMaxDeg][kMaxDeg] = {{0}};
    P[0][0] = 1.0;
    if (kMaxDeg > 1) P[1][1] = 1.0;
    for (int n = 1; n + 1 < kMaxDeg; n++) {            // (n+1) P_{n+1} = (2n+1) z P_n - n P_{n-1}
        for (int k = 0; k <= n; k++) P[n + 1][k + 1] += (2.0 * n + 1) * P[n][k] / (n + 1);
        for (int k = 0; k <= n - 1; k++) P[n + 1][k] -= (double)n * P[n - 1][k] / (n + 1);
    }
    for (int l = 0; l < kMaxDeg; l++) {
        double q[kMaxDeg];
        for (int k = 0; k < kMaxDeg; k++) q[k] = P[l][k];
        for (int m = 0; m <= l + 1 && m <= kMaxDeg; m++) {
            for (int k = 0; k < kMaxDeg; k++) t->q[l][m][k] = (float)q[k];
            if (m <= l && m < kMaxDeg) {
                double f = 1.0;
                for (int i = l - m + 1; i <= l + m; i++) f *= i;        // (l+m)!/(l-m)!
                double N = sqrt((2.0 * l + 1) / (4.0 * M_PI) / f);
                if (m) N *= sqrt(2.0);
                if (m & 1) N = -N;
                t->n[l][m] = (float)N;
            }
            for (int k = 0; k + 1 < kMaxDeg; k++) q[k] = q[k + 1] * (k + 1);   // differentiate
            q[kMaxDeg - 1] = 0.0;
        }
    }
}

__device__ __forceinline__ float horner(const float* c, int deg, float z) {
    float r = 0.f;
    for (int k = deg; k >= 0; k--) r = fmaf(r, z, c[k]);
    return r;
}

__global__ void __launch_bounds__(128) k_sh_fwd(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B, uint32_t D,
                                                uint32_t degree, float* __restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t C2 = degree * degree;
    const float x = inputs[(size_t)b * D], y = inputs[(size_t)b * D + 1], z = inputs[(size_t)b * D + 2];
    float A[kMaxDeg], Bm[kMaxDeg];
    A[0] = 1.f; Bm[0] = 0.f;
#pragma unroll
    for (int m = 1; m < kMaxDeg; m++) { A[m] = A[m-1] * x - Bm[m-1] * y; Bm[m] = A[m-1] * y + Bm[m-1] * x; }
    float* out = outputs + (size_t)b * C2;
    float* ddx = dy_dx ? dy_dx + (size_t)b * D * C2 : nullptr;
    float* ddy = ddx ? ddx + C2 : nullptr;
    float* ddz = ddx ? ddy + C2 : nullptr;
    for (int l = 0; l < (int)degree; l++) {
#pragma unroll
        for (int m = 0; m < kMaxDeg; m++) {
            if (m > l) break;
            const float N = c_sh.n[l][m];
            const float Q = horner(c_sh.q[l][m], l - m, z);
            const int ip = l * l + l + m, in_ = l * l + l - m;
            const float NQ = N * Q;
            out[ip] = NQ * A[m];
            if (m) out[in_] = NQ * Bm[m];
            if (ddx) {
                const float Q1 = (l - m - 1 >= 0) ? horner(c_sh.q[l][m + 1], l - m - 1, z) : 0.f;
                const float NQ1 = N * Q1;
                const float fm = (float)m;
                const float Am1 = m ? A[m - 1] : 0.f, Bm1 = m ? Bm[m - 1] : 0.f;
                ddx[ip] = NQ * fm * Am1; ddy[ip] = -NQ * fm * Bm1; ddz[ip] = NQ1 * A[m];
                if (m) { ddx[in_] = NQ * fm * Bm1; ddy[in_] = NQ * fm * Am1; ddz[in_] = NQ1 * Bm[m]; }
            }
        }
    }
}

// grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]     (accumulating, shencoder.cu:377-380)
__global__ void k_sh_bwd(const float* __restrict__ grad, uint32_t B, uint32_t D, uint32_t degree, const float* __restrict__ dy_dx,
                         float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = t / D;
    if (b >= B) return;
    const uint32_t d = t - b * D, C2 = degree * degree;
    const float* g = grad + (size_t)b * C2;
    const float* dd = dy_dx + (size_t)b * D * C2 + (size_t)d * C2;
    float acc = grad_inputs[t];
    for (uint32_t ch = 0; ch < C2; ch++) acc += g[ch] * dd[ch];
    grad_inputs[t] = acc;
}

bool g_sh_ready[64] = {false};

}  // namespace

SDF_API int sdf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs, void* stream) {
    if (B == 0) return SDF_OK;
    SDF_CHECK_ARG(inputs && outputs, "freq_encode_forward: null pointer");
    SDF_CHECK_ARG(C == D + D * 2 * deg, "freq_encode_forward: output_dim must be D + 2*D*degree");
    k_freq_fwd<<<cdiv(B * C, 256), 256, 0, (cudaStream_t)stream>>>(inputs, B, D, C, outputs);
    SDF_CHECK_LAUNCH("freq_encode_forward");
    return SDF_OK;
}

SDF_API int sdf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                     float* grad_inputs, void* stream) {
    if (B == 0) return SDF_OK;
    SDF_CHECK_ARG(grad && outputs && grad_inputs, "freq_encode_backward: null pointer");
    SDF_CHECK_ARG(C == D + D * 2 * deg, "freq_encode_backward: output_dim must be D + 2*D*degree");
    k_freq_bwd<<<cdiv(B * D, 256), 256, 0, (cudaStream_t)stream>>>(grad, outputs, B, D, deg, C, grad_inputs);
    SDF_CHECK_LAUNCH("freq_encode_backward");
    return SDF_OK;
}

SDF_API int sdf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx, void* stream) {
    if (B == 0) return SDF_OK;
    SDF_CHECK_ARG(inputs && outputs, "sh_encode_forward: null pointer");
    SDF_CHECK_ARG(D == 3, "SH encoder only supports input dim == 3");
    SDF_CHECK_ARG(degree >= 1 && degree <= (uint32_t)kMaxDeg, "SH encoder only supports degree in [1, 8]");
    int dev = 0;
    SDF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !g_sh_ready[dev]) {
        static ShTable host_table;
        build_sh_table(&host_table);
        SDF_CHECK_CUDA(cudaMemcpyToSymbol(c_sh, &host_table, sizeof(ShTable)));
        g_sh_ready[dev] = true;
    }
    k_sh_fwd<<<cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(inputs, outputs, B, D, degree, dy_dx);
    SDF_CHECK_LAUNCH("sh_encode_forward");
    return SDF_OK;
}

SDF_API int sdf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                                   float* grad_inputs, void* stream) {
    (void)inputs;
    if (B == 0) return SDF_OK;
    SDF_CHECK_ARG(grad && dy_dx && grad_inputs, "sh_encode_backward: null pointer");
    SDF_CHECK_ARG(D == 3 && degree >= 1 && degree <= (uint32_t)kMaxDeg, "sh_encode_backward: bad D/degree");
    k_sh_bwd<<<cdiv(B * D, 256), 256, 0, (cudaStream_t)stream>>>(grad, B, D, degree, dy_dx, grad_inputs);
    SDF_CHECK_LAUNCH("sh_encode_backward");
    return SDF_OK;
}
