// raymarch.cu — occupancy-grid ray marching and volume compositing for sm_100a.
//
// Replaces the reference's raymarching/src/raymarching.cu (thread-per-ray, two
// host-synchronised passes, atomic offsets).  Design here:
//   * warp-per-ray marching.  The reference's t-recurrence
//       t_{k+1} = t_k + clamp(t_k*dt_gamma, dt_min, dt_max)      (raymarching.cu:396-465)
//     does not depend on the grid: occupied steps and the empty-voxel
//     "do t += dt while t < tt" loop walk the SAME chain.  A warp therefore
//     evaluates 32 consecutive chain points at once (independent bitfield
//     loads -> memory-level parallelism), then resolves which of them the
//     sequential algorithm would have *visited* with ballots.  The fp32 chain is
//     advanced with the reference's exact operation order, so counts, xyzs and
//     ts are bit-identical to the reference kernel.
//   * deterministic sample offsets: exclusive prefix sum over rays instead of
//     atomicAdd order (raymarching.cu:470-474), computed on the device; M is
//     left in device memory (and mirrored to a pinned host word) so a fused
//     caller never has to synchronise.
//   * warp-per-ray compositing with shuffle scans, coalesced sample reads.
#include "common.cuh"
#include <float.h>

namespace {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__device__ __forceinline__ uint32_t compact_bits(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
// frexp exponent clamped to [0, C-1]   (raymarching.cu:42-54)
__device__ __forceinline__ int mip_level(float mx, float Cf) {
    int e;
    frexpf(mx, &e);
    return (int)fminf(Cf - 1.0f, fmaxf(0.0f, (float)e));
}

// ---------------------------------------------------------------- small utils
__global__ void k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                           const float* __restrict__ aabb, uint32_t N, float min_near,
                           float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n*3], oy = rays_o[n*3+1], oz = rays_o[n*3+2];
    const float rdx = __frcp_rn(rays_d[n*3]), rdy = __frcp_rn(rays_d[n*3+1]), rdz = __frcp_rn(rays_d[n*3+2]);
    float near = __fmul_rn(aabb[0] - ox, rdx), far = __fmul_rn(aabb[3] - ox, rdx), t;
    if (near > far) { t = near; near = far; far = t; }
    float ny = __fmul_rn(aabb[1] - oy, rdy), fy = __fmul_rn(aabb[4] - oy, rdy);
    if (ny > fy) { t = ny; ny = fy; fy = t; }
    bool miss = (near > fy || ny > far);
    if (!miss) {
        if (ny > near) near = ny;
        if (fy < far) far = fy;
        float nz = __fmul_rn(aabb[2] - oz, rdz), fz = __fmul_rn(aabb[5] - oz, rdz);
        if (nz > fz) { t = nz; nz = fz; fz = t; }
        miss = (near > fz || nz > far);
        if (!miss) {
            if (nz > near) near = nz;
            if (fz < far) far = fz;
            if (near < min_near) near = min_near;
        }
    }
    if (miss) near = far = FLT_MAX;
    nears[n] = near;
    fars[n] = far;
}

__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                               uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float RPI = 0.3183098861837907f;
    const float ox = rays_o[n*3], oy = rays_o[n*3+1], oz = rays_o[n*3+2];
    const float dx = rays_d[n*3], dy = rays_d[n*3+1], dz = rays_d[n*3+2];
    const float A = dx*dx + dy*dy + dz*dz;
    const float B = ox*dx + oy*dy + oz*dz;
    const float C = ox*ox + oy*oy + oz*oz - radius*radius;
    const float t = (-B + sqrtf(B*B - A*C)) / A;
    const float x = ox + t*dx, y = oy + t*dy, z = oz + t*dz;
    const float theta = atan2f(sqrtf(x*x + z*z), y);
    const float phi = atan2f(z, x);
    coords[n*2] = 2 * theta * RPI - 1;
    coords[n*2+1] = phi * RPI;
}

__global__ void k_morton3D(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int)morton3d((uint32_t)coords[n*3], (uint32_t)coords[n*3+1], (uint32_t)coords[n*3+2]);
}
__global__ void k_morton3D_invert(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[n*3] = (int)compact_bits((uint32_t)(ind >> 0));
    coords[n*3+1] = (int)compact_bits((uint32_t)(ind >> 1));
    coords[n*3+2] = (int)compact_bits((uint32_t)(ind >> 2));
}

// 8 floats -> 1 byte; one thread reads two float4 (32 B) so a warp reads 1 KB contiguous.
__global__ void k_packbits(const float4* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float4 a = grid[(size_t)n*2], b = grid[(size_t)n*2+1];
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// warp per ray: coalesced fill of res[offset .. offset+count) = n
__global__ void k_flatten_rays(const int* __restrict__ rays, uint32_t N, uint32_t M, int* __restrict__ res) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
    for (uint32_t i = lane; i < cnt; i += 32)
        if (off + i < M) res[off + i] = (int)n;
}

// ---------------------------------------------------------------- marching
struct MarchParams {
    float bound, dt_gamma, dt_min, dt_max, rH, H3, Hf, Hm1f, Cf;
    uint32_t H;
    int contract;
};

static MarchParams make_march_params(float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    MarchParams p;
    p.bound = bound; p.dt_gamma = dt_gamma; p.contract = contract; p.H = H;
    p.Hf = (float)H; p.Hm1f = (float)(H - 1); p.Cf = (float)C;
    p.rH = 1.0f / (float)H;
    p.H3 = (float)(H * H * H);
    p.dt_min = (2.0f * 1.7320508075688772f) / (float)max_steps;      // host IEEE fp32 division == device div.rn
    p.dt_max = ((2.0f * 1.7320508075688772f) * bound) / (float)H;
    return p;
}

// One warp marches one ray.  Emits at most max_emit samples starting at chain
// value t0; returns the number emitted (uniform across the warp).
// WRITE=false: count only.  Sample s goes to xyzs[s], dirs[s], ts[s]
// (pointers already offset to this ray's first slot).
template <bool WRITE>
__device__ __forceinline__ uint32_t march_ray_warp(const MarchParams& p, const uint8_t* __restrict__ grid,
                                                   float ox, float oy, float oz, float dx, float dy, float dz,
                                                   float t0, float far, uint32_t max_emit,
                                                   float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ ts) {
    const uint32_t lane = threadIdx.x & 31;
    const float rdx = __frcp_rn(dx), rdy = __frcp_rn(dy), rdz = __frcp_rn(dz);
    const float sx = copysignf(1.0f, dx), sy = copysignf(1.0f, dy), sz = copysignf(1.0f, dz);
    uint32_t count = 0;
    float t_base = t0;
    float skip_until = -INFINITY;   // chain points with t < skip_until were jumped over by an empty-voxel skip
    bool done = !(t0 < far) || max_emit == 0;

    while (!done) {
        // --- 32 consecutive chain points; lane L keeps (t_{base+L}, dt_{base+L})
        float t = t_base, my_t = t_base, my_dt = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float dtj = clampf(__fmul_rn(t, p.dt_gamma), p.dt_min, p.dt_max);
            if ((int)lane == j) { my_t = t; my_dt = dtj; }
            t = __fadd_rn(t, dtj);
        }
        t_base = t;

        const bool valid = my_t < far;
        bool occ = false;
        float tt = -INFINITY, cx = 0.f, cy = 0.f, cz = 0.f;
        if (valid) {
            const float x = clampf(__fmaf_rn(my_t, dx, ox), -p.bound, p.bound);
            const float y = clampf(__fmaf_rn(my_t, dy, oy), -p.bound, p.bound);
            const float z = clampf(__fmaf_rn(my_t, dz, oz), -p.bound, p.bound);
            const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
            const int level = max(mip_level(mag, p.Cf), mip_level(__fmul_rn(__fmul_rn(my_dt, p.Hf), 0.5f), p.Cf));
            const float mip_bound = fminf(scalbnf(1.0f, level), p.bound);
            const float mip_rbound = __frcp_rn(mip_bound);
            cx = x; cy = y; cz = z;
            const bool contracted = p.contract && mag > 1.0f;
            if (contracted) {
                const float s = __fdiv_rn(2.0f - __fdiv_rn(1.0f, mag), mag);
                cx = __fmul_rn(cx, s); cy = __fmul_rn(cy, s); cz = __fmul_rn(cz, s);
            }
            // 0.5 * (c * rbound + 1) * H : the reference evaluates the products in double; both are exact scalings
            const int nx = (int)clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(cx, mip_rbound, 1.0f)), p.Hf), 0.0f, p.Hm1f);
            const int ny = (int)clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(cy, mip_rbound, 1.0f)), p.Hf), 0.0f, p.Hm1f);
            const int nz = (int)clampf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(cz, mip_rbound, 1.0f)), p.Hf), 0.0f, p.Hm1f);
            const uint32_t index = (uint32_t)__fmaf_rn((float)level, p.H3, (float)morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
            occ = (__ldg(grid + (index >> 3)) >> (index & 7)) & 1;
            if (!occ && !contracted) {
                const float tx = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(__fmaf_rn(0.5f, sx, __fadd_rn((float)nx, 0.5f)), p.rH), 2.0f, -1.0f), -cx), rdx);
                const float ty = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(__fmaf_rn(0.5f, sy, __fadd_rn((float)ny, 0.5f)), p.rH), 2.0f, -1.0f), -cy), rdy);
                const float tz = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(__fmaf_rn(0.5f, sz, __fadd_rn((float)nz, 0.5f)), p.rH), 2.0f, -1.0f), -cz), rdz);
                tt = __fadd_rn(my_t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
            }
        }
        const uint32_t occ_mask = __ballot_sync(0xffffffffu, occ && valid);
        const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);

        // --- resolve which chain points the sequential march visits
        uint32_t cur = 0;
        while (cur < 32) {
            const uint32_t ok = __ballot_sync(0xffffffffu, my_t >= skip_until) & (0xffffffffu << cur);
            if (ok == 0) break;                           // rest of the chunk was skipped over; carry skip_until
            cur = __ffs(ok) - 1;
            if (!((valid_mask >> cur) & 1u)) { done = true; break; }      // t >= far
            const uint32_t rest = occ_mask >> cur;
            const uint32_t run = (rest == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~rest) - 1);   // consecutive occupied points from cur
            if (run > 0) {
                const uint32_t take = min(run, max_emit - count);
                if (WRITE) {
                    if (lane >= cur && lane < cur + take) {
                        const uint32_t s = count + (lane - cur);
                        xyzs[s*3] = cx; xyzs[s*3+1] = cy; xyzs[s*3+2] = cz;
                        dirs[s*3] = dx; dirs[s*3+1] = dy; dirs[s*3+2] = dz;
                        ts[s*2] = __fadd_rn(my_t, my_dt); ts[s*2+1] = my_dt;
                    }
                }
                count += take;
                if (count >= max_emit) { done = true; break; }
                cur += run;
                skip_until = -INFINITY;
                if (cur >= 32) break;
            } else {
                // visited, empty: jump to the first chain point at or beyond the voxel exit
                skip_until = __shfl_sync(0xffffffffu, tt, cur);
                cur += 1;
            }
        }
        if (!(t_base < far) && !done) {
            // every remaining chain point is beyond far
            done = true;
        }
    }
    return count;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) k_march_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                     const uint8_t* __restrict__ grid, MarchParams p, uint32_t max_steps, uint32_t N,
                                                     const float* __restrict__ nears, const float* __restrict__ fars,
                                                     const float* __restrict__ noises,
                                                     float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ ts,
                                                     int* __restrict__ rays, uint32_t capacity) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const uint32_t lane = threadIdx.x & 31;
    const float ox = rays_o[n*3], oy = rays_o[n*3+1], oz = rays_o[n*3+2];
    const float dx = rays_d[n*3], dy = rays_d[n*3+1], dz = rays_d[n*3+2];
    const float near = nears[n], far = fars[n];
    const float noise = noises ? noises[n] : 0.0f;
    const float t0 = __fmaf_rn(clampf(__fmul_rn(near, p.dt_gamma), p.dt_min, p.dt_max), noise, near);
    if (!WRITE) {
        const uint32_t c = march_ray_warp<false>(p, grid, ox, oy, oz, dx, dy, dz, t0, far, max_steps, nullptr, nullptr, nullptr);
        if (lane == 0) rays[n*2+1] = (int)c;
    } else {
        const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
        if (cnt == 0 || off + cnt > capacity) return;
        march_ray_warp<true>(p, grid, ox, oy, oz, dx, dy, dz, t0, far, cnt,
                             xyzs + (size_t)off*3, dirs + (size_t)off*3, ts + (size_t)off*2);
    }
}

// Exclusive scan of rays[:,1] into rays[:,0]; total -> counter[0] (and *host_mirror if given).
// One block; N is the ray count of a step (<= a few hundred thousand).
__global__ void __launch_bounds__(1024) k_scan_rays(int* __restrict__ rays, uint32_t N, int* __restrict__ counter,
                                                    volatile int* host_mirror) {
    __shared__ uint32_t warp_excl[32];
    __shared__ uint32_t tile_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t carry = 0;   // identical in every thread
    for (uint32_t base = 0; base < N; base += 1024) {
        const uint32_t i = base + tid;
        const uint32_t v = i < N ? (uint32_t)rays[i*2+1] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
        if (lane == 31) warp_excl[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = warp_excl[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o); if ((int)lane >= o) wi += u; }
            warp_excl[lane] = wi - w;
            if (lane == 31) tile_total = wi;
        }
        __syncthreads();
        if (i < N) rays[i*2] = (int)(carry + warp_excl[wid] + (incl - v));
        carry += tile_total;
        __syncthreads();
    }
    if (tid == 0) {
        counter[0] = (int)carry;
        if (host_mirror) *host_mirror = (int)carry;
    }
}

// ---------------------------------------------------------------- compositing (train)
// warp per ray, 32 samples per iteration.  Writes weights for every sample of
// the ray (0 past early termination), so the caller need not pre-zero.
__global__ void __launch_bounds__(256) k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                             const float* __restrict__ ts, const int* __restrict__ rays,
                                                             uint32_t M, uint32_t N, float T_thresh, int binarize,
                                                             float* __restrict__ weights, float* __restrict__ weights_sum,
                                                             float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
    float r = 0.f, g = 0.f, b = 0.f, ws = 0.f, d = 0.f;
    if (cnt != 0 && off + cnt <= M) {
        float T_carry = 1.0f;
        bool alive = true;
        for (uint32_t base = 0; base < cnt; base += 32) {
            const uint32_t s = base + lane;
            const bool act = s < cnt;
            const size_t i = (size_t)off + s;
            float w = 0.f;
            if (alive) {
                float sigma = 0.f, dt = 0.f, tv = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
                if (act) {
                    sigma = sigmas[i];
                    const float2 t2 = *reinterpret_cast<const float2*>(ts + i*2);
                    tv = t2.x; dt = t2.y;
                    cr = rgbs[i*3]; cg = rgbs[i*3+1]; cb = rgbs[i*3+2];
                }
                const float real_alpha = 1.0f - __expf(-sigma * dt);
                float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
                if (!act) alpha = 0.f;
                float P = 1.0f - alpha;          // inclusive prefix product over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, P, o); if ((int)lane >= o) P *= u; }
                float Pex = __shfl_up_sync(0xffffffffu, P, 1);
                if (lane == 0) Pex = 1.0f;
                const float T_ex = T_carry * Pex;
                // the reference stops after the first sample whose trailing T < T_thresh
                const bool proc = act && (s == 0 || T_ex >= T_thresh);
                w = proc ? alpha * T_ex : 0.f;
                r = fmaf(w, cr, r); g = fmaf(w, cg, g); b = fmaf(w, cb, b);
                ws += w; d = fmaf(w, tv, d);
                T_carry = T_carry * __shfl_sync(0xffffffffu, P, 31);
                if (T_carry < T_thresh) alive = false;
            }
            if (act) weights[i] = w;
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    }
    if (lane == 0) {
        weights_sum[n] = ws; depth[n] = d;
        image[n*3] = r; image[n*3+1] = g; image[n*3+2] = b;
    }
}

__device__ __forceinline__ float warp_incl_sum(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += u; }
    return v;
}

// raymarching.cu:606-695 restructured: per sample
//   grad_sigma = dt * ( T*(gi.rgb + gd*t) - (F - A) + (gws + gw_i) * (T - (WS - ws)) )
// with A = running sum of w*(gi.rgb + gd*t), F its final value (= gi.image + gd*depth of the forward),
// ws the running weight sum, T the transmittance after the sample.
__global__ void __launch_bounds__(256) k_composite_train_bwd(const float* __restrict__ grad_weights, const float* __restrict__ grad_weights_sum,
                                                             const float* __restrict__ grad_depth, const float* __restrict__ grad_image,
                                                             const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                             const float* __restrict__ ts, const int* __restrict__ rays,
                                                             const float* __restrict__ weights_sum, const float* __restrict__ depth,
                                                             const float* __restrict__ image,
                                                             uint32_t M, uint32_t N, float T_thresh, int binarize,
                                                             float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
    if (cnt == 0 || off + cnt > M) return;
    const float gi0 = grad_image[n*3], gi1 = grad_image[n*3+1], gi2 = grad_image[n*3+2];
    const float gws = grad_weights_sum ? grad_weights_sum[n] : 0.f;
    const float gd = grad_depth ? grad_depth[n] : 0.f;
    const float F = gi0 * image[n*3] + gi1 * image[n*3+1] + gi2 * image[n*3+2] + gd * depth[n];
    const float WS = weights_sum[n];
    float T_carry = 1.0f, A_carry = 0.f, ws_carry = 0.f;
    bool alive = true;
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t s = base + lane;
        const bool act = s < cnt;
        const size_t i = (size_t)off + s;
        float gs = 0.f, gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
        if (alive) {
            float sigma = 0.f, dt = 0.f, tv = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, gw = 0.f;
            if (act) {
                sigma = sigmas[i];
                const float2 t2 = *reinterpret_cast<const float2*>(ts + i*2);
                tv = t2.x; dt = t2.y;
                cr = rgbs[i*3]; cg = rgbs[i*3+1]; cb = rgbs[i*3+2];
                if (grad_weights) gw = grad_weights[i];
            }
            const float real_alpha = 1.0f - __expf(-sigma * dt);
            float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            if (!act) alpha = 0.f;
            float P = 1.0f - alpha;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, P, o); if ((int)lane >= o) P *= u; }
            float Pex = __shfl_up_sync(0xffffffffu, P, 1);
            if (lane == 0) Pex = 1.0f;
            const float T_ex = T_carry * Pex, T_in = T_carry * P;
            const bool proc = act && (s == 0 || T_ex >= T_thresh);
            const float w = proc ? alpha * T_ex : 0.f;
            const float q = gi0 * cr + gi1 * cg + gi2 * cb + gd * tv;
            const float A = A_carry + warp_incl_sum(w * q, lane);
            const float wsi = ws_carry + warp_incl_sum(w, lane);
            if (proc) {
                gs = dt * ((T_in * q - (F - A)) + (gws + gw) * (T_in - (WS - wsi)));
                gr0 = gi0 * w; gr1 = gi1 * w; gr2 = gi2 * w;
            }
            A_carry = __shfl_sync(0xffffffffu, A, 31);
            ws_carry = __shfl_sync(0xffffffffu, wsi, 31);
            T_carry = __shfl_sync(0xffffffffu, T_in, 31);
            if (T_carry < T_thresh) alive = false;
        }
        if (act) {
            grad_sigmas[i] = gs;
            grad_rgbs[i*3] = gr0; grad_rgbs[i*3+1] = gr1; grad_rgbs[i*3+2] = gr2;
        }
    }
}

// ---------------------------------------------------------------- inference
__global__ void __launch_bounds__(256) k_march_infer(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
                                                     const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                                                     const float* __restrict__ rays_d, MarchParams p, const uint8_t* __restrict__ grid,
                                                     const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                                                     float* __restrict__ ts, const float* __restrict__ noises) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float ox = rays_o[index*3], oy = rays_o[index*3+1], oz = rays_o[index*3+2];
    const float dx = rays_d[index*3], dy = rays_d[index*3+1], dz = rays_d[index*3+2];
    const float noise = noises ? noises[n] : 0.f;
    float t = rays_t[index];
    t = __fmaf_rn(clampf(__fmul_rn(t, p.dt_gamma), p.dt_min, p.dt_max), noise, t);
    march_ray_warp<true>(p, grid, ox, oy, oz, dx, dy, dz, t, fars[index], n_step,
                         xyzs + (size_t)n*n_step*3, dirs + (size_t)n*n_step*3, ts + (size_t)n*n_step*2);
}

// raymarching.cu:843-925: at most n_step (<= 8 in the renderer) samples per ray, kept sequential.
__global__ void k_composite_infer(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize,
                                  int* __restrict__ rays_alive, float* __restrict__ rays_t,
                                  const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ts,
                                  float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    sigmas += (size_t)n*n_step; rgbs += (size_t)n*n_step*3; ts += (size_t)n*n_step*2;
    float t = 0.f, d = depth[index], r = image[index*3], g = image[index*3+1], b = image[index*3+2], wsum = weights_sum[index];
    uint32_t step = 0;
    while (step < n_step) {
        if (ts[0] == 0.f) break;
        const float real_alpha = 1.0f - __expf(-sigmas[0] * ts[1]);
        const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
        const float T = 1.0f - wsum;
        const float w = alpha * T;
        wsum += w;
        t = ts[0];
        d = fmaf(w, t, d); r = fmaf(w, rgbs[0], r); g = fmaf(w, rgbs[1], g); b = fmaf(w, rgbs[2], b);
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; ts += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
    weights_sum[index] = wsum; depth[index] = d;
    image[index*3] = r; image[index*3+1] = g; image[index*3+2] = b;
}


// ---------------------------------------------------------------- inference loop with its bookkeeping on the device
// nerf/renderer.py:759-794 re-reads the alive count on the host every iteration (boolean indexing = a sync + a compaction through
// three PyTorch kernels).  Here the loop state lives in device memory: every kernel of an iteration is launched with capacity N and
// reads (n_alive, n_step) itself; compaction is a warp-aggregated scatter whose last block advances the state.  The host only
// enqueues and polls a pinned mirror of n_alive (never blocks on it) to stop enqueueing once the frame is done.
struct InferState {
    int n_alive;        // rays still marching
    int n_step;         // samples per ray this iteration = clamp(N / n_alive, 1, 8)   (renderer.py:771)
    int M;              // n_alive * n_step: live rows of xyzs / sigmas this iteration (the fused field's m_dev)
    int step;           // sum of n_step so far; the loop ends at max_steps (renderer.py:763)
    int next_alive;     // compaction cursor
    unsigned ticket;    // blocks done in the compaction kernel
    int N, max_steps;
};

__device__ __forceinline__ void infer_set_counts(InferState* s, int n_alive) {
    if (s->step >= s->max_steps) n_alive = 0;
    s->n_alive = n_alive;
    const int n_step = n_alive > 0 ? max(min(s->N / n_alive, 8), 1) : 0;
    s->n_step = n_step;
    s->M = n_alive * n_step;
}

__global__ void k_infer_begin(InferState* __restrict__ st, uint32_t N, uint32_t max_steps, int* __restrict__ rays_alive,
                              float* __restrict__ rays_t, const float* __restrict__ nears, float* __restrict__ weights_sum,
                              float* __restrict__ depth, float* __restrict__ image, volatile int* host_alive) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) {
        st->N = (int)N; st->max_steps = (int)max_steps; st->step = 0; st->next_alive = 0; st->ticket = 0;
        infer_set_counts(st, (int)N);
        if (host_alive) *host_alive = (int)N;
    }
    if (n >= N) return;
    rays_alive[n] = (int)n;
    rays_t[n] = nears[n];
    weights_sum[n] = 0.f; depth[n] = 0.f;
    image[n*3] = 0.f; image[n*3+1] = 0.f; image[n*3+2] = 0.f;
}

__global__ void __launch_bounds__(256) k_march_infer_dev(const InferState* __restrict__ st, const int* __restrict__ rays_alive,
                                                         const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                                                         const float* __restrict__ rays_d, MarchParams p, const uint8_t* __restrict__ grid,
                                                         const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                                                         float* __restrict__ ts, const float* __restrict__ noises) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_alive = (uint32_t)st->n_alive, n_step = (uint32_t)st->n_step;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float ox = rays_o[index*3], oy = rays_o[index*3+1], oz = rays_o[index*3+2];
    const float dx = rays_d[index*3], dy = rays_d[index*3+1], dz = rays_d[index*3+2];
    const float noise = noises ? noises[n] : 0.f;
    float t = rays_t[index];
    t = __fmaf_rn(clampf(__fmul_rn(t, p.dt_gamma), p.dt_min, p.dt_max), noise, t);
    float* tsr = ts + (size_t)n*n_step*2;
    const uint32_t c = march_ray_warp<true>(p, grid, ox, oy, oz, dx, dy, dz, t, fars[index], n_step,
                                            xyzs + (size_t)n*n_step*3, dirs + (size_t)n*n_step*3, tsr);
    // unwritten tail slots: ts[0] == 0 is the compositor's end-of-ray sentinel (the reference relies on zero-initialised buffers);
    // their xyz must be finite for the field kernel
    for (uint32_t s = c + lane; s < n_step; s += 32) {
        tsr[s*2] = 0.f; tsr[s*2+1] = 0.f;
        float* x = xyzs + ((size_t)n*n_step + s)*3;
        x[0] = 0.f; x[1] = 0.f; x[2] = 0.f;
    }
}

__global__ void k_composite_infer_dev(const InferState* __restrict__ st, float T_thresh, int binarize,
                                      int* __restrict__ rays_alive, float* __restrict__ rays_t,
                                      const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ts,
                                      float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_alive = (uint32_t)st->n_alive, n_step = (uint32_t)st->n_step;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    sigmas += (size_t)n*n_step; rgbs += (size_t)n*n_step*3; ts += (size_t)n*n_step*2;
    float t = 0.f, d = depth[index], r = image[index*3], g = image[index*3+1], b = image[index*3+2], wsum = weights_sum[index];
    uint32_t step = 0;
    while (step < n_step) {
        if (ts[0] == 0.f) break;
        const float real_alpha = 1.0f - __expf(-sigmas[0] * ts[1]);
        const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
        const float T = 1.0f - wsum;
        const float w = alpha * T;
        wsum += w;
        t = ts[0];
        d = fmaf(w, t, d); r = fmaf(w, rgbs[0], r); g = fmaf(w, rgbs[1], g); b = fmaf(w, rgbs[2], b);
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; ts += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
    weights_sum[index] = wsum; depth[index] = d;
    image[index*3] = r; image[index*3+1] = g; image[index*3+2] = b;
}

// rays_out <- the entries of rays_in[0 .. n_alive) that are >= 0 (order within a warp kept, warps in arrival order); the last block
// to finish advances the loop state and mirrors the new count to pinned host memory
__global__ void __launch_bounds__(256) k_compact_alive(InferState* __restrict__ st, const int* __restrict__ rays_in, int* __restrict__ rays_out,
                                                       volatile int* host_alive) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const uint32_t n_alive = (uint32_t)st->n_alive;
    const int idx = n < n_alive ? rays_in[n] : -1;
    const bool keep = idx >= 0;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&st->next_alive, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) rays_out[base + __popc(m & ((1u << lane) - 1u))] = idx;
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        const int alive = atomicAdd(&st->next_alive, 0);
        st->step += st->n_step;
        st->next_alive = 0; st->ticket = 0;
        infer_set_counts(st, alive);
        if (host_alive) *host_alive = st->n_alive;
    }
}

}  // namespace

// ================================================================= C ABI
SDF_API int sdf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                                   float* nears, float* fars, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(rays_o && rays_d && aabb && nears && fars, "near_far_from_aabb: null pointer");
    k_near_far<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    SDF_CHECK_LAUNCH("near_far_from_aabb");
    return SDF_OK;
}

SDF_API int sdf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(rays_o && rays_d && coords, "sph_from_ray: null pointer");
    k_sph_from_ray<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    SDF_CHECK_LAUNCH("sph_from_ray");
    return SDF_OK;
}

SDF_API int sdf_morton3D(const int* coords, uint32_t N, int* indices, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(coords && indices, "morton3D: null pointer");
    k_morton3D<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    SDF_CHECK_LAUNCH("morton3D");
    return SDF_OK;
}

SDF_API int sdf_morton3D_invert(const int* indices, uint32_t N, int* coords, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(coords && indices, "morton3D_invert: null pointer");
    k_morton3D_invert<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    SDF_CHECK_LAUNCH("morton3D_invert");
    return SDF_OK;
}

SDF_API int sdf_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(grid && bitfield, "packbits: null pointer");
    SDF_CHECK_ARG(((uintptr_t)grid & 15) == 0, "packbits: grid must be 16-byte aligned");
    k_packbits<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(grid), N, thresh, bitfield);
    SDF_CHECK_LAUNCH("packbits");
    return SDF_OK;
}

SDF_API int sdf_flatten_rays(const int* rays, uint32_t N, uint32_t M, int* res, void* stream) {
    if (N == 0 || M == 0) return SDF_OK;
    SDF_CHECK_ARG(rays && res, "flatten_rays: null pointer");
    k_flatten_rays<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(rays, N, M, res);
    SDF_CHECK_LAUNCH("flatten_rays");
    return SDF_OK;
}

// Pass 1 of march_rays_train: per-ray sample counts, exclusive offsets, total.
//   rays    [N,2] int32 out: (offset, count), offsets in ray order
//   counter [1]   int32 out: M (device)
//   host_M  optional pinned-host int32 that receives M from the device (mapped write), may be NULL
SDF_API int sdf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, int contract,
                                       float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                       const float* nears, const float* fars, const float* noises,
                                       int* rays, int* counter, int* host_M, void* stream) {
    SDF_CHECK_ARG(counter, "march_rays_train_count: null counter");
    SDF_CHECK_ARG(N == 0 || (rays_o && rays_d && grid && nears && fars && rays), "march_rays_train_count: null pointer");
    SDF_CHECK_ARG(max_steps > 0 && H > 0 && C > 0 && H <= 1024, "march_rays_train_count: bad max_steps/H/C");
    SDF_CHECK_ARG((uint64_t)C * H * H * H <= (1ull << 24), "march_rays_train_count: C*H^3 must be <= 2^24 (fp32 index arithmetic of the reference)");
    cudaStream_t st = (cudaStream_t)stream;
    if (N) {
        const MarchParams p = make_march_params(bound, contract, dt_gamma, max_steps, C, H);
        k_march_train<false><<<cdiv(N, 8), 256, 0, st>>>(rays_o, rays_d, grid, p, max_steps, N, nears, fars, noises,
                                                        nullptr, nullptr, nullptr, rays, 0);
        SDF_CHECK_LAUNCH("march_rays_train(count)");
    }
    k_scan_rays<<<1, 1024, 0, st>>>(rays, N, counter, host_M);
    SDF_CHECK_LAUNCH("march_rays_train(scan)");
    return SDF_OK;
}

// Pass 2: write samples at the offsets in rays.  capacity = rows available in xyzs/dirs/ts;
// rays whose samples would not fit are skipped (reference: offset + count > M, raymarching.cu:521).
SDF_API int sdf_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, int contract,
                                       float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                       const float* nears, const float* fars, const float* noises,
                                       float* xyzs, float* dirs, float* ts, const int* rays, uint32_t capacity, void* stream) {
    if (N == 0 || capacity == 0) return SDF_OK;
    SDF_CHECK_ARG(rays_o && rays_d && grid && nears && fars && rays && xyzs && dirs && ts, "march_rays_train_write: null pointer");
    SDF_CHECK_ARG(max_steps > 0 && H > 0 && C > 0 && (uint64_t)C * H * H * H <= (1ull << 24), "march_rays_train_write: bad max_steps/H/C");
    const MarchParams p = make_march_params(bound, contract, dt_gamma, max_steps, C, H);
    k_march_train<true><<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, grid, p, max_steps, N, nears, fars, noises,
                                                                      xyzs, dirs, ts, const_cast<int*>(rays), capacity);
    SDF_CHECK_LAUNCH("march_rays_train(write)");
    return SDF_OK;
}

SDF_API int sdf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts, const int* rays,
                                             uint32_t M, uint32_t N, float T_thresh, int binarize,
                                             float* weights, float* weights_sum, float* depth, float* image, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(rays && weights_sum && depth && image, "composite_rays_train_forward: null pointer");
    SDF_CHECK_ARG(M == 0 || (sigmas && rgbs && ts && weights), "composite_rays_train_forward: null sample pointer");
    k_composite_train_fwd<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, ts, rays, M, N, T_thresh, binarize,
                                                                        weights, weights_sum, depth, image);
    SDF_CHECK_LAUNCH("composite_rays_train_forward");
    return SDF_OK;
}

SDF_API int sdf_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum, const float* grad_depth,
                                              const float* grad_image, const float* sigmas, const float* rgbs, const float* ts,
                                              const int* rays, const float* weights_sum, const float* depth, const float* image,
                                              uint32_t M, uint32_t N, float T_thresh, int binarize,
                                              float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0 || M == 0) return SDF_OK;
    SDF_CHECK_ARG(grad_image && sigmas && rgbs && ts && rays && weights_sum && depth && image && grad_sigmas && grad_rgbs,
                  "composite_rays_train_backward: null pointer");
    k_composite_train_bwd<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(grad_weights, grad_weights_sum, grad_depth, grad_image,
                                                                        sigmas, rgbs, ts, rays, weights_sum, depth, image,
                                                                        M, N, T_thresh, binarize, grad_sigmas, grad_rgbs);
    SDF_CHECK_LAUNCH("composite_rays_train_backward");
    return SDF_OK;
}

SDF_API int sdf_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t, const float* rays_o,
                           const float* rays_d, float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                           const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* ts,
                           const float* noises, void* stream) {
    (void)nears;
    if (n_alive == 0 || n_step == 0) return SDF_OK;
    SDF_CHECK_ARG(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && ts, "march_rays: null pointer");
    SDF_CHECK_ARG(max_steps > 0 && H > 0 && C > 0 && (uint64_t)C * H * H * H <= (1ull << 24), "march_rays: bad max_steps/H/C");
    const MarchParams p = make_march_params(bound, contract, dt_gamma, max_steps, C, H);
    k_march_infer<<<cdiv(n_alive, 8), 256, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, p, grid,
                                                                      fars, xyzs, dirs, ts, noises);
    SDF_CHECK_LAUNCH("march_rays");
    return SDF_OK;
}

SDF_API int sdf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize, int* rays_alive, float* rays_t,
                               const float* sigmas, const float* rgbs, const float* ts, float* weights_sum, float* depth,
                               float* image, void* stream) {
    if (n_alive == 0) return SDF_OK;
    SDF_CHECK_ARG(rays_alive && rays_t && sigmas && rgbs && ts && weights_sum && depth && image, "composite_rays: null pointer");
    k_composite_infer<<<cdiv(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, T_thresh, binarize, rays_alive, rays_t,
                                                                            sigmas, rgbs, ts, weights_sum, depth, image);
    SDF_CHECK_LAUNCH("composite_rays");
    return SDF_OK;
}

// ---- inference loop with device-side bookkeeping (state: 8 ints of device memory; host_alive: optional pinned int32 mirror)
SDF_API int sdf_infer_begin(void* state, uint32_t N, uint32_t max_steps, int* rays_alive, float* rays_t, const float* nears,
                            float* weights_sum, float* depth, float* image, int* host_alive, void* stream) {
    SDF_CHECK_ARG(state && rays_alive && rays_t && nears && weights_sum && depth && image, "infer_begin: null pointer");
    k_infer_begin<<<cdiv(N > 0 ? N : 1, 256), 256, 0, (cudaStream_t)stream>>>((InferState*)state, N, max_steps, rays_alive, rays_t, nears, weights_sum,
                                                                             depth, image, host_alive);
    SDF_CHECK_LAUNCH("infer_begin");
    return SDF_OK;
}

SDF_API int sdf_infer_march(const void* state, uint32_t N, const int* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d,
                            float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                            const float* fars, float* xyzs, float* dirs, float* ts, const float* noises, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(state && rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && ts, "infer_march: null pointer");
    SDF_CHECK_ARG(max_steps > 0 && H > 0 && C > 0 && (uint64_t)C * H * H * H <= (1ull << 24), "infer_march: bad max_steps/H/C");
    const MarchParams p = make_march_params(bound, contract, dt_gamma, max_steps, C, H);
    k_march_infer_dev<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>((const InferState*)state, rays_alive, rays_t, rays_o, rays_d, p, grid, fars, xyzs,
                                                                    dirs, ts, noises);
    SDF_CHECK_LAUNCH("infer_march");
    return SDF_OK;
}

SDF_API int sdf_infer_composite(const void* state, uint32_t N, float T_thresh, int binarize, int* rays_alive, float* rays_t, const float* sigmas,
                                const float* rgbs, const float* ts, float* weights_sum, float* depth, float* image, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(state && rays_alive && rays_t && sigmas && rgbs && ts && weights_sum && depth && image, "infer_composite: null pointer");
    k_composite_infer_dev<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>((const InferState*)state, T_thresh, binarize, rays_alive, rays_t, sigmas, rgbs,
                                                                          ts, weights_sum, depth, image);
    SDF_CHECK_LAUNCH("infer_composite");
    return SDF_OK;
}

SDF_API int sdf_infer_compact(void* state, uint32_t N, const int* rays_in, int* rays_out, int* host_alive, void* stream) {
    if (N == 0) return SDF_OK;
    SDF_CHECK_ARG(state && rays_in && rays_out && rays_in != rays_out, "infer_compact: bad arguments");
    k_compact_alive<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>((InferState*)state, rays_in, rays_out, host_alive);
    SDF_CHECK_LAUNCH("infer_compact");
    return SDF_OK;
}
