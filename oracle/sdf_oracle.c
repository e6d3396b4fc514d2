/*
 * sdf_oracle.c — CPU restatement of the reference's native operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package under
 * stable-dreamfusion_b200/) may include, link or call this file.  It is used
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker.  Parity status: pinned against the reference's own CUDA kernels
 * (oracle/_ref, built by oracle/build_ref.py) on the GPU box; the outputs of
 * that run are committed as fixtures under tests/golden/ (see
 * tests/golden/make_golden_gpu.py) and re-checked on CPU by
 * tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to /root/reference).  Arithmetic is fp32 in the association order of the
 * reference kernels; where nvcc contracts a*b+c into one FMA in the reference
 * build (verified on the SASS of oracle/_ref/_raymarching.so) the restatement
 * calls fmaf() explicitly and this file is compiled with -ffp-contract=off,
 * so integer outputs (sample counts, indices) agree bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o libsdf_oracle.so sdf_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf_(float x) { return copysignf(1.0f, x); }

/* ------------------------------------------------------------------ */
/* raymarching/src/raymarching.cu:42-81 — mip level + Morton helpers   */
/* ------------------------------------------------------------------ */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    /* raymarching.cu:42-47 */
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e; frexpf(mx, &e);
    return (int)fminf(max_cascade - 1.0f, fmaxf(0.0f, (float)e));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    /* raymarching.cu:49-54: dt*H in fp32, *0.5 exact */
    const float mx = (dt * H) * 0.5f;
    int e; frexpf(mx, &e);
    return (int)fminf(max_cascade - 1.0f, fmaxf(0.0f, (float)e));
}
static inline uint32_t expand_bits(uint32_t v) {
    /* raymarching.cu:56-63 */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3d_invert1(uint32_t x) {
    /* raymarching.cu:73-81 */
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:92-145 */
ORACLE_API void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                                          uint32_t N, float min_near, float* nears, float* fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n*3], oy = rays_o[n*3+1], oz = rays_o[n*3+2];
        const float rdx = 1.0f / rays_d[n*3], rdy = 1.0f / rays_d[n*3+1], rdz = 1.0f / rays_d[n*3+2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, t;
        if (near > far) { t = near; near = far; far = t; }
        float ny = (aabb[1] - oy) * rdy, fy = (aabb[4] - oy) * rdy;
        if (ny > fy) { t = ny; ny = fy; fy = t; }
        if (near > fy || ny > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (ny > near) near = ny;
        if (fy < far) far = fy;
        float nz = (aabb[2] - oz) * rdz, fz = (aabb[5] - oz) * rdz;
        if (nz > fz) { t = nz; nz = fz; fz = t; }
        if (near > fz || nz > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (nz > near) near = nz;
        if (fz < far) far = fz;
        if (near < min_near) near = min_near;
        nears[n] = near; fars[n] = far;
    }
}

/* raymarching.cu:163-198 */
ORACLE_API void oracle_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
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
}

/* raymarching.cu:214-226 / 237-254 */
ORACLE_API void oracle_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)morton3d((uint32_t)coords[n*3], (uint32_t)coords[n*3+1], (uint32_t)coords[n*3+2]);
}
ORACLE_API void oracle_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n];
        coords[n*3]   = (int32_t)morton3d_invert1((uint32_t)(ind >> 0));
        coords[n*3+1] = (int32_t)morton3d_invert1((uint32_t)(ind >> 1));
        coords[n*3+2] = (int32_t)morton3d_invert1((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:268-289 */
ORACLE_API void oracle_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n*8+i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:303-319 */
ORACLE_API void oracle_flatten_rays(const int32_t* rays, uint32_t N, uint32_t M, int32_t* res) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
        for (uint32_t i = 0; i < cnt && off + i < M; i++) res[off + i] = (int32_t)n;
    }
}

/* ------------------------------------------------------------------ */
/* raymarching.cu:338-475 (train) and :714-829 (inference) share this   */
/* stepping body.  Returns number of samples emitted; advances *t_io.   */
/* ------------------------------------------------------------------ */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, H3, Hf, Cf;
    uint32_t H; int contract;
    const uint8_t* grid;
} march_ctx;

static void march_ctx_init(march_ctx* c, const float* o, const float* d, float bound, int contract,
                           float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid) {
    c->ox = o[0]; c->oy = o[1]; c->oz = o[2]; c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
    c->rdx = 1.0f / c->dx; c->rdy = 1.0f / c->dy; c->rdz = 1.0f / c->dz;
    c->bound = bound; c->dt_gamma = dt_gamma; c->contract = contract; c->H = H; c->grid = grid;
    c->Hf = (float)H; c->Cf = (float)C;
    c->rH = 1.0f / (float)H;
    c->H3 = (float)(H * H * H);
    c->dt_min = (2.0f * 1.7320508075688772f) / (float)max_steps;           /* :385 */
    c->dt_max = ((2.0f * 1.7320508075688772f) * bound) / (float)H;          /* :386 */
}

static uint32_t march_body(const march_ctx* c, float* t_io, float far, uint32_t max_emit,
                           float* xyzs, float* dirs, float* ts) {
    float t = *t_io;
    uint32_t step = 0;
    while (t < far && step < max_emit) {
        /* :398-402 — position = fma(t, d, o), clamped */
        const float x = clampf(fmaf(t, c->dx, c->ox), -c->bound, c->bound);
        const float y = clampf(fmaf(t, c->dy, c->oy), -c->bound, c->bound);
        const float z = clampf(fmaf(t, c->dz, c->oz), -c->bound, c->bound);
        float dt = clampf(t * c->dt_gamma, c->dt_min, c->dt_max);
        /* :405-408 */
        const int lp = mip_from_pos(x, y, z, c->Cf), ld = mip_from_dt(dt, c->Hf, c->Cf);
        const int level = lp > ld ? lp : ld;
        const float mip_bound = fminf(scalbnf(1.0f, level), c->bound);
        const float mip_rbound = 1.0f / mip_bound;
        /* :411-419 contraction */
        float cx = x, cy = y, cz = z;
        const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        if (c->contract && mag > 1) {
            const float s = (2 - 1 / mag) / mag;
            cx *= s; cy *= s; cz *= s;
        }
        /* :422-424 — 0.5*(c*rb+1)*H in double (exact), clamp, truncate */
        const int nx = (int)clampf((float)(0.5 * (double)fmaf(cx, mip_rbound, 1.0f) * (double)c->H), 0.0f, (float)(c->H - 1));
        const int ny = (int)clampf((float)(0.5 * (double)fmaf(cy, mip_rbound, 1.0f) * (double)c->H), 0.0f, (float)(c->H - 1));
        const int nz = (int)clampf((float)(0.5 * (double)fmaf(cz, mip_rbound, 1.0f) * (double)c->H), 0.0f, (float)(c->H - 1));
        /* :426-427 — index computed in float */
        const uint32_t index = (uint32_t)fmaf((float)level, c->H3, (float)morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
        const int occ = (c->grid[index / 8] >> (index % 8)) & 1;
        if (occ) {
            /* :432-447 */
            t += dt;
            if (xyzs) {
                xyzs[step*3] = cx; xyzs[step*3+1] = cy; xyzs[step*3+2] = cz;
                dirs[step*3] = c->dx; dirs[step*3+1] = c->dy; dirs[step*3+2] = c->dz;
                ts[step*2] = t; ts[step*2+1] = dt;
            }
            step++;
        } else if (c->contract && mag > 1) {
            t += dt;                                                         /* :449-450 */
        } else {
            /* :452-463 — distance to voxel exit, then step until past it */
            const float tx = fmaf(mip_bound, fmaf((fmaf(0.5f, signf_(c->dx), (float)nx + 0.5f)) * c->rH, 2.0f, -1.0f), -cx) * c->rdx;
            const float ty = fmaf(mip_bound, fmaf((fmaf(0.5f, signf_(c->dy), (float)ny + 0.5f)) * c->rH, 2.0f, -1.0f), -cy) * c->rdy;
            const float tz = fmaf(mip_bound, fmaf((fmaf(0.5f, signf_(c->dz), (float)nz + 0.5f)) * c->rH, 2.0f, -1.0f), -cz) * c->rdz;
            const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
            do {
                dt = clampf(t * c->dt_gamma, c->dt_min, c->dt_max);
                t += dt;
            } while (t < tt);
        }
    }
    *t_io = t;
    return step;
}

/* raymarching.cu:338-475 + raymarching/raymarching.py:197-258.
 * Pass 1 (xyzs == NULL): rays[n] = (offset, count) with offsets assigned in
 * ray order (the reference's atomicAdd order is non-deterministic; canonical
 * order = exclusive prefix sum over n).  Returns M.
 * Pass 2: writes samples at the offsets found in rays. */
ORACLE_API uint32_t oracle_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                            float bound, int contract, float dt_gamma, uint32_t max_steps,
                                            uint32_t N, uint32_t C, uint32_t H,
                                            const float* nears, const float* fars, const float* noises,
                                            float* xyzs, float* dirs, float* ts, int32_t* rays) {
    uint32_t M = 0;
    for (uint32_t n = 0; n < N; n++) {
        march_ctx c; march_ctx_init(&c, rays_o + n*3, rays_d + n*3, bound, contract, dt_gamma, max_steps, C, H, grid);
        const float near = nears[n], far = fars[n];
        /* :389-391 */
        float t = fmaf(clampf(near * dt_gamma, c.dt_min, c.dt_max), noises[n], near);
        if (!xyzs) {
            const uint32_t cnt = march_body(&c, &t, far, max_steps, NULL, NULL, NULL);
            rays[n*2] = (int32_t)M; rays[n*2+1] = (int32_t)cnt;
            M += cnt;
        } else {
            const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
            march_body(&c, &t, far, cnt, xyzs + (size_t)off*3, dirs + (size_t)off*3, ts + (size_t)off*2);
            M += cnt;
        }
    }
    return M;
}

/* raymarching.cu:501-579 */
ORACLE_API void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts,
                                                    const int32_t* rays, uint32_t M, uint32_t N, float T_thresh, int binarize,
                                                    float* weights, float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
        if (cnt == 0 || off + cnt > M) {
            weights_sum[n] = 0; depth[n] = 0; image[n*3] = image[n*3+1] = image[n*3+2] = 0; continue;
        }
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
        for (uint32_t s = 0; s < cnt; s++) {
            const uint32_t i = off + s;
            const float real_alpha = 1.0f - expf(-sigmas[i] * ts[i*2+1]);   /* __expf on device */
            const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            const float w = alpha * T;
            weights[i] = w;
            r = fmaf(w, rgbs[i*3], r); g = fmaf(w, rgbs[i*3+1], g); b = fmaf(w, rgbs[i*3+2], b);
            ws += w; d = fmaf(w, ts[i*2], d);
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
        }
        weights_sum[n] = ws; depth[n] = d; image[n*3] = r; image[n*3+1] = g; image[n*3+2] = b;
    }
}

/* raymarching.cu:606-695 */
ORACLE_API void oracle_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum,
                                                     const float* grad_depth, const float* grad_image,
                                                     const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                                                     const float* weights_sum, const float* depth, const float* image,
                                                     uint32_t M, uint32_t N, float T_thresh, int binarize,
                                                     float* grad_sigmas, float* grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t off = (uint32_t)rays[n*2], cnt = (uint32_t)rays[n*2+1];
        if (cnt == 0 || off + cnt > M) continue;
        const float gi0 = grad_image[n*3], gi1 = grad_image[n*3+1], gi2 = grad_image[n*3+2];
        const float rf = image[n*3], gf = image[n*3+1], bf = image[n*3+2], wsf = weights_sum[n], df = depth[n];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
        for (uint32_t s = 0; s < cnt; s++) {
            const uint32_t i = off + s;
            const float real_alpha = 1.0f - expf(-sigmas[i] * ts[i*2+1]);
            const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            const float w = alpha * T;
            r = fmaf(w, rgbs[i*3], r); g = fmaf(w, rgbs[i*3+1], g); b = fmaf(w, rgbs[i*3+2], b);
            ws += w; d = fmaf(w, ts[i*2], d);
            T *= 1.0f - alpha;
            grad_rgbs[i*3] = gi0 * w; grad_rgbs[i*3+1] = gi1 * w; grad_rgbs[i*3+2] = gi2 * w;
            grad_sigmas[i] = ts[i*2+1] * (
                gi0 * (T * rgbs[i*3]   - (rf - r)) +
                gi1 * (T * rgbs[i*3+1] - (gf - g)) +
                gi2 * (T * rgbs[i*3+2] - (bf - b)) +
                (grad_weights_sum[n] + grad_weights[i]) * (T - (wsf - ws)) +
                grad_depth[n] * (T * ts[i*2] - (df - d)));
            if (T < T_thresh) break;
        }
    }
}

/* raymarching.cu:714-829 */
ORACLE_API void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                                  const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                                  uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                                  const float* nears, const float* fars, float* xyzs, float* dirs, float* ts, const float* noises) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        march_ctx c; march_ctx_init(&c, rays_o + (size_t)index*3, rays_d + (size_t)index*3, bound, contract, dt_gamma, max_steps, C, H, grid);
        float t = rays_t[index];
        t = fmaf(clampf(t * dt_gamma, c.dt_min, c.dt_max), noises[n], t);     /* :756-757 */
        (void)nears;
        march_body(&c, &t, fars[index], n_step, xyzs + (size_t)n*n_step*3, dirs + (size_t)n*n_step*3, ts + (size_t)n*n_step*2);
    }
}

/* raymarching.cu:843-925 */
ORACLE_API void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize,
                                      int32_t* rays_alive, float* rays_t, const float* sigmas, const float* rgbs, const float* ts,
                                      float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float* sg = sigmas + (size_t)n*n_step; const float* rg = rgbs + (size_t)n*n_step*3; const float* tp = ts + (size_t)n*n_step*2;
        float t = 0, d = depth[index], r = image[index*3], g = image[index*3+1], b = image[index*3+2], wsum = weights_sum[index];
        uint32_t step = 0;
        while (step < n_step) {
            if (tp[0] == 0) break;
            const float real_alpha = 1.0f - expf(-sg[0] * tp[1]);
            const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            const float T = 1 - wsum;
            const float w = alpha * T;
            wsum += w;
            t = tp[0];
            d = fmaf(w, t, d); r = fmaf(w, rg[0], r); g = fmaf(w, rg[1], g); b = fmaf(w, rg[2], b);
            if (T < T_thresh) break;
            sg++; rg += 3; tp += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = wsum; depth[index] = d; image[index*3] = r; image[index*3+1] = g; image[index*3+2] = b;
    }
}

/* ------------------------------------------------------------------ */
/* gridencoder/src/gridencoder.cu                                      */
/* ------------------------------------------------------------------ */
static inline float h2f_round(float v) { return (float)(_Float16)v; }   /* round-trip through fp16 (RNE) */

static inline uint32_t grid_index(uint32_t gridtype, uint32_t D, uint32_t hashmap_size, uint32_t resolution, const uint32_t* pg) {
    /* gridencoder.cu:45-79 (without the *C + ch) */
    static const uint32_t primes[7] = { 1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u };
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) { index += pg[d] * stride; stride *= resolution; }
    if (gridtype == 0 && stride > hashmap_size) {
        index = 0;
        for (uint32_t d = 0; d < D; d++) index ^= pg[d] * primes[d];
    }
    return index % hashmap_size;
}

ORACLE_API uint32_t oracle_grid_resolution(uint32_t level, float S, uint32_t H) {
    /* gridencoder.cu:133 */
    return (uint32_t)ceilf(exp2f((float)level * S) * (float)H);
}

/* gridencoder.cu:83-249.  half != 0 emulates the scalar_t = at::Half build:
 * the table and outputs are given as float arrays holding fp16-representable
 * values and every `results[ch] += w * grid[...]` rounds product and sum to
 * fp16 (Half += float semantics of c10::Half).  res_override (optional, [L])
 * replaces ceil(exp2f(level*S)*H) so a test can feed the device's values. */
ORACLE_API void oracle_grid_encode_forward(const float* inputs, const float* grid, const int32_t* offsets, float* outputs,
                                           uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                           float* dy_dx, uint32_t gridtype, int align_corners, uint32_t interp, int half,
                                           const uint32_t* res_override) {
    for (uint32_t level = 0; level < max_level; level++) {
        const float* g = grid + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level+1] - offsets[level]);
        const uint32_t resolution = res_override ? res_override[level] : oracle_grid_resolution(level, S, H);
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b*D;
            float* out = outputs + ((size_t)level*B + b)*C;
            float* dd = dy_dx ? dy_dx + (size_t)b*D*L*C + (size_t)level*D*C : NULL;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) {
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
                if (dd) for (uint32_t i = 0; i < D*C; i++) dd[i] = 0;
                continue;
            }
            float pos[5], pos_deriv[5]; uint32_t pg[5];
            for (uint32_t d = 0; d < D; d++) {
                if (align_corners) {
                    pos[d] = in[d] * (float)(resolution - 1);
                    uint32_t f = (uint32_t)floorf(pos[d]); pg[d] = f < resolution - 2 ? f : resolution - 2;
                } else {
                    pos[d] = fminf(fmaxf(fmaf(in[d], (float)resolution, -0.5f), 0.0f), (float)(resolution - 1));
                    pg[d] = (uint32_t)floorf(pos[d]);
                }
                pos[d] -= (float)pg[d];
                if (interp == 1) {
                    const float v = pos[d];
                    pos_deriv[d] = 6*v*(1.0f - v);
                    pos[d] = v*v*(3.0f - 2.0f*v);
                } else pos_deriv[d] = 1.0f;
            }
            float res[32]; for (uint32_t ch = 0; ch < C; ch++) res[ch] = 0;
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1; uint32_t pl[5];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1 < resolution - 1 ? pg[d] + 1 : resolution - 1; }
                }
                const uint32_t index = grid_index(gridtype, D, hashmap_size, resolution, pl) * C;
                for (uint32_t ch = 0; ch < C; ch++) {
                    if (half) res[ch] = h2f_round(res[ch] + h2f_round(w * g[index + ch]));
                    else res[ch] = fmaf(w, g[index + ch], res[ch]);
                }
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = res[ch];
            if (dd) {
                for (uint32_t gd = 0; gd < D; gd++) {
                    float rg[32]; for (uint32_t ch = 0; ch < C; ch++) rg[ch] = 0;
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = (float)(align_corners ? resolution - 1 : resolution);
                        uint32_t pl[5];
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                            else { w *= pos[d]; pl[d] = pg[d] + 1 < resolution - 1 ? pg[d] + 1 : resolution - 1; }
                        }
                        pl[gd] = pg[gd];
                        const uint32_t il = grid_index(gridtype, D, hashmap_size, resolution, pl) * C;
                        pl[gd] = pg[gd] + 1 < resolution - 1 ? pg[gd] + 1 : resolution - 1;
                        const uint32_t ir = grid_index(gridtype, D, hashmap_size, resolution, pl) * C;
                        for (uint32_t ch = 0; ch < C; ch++) {
                            if (half) {
                                /* Half - Half -> Half (rounded); float*Half*float -> float; Half += float */
                                const float diff = h2f_round(g[ir + ch] - g[il + ch]);
                                rg[ch] = h2f_round(rg[ch] + h2f_round(w * diff * pos_deriv[gd]));
                            } else rg[ch] += w * (g[ir + ch] - g[il + ch]) * pos_deriv[gd];
                        }
                    }
                    for (uint32_t ch = 0; ch < C; ch++) dd[gd*C + ch] = rg[ch];
                }
            }
        }
    }
}

/* gridencoder.cu:253-349.  Accumulates in double so the result is the
 * order-independent exact sum the atomics approximate. grad: [L,B,C]. */
ORACLE_API void oracle_grid_encode_backward(const float* grad, const float* inputs, const int32_t* offsets, double* grad_grid,
                                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                            uint32_t gridtype, int align_corners, uint32_t interp, int half,
                                            const uint32_t* res_override) {
    (void)L;
    for (uint32_t level = 0; level < max_level; level++) {
        double* gg = grad_grid + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level+1] - offsets[level]);
        const uint32_t resolution = res_override ? res_override[level] : oracle_grid_resolution(level, S, H);
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b*D;
            const float* gr = grad + ((size_t)level*B + b)*C;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            float pos[5]; uint32_t pg[5];
            for (uint32_t d = 0; d < D; d++) {
                if (align_corners) {
                    pos[d] = in[d] * (float)(resolution - 1);
                    uint32_t f = (uint32_t)floorf(pos[d]); pg[d] = f < resolution - 2 ? f : resolution - 2;
                } else {
                    pos[d] = fminf(fmaxf(fmaf(in[d], (float)resolution, -0.5f), 0.0f), (float)(resolution - 1));
                    pg[d] = (uint32_t)floorf(pos[d]);
                }
                pos[d] -= (float)pg[d];
                if (interp == 1) { const float v = pos[d]; pos[d] = v*v*(3.0f - 2.0f*v); }
            }
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1; uint32_t pl[5];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1 < resolution - 1 ? pg[d] + 1 : resolution - 1; }
                }
                const uint32_t index = grid_index(gridtype, D, hashmap_size, resolution, pl) * C;
                for (uint32_t ch = 0; ch < C; ch++) {
                    const float v = w * gr[ch];
                    gg[index + ch] += half ? (double)h2f_round(v) : (double)v;
                }
            }
        }
    }
}

/* gridencoder.cu:353-378 */
ORACLE_API void oracle_grid_input_backward(const float* grad, const float* dy_dx, float* grad_inputs,
                                           uint32_t B, uint32_t D, uint32_t C, uint32_t L, int half) {
    for (uint32_t b = 0; b < B; b++) for (uint32_t d = 0; d < D; d++) {
        const float* dd = dy_dx + (size_t)b*L*D*C;
        float result = 0;
        for (uint32_t l = 0; l < L; l++) for (uint32_t ch = 0; ch < C; ch++) {
            const float p = grad[((size_t)l*B + b)*C + ch] * dd[(size_t)l*D*C + d*C + ch];
            result = half ? h2f_round(result + h2f_round(p)) : result + p;
        }
        grad_inputs[(size_t)b*D + d] = result;
    }
}

/* gridencoder.cu:526-631 (fp32 only; the wrapper runs with autocast disabled, grid.py:172) */
ORACLE_API void oracle_grad_total_variation(const float* inputs, const float* grid, double* grad, const int32_t* offsets, float weight,
                                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                            uint32_t gridtype, int align_corners, const uint32_t* res_override) {
    for (uint32_t level = 0; level < L; level++) {
        const float* g = grid + (size_t)(uint32_t)offsets[level] * C;
        double* gr = grad + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level+1] - offsets[level]);
        const uint32_t resolution = res_override ? res_override[level] : oracle_grid_resolution(level, S, H);
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b*D;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            uint32_t pg[5];
            for (uint32_t d = 0; d < D; d++) {
                if (align_corners) {
                    const float p = in[d] * (float)(resolution - 1);
                    uint32_t f = (uint32_t)floorf(p); pg[d] = f < resolution - 2 ? f : resolution - 2;
                } else {
                    const float p = fminf(fmaxf(fmaf(in[d], (float)resolution, -0.5f), 0.0f), (float)(resolution - 1));
                    pg[d] = (uint32_t)floorf(p);
                }
            }
            float results[32], idelta[32];
            for (uint32_t ch = 0; ch < C; ch++) results[ch] = idelta[ch] = 0;
            const uint32_t index = grid_index(gridtype, D, hashmap_size, resolution, pg) * C;
            const float w = weight / (float)(2 * D);
            for (uint32_t d = 0; d < D; d++) {
                const uint32_t cur = pg[d];
                if (cur < resolution) {
                    pg[d] = cur + 1;
                    const uint32_t ir = grid_index(gridtype, D, hashmap_size, resolution, pg) * C;
                    for (uint32_t ch = 0; ch < C; ch++) { const float v = g[index+ch] - g[ir+ch]; results[ch] += v; idelta[ch] += v*v; }
                }
                if (cur > 0) {
                    pg[d] = cur - 1;
                    const uint32_t il = grid_index(gridtype, D, hashmap_size, resolution, pg) * C;
                    for (uint32_t ch = 0; ch < C; ch++) { const float v = g[index+ch] - g[il+ch]; results[ch] += v; idelta[ch] += v*v; }
                }
                pg[d] = cur;
            }
            for (uint32_t ch = 0; ch < C; ch++) gr[index + ch] += (double)(w * results[ch] * (1.0f / sqrtf(idelta[ch] + 1e-9f)));
        }
    }
}

/* gridencoder.cu:671-703 */
ORACLE_API void oracle_grad_weight_decay(const float* grid, float* grad, const int32_t* offsets, float weight,
                                         uint32_t B, uint32_t C, uint32_t L) {
    for (uint32_t b = 0; b < B * C; b++) {
        const uint32_t n = b / C;
        uint32_t level = 0, l = 0, r = L;
        while (l < r) { const uint32_t m = (l + r) / 2; if ((uint32_t)offsets[m] <= n) { level = m; l = m + 1; } else r = m; }
        const uint32_t hashmap_size = (uint32_t)(offsets[level+1] - offsets[level]);
        grad[b] += 2 * weight * grid[b] / (float)hashmap_size;
    }
}

/* ------------------------------------------------------------------ */
/* freqencoder/src/freqencoder.cu:30-94                                */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs) {
    (void)deg;
    const float HALF_PI = 3.141592653589793f / 2;
    for (uint32_t b = 0; b < B; b++) for (uint32_t c = 0; c < C; c++) {
        float* o = outputs + (size_t)b*C + c;
        if (c < D) { *o = inputs[(size_t)b*D + c]; continue; }
        const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
        const float phase = (float)(col % 2) * HALF_PI;
        *o = sinf(scalbnf(inputs[(size_t)b*D + d], (int)freq) + phase);     /* __sinf on device */
    }
}
ORACLE_API void oracle_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* grad_inputs) {
    for (uint32_t b = 0; b < B; b++) for (uint32_t d = 0; d < D; d++) {
        const float* g = grad + (size_t)b*C; const float* o = outputs + (size_t)b*C;
        float result = g[d];
        g += D; o += D;
        for (uint32_t f = 0; f < deg; f++) {
            result += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
            g += 2*D; o += 2*D;
        }
        grad_inputs[(size_t)b*D + d] = result;
    }
}

/* ------------------------------------------------------------------ */
/* shencoder/src/shencoder.cu:28-383.                                  */
/* The reference lists 64 closed-form polynomials and their 192        */
/* partial derivatives.  They are the real spherical harmonics         */
/*   Y_l^m = (-1)^m N_lm * Q_l^|m|(z) * (m>=0 ? Re : Im)(x+iy)^|m| * (m?sqrt2:1) */
/* with Q_l^m = d^m/dz^m P_l(z) and N_lm = sqrt((2l+1)/(4pi) (l-m)!/(l+m)!), */
/* written as polynomials in independent x,y,z (index l*l+l+m,         */
/* shencoder.cu:50-120); the dx/dy/dz tables (:130-350) are the exact  */
/* partials of those polynomials.  Restated here generatively in       */
/* double, rounded to fp32 on output.                                  */
/* ------------------------------------------------------------------ */
static void legendre_coeffs(int l, double* c /* [l+1], ascending powers */) {
    /* P_l via Bonnet recurrence on coefficient vectors */
    double p0[16] = {0}, p1[16] = {0}, p2[16];
    p0[0] = 1; p1[1] = 1;
    if (l == 0) { memcpy(c, p0, sizeof(double)*(size_t)(l+1)); return; }
    if (l == 1) { memcpy(c, p1, sizeof(double)*(size_t)(l+1)); return; }
    for (int n = 1; n < l; n++) {
        memset(p2, 0, sizeof p2);
        for (int k = 0; k <= n; k++) p2[k+1] += (2.0*n + 1) * p1[k] / (n + 1);
        for (int k = 0; k <= n - 1; k++) p2[k] -= (double)n * p0[k] / (n + 1);
        memcpy(p0, p1, sizeof p0); memcpy(p1, p2, sizeof p1);
    }
    memcpy(c, p1, sizeof(double)*(size_t)(l+1));
}
static double polyval(const double* c, int deg, double z) { double r = 0; for (int k = deg; k >= 0; k--) r = r*z + c[k]; return deg < 0 ? 0 : r; }
static void polyder(double* c, int* deg) { if (*deg <= 0) { c[0] = 0; *deg = (*deg == 0) ? -1 : *deg; return; } for (int k = 0; k < *deg; k++) c[k] = c[k+1]*(k+1); (*deg)--; }
static double factorial(int n) { double r = 1; for (int i = 2; i <= n; i++) r *= i; return r; }

ORACLE_API void oracle_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx) {
    const uint32_t C2 = degree * degree;
    const double PI_ = 3.14159265358979323846;
    for (uint32_t b = 0; b < B; b++) {
        const double x = inputs[(size_t)b*D], y = inputs[(size_t)b*D+1], z = inputs[(size_t)b*D+2];
        /* A_m + i B_m = (x+iy)^m */
        double A[9], Bm[9]; A[0] = 1; Bm[0] = 0;
        for (int m = 1; m <= 8; m++) { A[m] = A[m-1]*x - Bm[m-1]*y; Bm[m] = A[m-1]*y + Bm[m-1]*x; }
        float* out = outputs + (size_t)b*C2;
        float* ddx = dy_dx ? dy_dx + (size_t)b*D*C2 : NULL;
        float* ddy = ddx ? ddx + C2 : NULL; float* ddz = ddx ? ddy + C2 : NULL;
        for (int l = 0; l < (int)degree; l++) {
            double c[16]; legendre_coeffs(l, c);
            int deg = l;
            double q[16]; memcpy(q, c, sizeof(double)*(size_t)(l+1));
            for (int m = 0; m <= l; m++) {
                /* q = Q_l^m (degree deg); qn = Q_l^{m+1} */
                double qn[16]; int degn = deg; memcpy(qn, q, sizeof q); polyder(qn, &degn);
                const double Qm = polyval(q, deg, z), Qm1 = (degn >= 0) ? polyval(qn, degn, z) : 0.0;
                const double N = sqrt((2.0*l + 1) / (4*PI_) * factorial(l - m) / factorial(l + m)) * (m ? sqrt(2.0) : 1.0) * ((m & 1) ? -1.0 : 1.0);
                const int ip = l*l + l + m, in_ = l*l + l - m;
                out[ip] = (float)(N * Qm * A[m]);
                if (m) out[in_] = (float)(N * Qm * Bm[m]);
                if (ddx) {
                    const double dAx = m ? m*A[m-1] : 0, dAy = m ? -m*Bm[m-1] : 0;
                    const double dBx = m ? m*Bm[m-1] : 0, dBy = m ? m*A[m-1] : 0;
                    ddx[ip] = (float)(N*Qm*dAx); ddy[ip] = (float)(N*Qm*dAy); ddz[ip] = (float)(N*Qm1*A[m]);
                    if (m) { ddx[in_] = (float)(N*Qm*dBx); ddy[in_] = (float)(N*Qm*dBy); ddz[in_] = (float)(N*Qm1*Bm[m]); }
                }
                memcpy(q, qn, sizeof q); deg = degn;
            }
        }
    }
}

/* shencoder.cu:359-383 (accumulates into caller-zeroed grad_inputs) */
ORACLE_API void oracle_sh_encode_backward(const float* grad, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx, float* grad_inputs) {
    const uint32_t C2 = degree * degree;
    for (uint32_t b = 0; b < B; b++) for (uint32_t d = 0; d < D; d++) {
        float acc = grad_inputs[(size_t)b*D + d];
        for (uint32_t ch = 0; ch < C2; ch++) acc += grad[(size_t)b*C2 + ch] * dy_dx[(size_t)b*D*C2 + (size_t)d*C2 + ch];
        grad_inputs[(size_t)b*D + d] = acc;
    }
}
