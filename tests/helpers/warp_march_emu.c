/*
 * warp_march_emu.c — scalar emulation of the warp-per-ray marching scheme used
 * by stable-dreamfusion_b200/csrc/raymarch.cu (march_ray_warp): 32 chain points
 * per iteration + ballot-style resolution of the visited set.  Test helper:
 * tests/test_march_algorithm.py checks it against the sequential restatement
 * in oracle/sdf_oracle.c, so the resolution logic is validated on CPU before
 * the CUDA kernel ever runs.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o libwarp_march_emu.so warp_march_emu.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v;
}
static inline uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
static inline int mip_level(float mx, float Cf) { int e; frexpf(mx, &e); return (int)fminf(Cf - 1.0f, fmaxf(0.0f, (float)e)); }
static inline int ffs32(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }

uint32_t emu_march_ray(const float* o, const float* d, const uint8_t* grid, float bound, int contract, float dt_gamma,
                       uint32_t max_steps, uint32_t C, uint32_t H, float near, float far, float noise, uint32_t max_emit,
                       float* xyzs, float* dirs, float* ts) {
    const float ox = o[0], oy = o[1], oz = o[2], dx = d[0], dy = d[1], dz = d[2];
    const float Hf = (float)H, Hm1f = (float)(H - 1), Cf = (float)C, rH = 1.0f / (float)H, H3 = (float)(H*H*H);
    const float dt_min = (2.0f * 1.7320508075688772f) / (float)max_steps;
    const float dt_max = ((2.0f * 1.7320508075688772f) * bound) / (float)H;
    const float rdx = 1.0f/dx, rdy = 1.0f/dy, rdz = 1.0f/dz;
    const float sx = copysignf(1.0f, dx), sy = copysignf(1.0f, dy), sz = copysignf(1.0f, dz);
    const float t0 = fmaf(clampf(near * dt_gamma, dt_min, dt_max), noise, near);
    uint32_t count = 0;
    float t_base = t0, skip_until = -INFINITY;
    int done = !(t0 < far) || max_emit == 0;
    while (!done) {
        float my_t[32], my_dt[32], tt[32], cx[32], cy[32], cz[32];
        int valid[32], occ[32];
        float t = t_base;
        for (int j = 0; j < 32; j++) { const float dtj = clampf(t * dt_gamma, dt_min, dt_max); my_t[j] = t; my_dt[j] = dtj; t = t + dtj; }
        t_base = t;
        uint32_t occ_mask = 0, valid_mask = 0;
        for (int L = 0; L < 32; L++) {
            valid[L] = my_t[L] < far; occ[L] = 0; tt[L] = -INFINITY; cx[L] = cy[L] = cz[L] = 0;
            if (valid[L]) {
                const float x = clampf(fmaf(my_t[L], dx, ox), -bound, bound);
                const float y = clampf(fmaf(my_t[L], dy, oy), -bound, bound);
                const float z = clampf(fmaf(my_t[L], dz, oz), -bound, bound);
                const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
                const int l1 = mip_level(mag, Cf), l2 = mip_level((my_dt[L] * Hf) * 0.5f, Cf);
                const int level = l1 > l2 ? l1 : l2;
                const float mip_bound = fminf(scalbnf(1.0f, level), bound);
                const float mip_rbound = 1.0f / mip_bound;
                cx[L] = x; cy[L] = y; cz[L] = z;
                const int contracted = contract && mag > 1.0f;
                if (contracted) { const float s = (2.0f - 1.0f / mag) / mag; cx[L] *= s; cy[L] *= s; cz[L] *= s; }
                const int nx = (int)clampf((0.5f * fmaf(cx[L], mip_rbound, 1.0f)) * Hf, 0.0f, Hm1f);
                const int ny = (int)clampf((0.5f * fmaf(cy[L], mip_rbound, 1.0f)) * Hf, 0.0f, Hm1f);
                const int nz = (int)clampf((0.5f * fmaf(cz[L], mip_rbound, 1.0f)) * Hf, 0.0f, Hm1f);
                const uint32_t index = (uint32_t)fmaf((float)level, H3, (float)morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
                occ[L] = (grid[index >> 3] >> (index & 7)) & 1;
                if (!occ[L] && !contracted) {
                    const float tx = fmaf(mip_bound, fmaf(fmaf(0.5f, sx, (float)nx + 0.5f) * rH, 2.0f, -1.0f), -cx[L]) * rdx;
                    const float ty = fmaf(mip_bound, fmaf(fmaf(0.5f, sy, (float)ny + 0.5f) * rH, 2.0f, -1.0f), -cy[L]) * rdy;
                    const float tz = fmaf(mip_bound, fmaf(fmaf(0.5f, sz, (float)nz + 0.5f) * rH, 2.0f, -1.0f), -cz[L]) * rdz;
                    tt[L] = my_t[L] + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
                }
            }
            if (occ[L] && valid[L]) occ_mask |= 1u << L;
            if (valid[L]) valid_mask |= 1u << L;
        }
        uint32_t cur = 0;
        while (cur < 32) {
            uint32_t ok = 0;
            for (int L = 0; L < 32; L++) if (my_t[L] >= skip_until) ok |= 1u << L;
            ok &= (0xffffffffu << cur);
            if (ok == 0) break;
            cur = (uint32_t)ffs32(ok) - 1;
            if (!((valid_mask >> cur) & 1u)) { done = 1; break; }
            const uint32_t rest = occ_mask >> cur;
            const uint32_t run = (rest == 0xffffffffu) ? 32u : (uint32_t)ffs32(~rest) - 1;
            if (run > 0) {
                const uint32_t take = run < max_emit - count ? run : max_emit - count;
                if (xyzs) for (uint32_t L = cur; L < cur + take; L++) {
                    const uint32_t s = count + (L - cur);
                    xyzs[s*3] = cx[L]; xyzs[s*3+1] = cy[L]; xyzs[s*3+2] = cz[L];
                    dirs[s*3] = dx; dirs[s*3+1] = dy; dirs[s*3+2] = dz;
                    ts[s*2] = my_t[L] + my_dt[L]; ts[s*2+1] = my_dt[L];
                }
                count += take;
                if (count >= max_emit) { done = 1; break; }
                cur += run;
                skip_until = -INFINITY;
                if (cur >= 32) break;
            } else {
                skip_until = tt[cur];
                cur += 1;
            }
        }
        if (!(t_base < far) && !done) done = 1;
    }
    return count;
}
