// field_common.cuh — device building blocks of the fused radiance-field kernels (fused_field.cu):
// hash-grid lookup emitted directly in mma.sync A-fragment order, the 32-64-64-4 MLP on tensor
// cores with register-chained activations, and the weight tiles in shared memory.
#pragma once
#include "grid_common.cuh"
#include <mma.h>

namespace field {

constexpr int kLevels = 16;          // instant-NGP config of the -O backbone (nerf/network_grid.py:49, encoding.py:57)
constexpr int kEncDim = 32;          // kLevels * 2 features
constexpr int kHidden = 64;
constexpr int kOut = 4;              // sigma logit + 3 albedo logits
constexpr int kW1Stride = kEncDim + 8;   // halfs; +8 keeps B-fragment LDS conflict-free
constexpr int kW2Stride = kHidden + 8;

struct LevelSmem {
    uint32_t offset, size, res, flags;   // flags bit0: hashed, bit1: size is a power of two
};

// Shared-memory weight block (fp16 weights, fp32 biases) — forward orientation W[out][in].
struct WeightsSmem {
    __half w1[kHidden][kW1Stride];
    __half w2[kHidden][kW2Stride];
    __half w3[8][kW2Stride];             // rows 4..7 are zero
    float b1[kHidden], b2[kHidden], b3[8];
    LevelSmem lv[kLevels];
};

struct FieldParams {
    const __half2* table;                // fp16 hash table, [n_entries] half2 (2 features)
    const float *w1, *b1, *w2, *b2, *w3, *b3;   // fp32 master weights of sigma_net (nn.Linear layout [out,in])
    const LevelParams* lp;
    float bound;
    uint32_t n_levels_active;            // levels >= this produce zeros (progressive max_level)
    float blob_density, blob_radius;     // density_blob (nerf/renderer.py:339-349), exp activation
    int interp_smoothstep;
};

__device__ __forceinline__ void load_weights(WeightsSmem& s, const FieldParams& p) {
    for (int i = threadIdx.x; i < kHidden * kEncDim; i += blockDim.x) s.w1[i / kEncDim][i % kEncDim] = __float2half_rn(p.w1[i]);
    for (int i = threadIdx.x; i < kHidden * kHidden; i += blockDim.x) s.w2[i / kHidden][i % kHidden] = __float2half_rn(p.w2[i]);
    for (int i = threadIdx.x; i < 8 * kHidden; i += blockDim.x)
        s.w3[i / kHidden][i % kHidden] = (i / kHidden) < kOut ? __float2half_rn(p.w3[i]) : __float2half_rn(0.f);
    for (int i = threadIdx.x; i < kHidden; i += blockDim.x) {
        // biases pass through fp16 like the autocast nn.Linear of the reference
        s.b1[i] = __half2float(__float2half_rn(p.b1[i]));
        s.b2[i] = __half2float(__float2half_rn(p.b2[i]));
    }
    if (threadIdx.x < 8) s.b3[threadIdx.x] = threadIdx.x < kOut ? __half2float(__float2half_rn(p.b3[threadIdx.x])) : 0.f;
    if (threadIdx.x < kLevels) {
        const uint32_t l = threadIdx.x;
        LevelSmem v;
        v.offset = p.lp->offset[l]; v.size = p.lp->size[l]; v.res = p.lp->res[l];
        // index rule of the reference (gridencoder.cu:62-79): dense while stride <= size, hash otherwise
        uint32_t stride = 1;
        for (int d = 0; d < 3; d++) if (stride <= v.size) stride *= v.res;
        v.flags = (stride > v.size ? 1u : 0u) | (((v.size & (v.size - 1)) == 0) ? 2u : 0u);
        s.lv[l] = v;
    }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Four 8x8 b16 matrices from shared memory in one instruction: lane l supplies the address of row (l & 7) of matrix (l >> 3);
// thread (g, t) receives M_i[g][2t..2t+1] in r[i] — exactly the mma.sync B fragment of a [n][k] (k contiguous) operand, or
// the A fragment of a [m][k] operand.  Replaces four 32-bit LDS per lane (the kernels are shared-memory-issue bound).
__device__ __forceinline__ void ldsm_x4(uint32_t r[4], const void* smem_row) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// B fragments of TWO consecutive k-steps (k0 .. k0+31) of the n-tile whose first row is `rows` (row stride in halfs):
// r[0], r[1] = (b0, b1) of k-step 0; r[2], r[3] = (b0, b1) of k-step 1.
__device__ __forceinline__ void ldsm_b2(uint32_t r[4], const __half* rows, int stride, int k0, int lane) {
    ldsm_x4(r, rows + (lane & 7) * stride + k0 + (lane >> 3) * 8);
}
// B fragments of ONE k-step (k0 .. k0+15) for the TWO n-tiles at n0 and n0 + 8 of an operand stored [n][k] (k contiguous):
// r[0], r[1] = (b0, b1) of n-tile n0; r[2], r[3] = of n-tile n0 + 8.
__device__ __forceinline__ void ldsm_b_nn(uint32_t r[4], const __half* base, int stride, int n0, int k0, int lane) {
    ldsm_x4(r, base + (n0 + (lane & 7) + (lane >> 4) * 8) * stride + k0 + ((lane >> 3) & 1) * 8);
}
// A fragment (a0..a3) of the 16 x 16 block at (rows, k0) of a [m][k] operand.
__device__ __forceinline__ void ldsm_a(uint32_t r[4], const __half* rows, int stride, int k0, int lane) {
    ldsm_x4(r, rows + ((lane & 7) + ((lane >> 3) & 1) * 8) * stride + k0 + (lane >> 4) * 8);
}

// Corner geometry of one (point, level): base cell, interpolation fractions (after smoothstep) and
// whether the point is inside the unit cube.
struct Cell {
    uint32_t pg[3];
    float f[3];
    float df[3];      // d(smoothstep)/d(pos) * res  (only filled when WITH_DERIV)
};

template <bool WITH_DERIV>
__device__ __forceinline__ void locate_cell(Cell& c, float x, float y, float z, uint32_t res, bool smooth) {
    const float in[3] = {x, y, z};
    const float rf = (float)res;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float pos = fminf(fmaxf(fmaf(in[d], rf, -0.5f), 0.0f), rf - 1.0f);
        const float fl = floorf(pos);
        c.pg[d] = (uint32_t)fl;
        pos -= fl;
        if (smooth) {
            if (WITH_DERIV) c.df[d] = 6.f * pos * (1.f - pos) * rf;
            pos = pos * pos * (3.f - 2.f * pos);
        } else if (WITH_DERIV) {
            c.df[d] = rf;
        }
        c.f[d] = pos;
    }
}

// The 8 corners of one (point, level): table indices (relative to the level) and trilinear weights in the reference's corner
// order (bit 0 = x, bit 1 = y, bit 2 = z; gridencoder.cu:150-175).  ONE branch per level (hashed / dense; lanes of a warp hold
// different levels, so it only diverges for the 4-level group that straddles the dense->hash switch) and branch-free index
// arithmetic inside, so the caller can issue all of its gathers back to back.
struct Corners {
    uint32_t idx[8];
    float w[8];
};

__device__ __forceinline__ uint32_t level_corners(Corners& c, const LevelSmem& lv, float x, float y, float z, bool smooth) {
    Cell cell;
    locate_cell<false>(cell, x, y, z, lv.res, smooth);
    const uint32_t x0 = cell.pg[0], y0 = cell.pg[1], z0 = cell.pg[2];
    const uint32_t cell_key = x0 | (y0 << 10) | (z0 << 20);          // unique while res <= 1024 (used on the coarse levels only)
    const uint32_t x1 = min(x0 + 1, lv.res - 1), y1 = min(y0 + 1, lv.res - 1), z1 = min(z0 + 1, lv.res - 1);
    const float fx = cell.f[0], fy = cell.f[1], fz = cell.f[2];
    const float wxy[4] = {(1.f - fx) * (1.f - fy), fx * (1.f - fy), (1.f - fx) * fy, fx * fy};
#pragma unroll
    for (int k = 0; k < 8; k++) c.w[k] = wxy[k & 3] * ((k & 4) ? fz : 1.f - fz);
    if (lv.flags & 1u) {
        // spatial hash (gridencoder.cu:53-60): x ^ y*2654435761 ^ z*805459861, modulo the level size
        const uint32_t ya[2] = {y0 * 2654435761u, y1 * 2654435761u}, za[2] = {z0 * 805459861u, z1 * 805459861u};
        const uint32_t xa[2] = {x0, x1};
        if (lv.flags & 2u) {
            const uint32_t mask = lv.size - 1;
#pragma unroll
            for (int k = 0; k < 8; k++) c.idx[k] = (xa[k & 1] ^ ya[(k >> 1) & 1] ^ za[k >> 2]) & mask;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) c.idx[k] = (xa[k & 1] ^ ya[(k >> 1) & 1] ^ za[k >> 2]) % lv.size;
        }
    } else {
        // dense levels (flag clear <=> res^3 <= size, load_weights): the reference's `index % size` (gridencoder.cu:79) is the
        // identity because x + y*res + z*res^2 <= res^3 - 1 < size
        const uint32_t ya[2] = {y0 * lv.res, y1 * lv.res}, za[2] = {z0 * lv.res * lv.res, z1 * lv.res * lv.res};
        const uint32_t xa[2] = {x0, x1};
#pragma unroll
        for (int k = 0; k < 8; k++) c.idx[k] = xa[k & 1] + ya[(k >> 1) & 1] + za[k >> 2];
    }
    return cell_key;
}

__device__ __forceinline__ uint32_t corner_index(const LevelSmem& lv, uint32_t x, uint32_t y, uint32_t z) {
    if (lv.flags & 1u) {
        const uint32_t h = x ^ (y * 2654435761u) ^ (z * 805459861u);
        return (lv.flags & 2u) ? (h & (lv.size - 1)) : (h % lv.size);
    }
    return x + (y + z * lv.res) * lv.res;
}

// Trilinear (smoothstepped) lookup of one level for TWO points in [0,1]^3 (the two rows a lane owns): all 16 gathers are issued
// before the first one is consumed.  Points outside the unit cube (valid == false) still gather (their cell is clamped) and
// are zeroed afterwards.
// Gather of one x-neighbour corner pair (corners 2j, 2j+1).  When both entries sit in the same aligned 8-byte pair of the table
// (always on hashed levels with even x0: x1 = x0 ^ 1 only flips bit 0 of the hash; on dense levels when the linear index is
// even) ONE 8-byte load replaces two 4-byte ones — the kernels are bound by the number of scattered lanes the LSU / L1 has to
// serve, not by bytes.  Predicated loads, no branches: all gathers of a level pair still issue back to back.
struct PairLoad { uint32_t lo, hi, a, b, sel; };
__device__ __forceinline__ void pair_issue(PairLoad& r, const __half2* t, uint32_t i0, uint32_t i1) {
    const uint32_t merged = ((i0 ^ i1) == 1u) ? 1u : 0u;
    r.sel = merged | ((i0 & 1u) << 1);
    r.lo = r.hi = r.a = r.b = 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %7, 0;\n\t"
                 "@p ld.global.nc.v2.u32 {%0, %1}, [%4];\n\t"
                 "@!p ld.global.nc.u32 %2, [%5];\n\t"
                 "@!p ld.global.nc.u32 %3, [%6];\n\t}"
                 : "+r"(r.lo), "+r"(r.hi), "+r"(r.a), "+r"(r.b)
                 : "l"(t + (i0 & ~1u)), "l"(t + i0), "l"(t + i1), "r"(merged));
}
__device__ __forceinline__ void pair_values(const PairLoad& r, float2& v0, float2& v1) {
    const bool merged = r.sel & 1u, odd = r.sel & 2u;
    const uint32_t u0 = merged ? (odd ? r.hi : r.lo) : r.a;
    const uint32_t u1 = merged ? (odd ? r.lo : r.hi) : r.b;
    v0 = __half22float2(*reinterpret_cast<const __half2*>(&u0));
    v1 = __half22float2(*reinterpret_cast<const __half2*>(&u1));
}

__device__ __forceinline__ void encode_level_pair(const __half2* __restrict__ table, const LevelSmem& lv, const float pa[3], bool va,
                                                  const float pb[3], bool vb, bool smooth, float2& ea, float2& eb) {
    Corners ca, cb;
    level_corners(ca, lv, pa[0], pa[1], pa[2], smooth);
    level_corners(cb, lv, pb[0], pb[1], pb[2], smooth);
    const __half2* t = table + lv.offset;          // level offsets are multiples of 8 entries: entry pairs stay 8-byte aligned
    PairLoad la[4], lb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { pair_issue(la[j], t, ca.idx[2 * j], ca.idx[2 * j + 1]); pair_issue(lb[j], t, cb.idx[2 * j], cb.idx[2 * j + 1]); }
    float2 aa = make_float2(0.f, 0.f), ab = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float2 v0, v1, u0, u1;
        pair_values(la[j], v0, v1);
        pair_values(lb[j], u0, u1);
        aa.x = fmaf(ca.w[2 * j], v0.x, aa.x); aa.y = fmaf(ca.w[2 * j], v0.y, aa.y);
        aa.x = fmaf(ca.w[2 * j + 1], v1.x, aa.x); aa.y = fmaf(ca.w[2 * j + 1], v1.y, aa.y);
        ab.x = fmaf(cb.w[2 * j], u0.x, ab.x); ab.y = fmaf(cb.w[2 * j], u0.y, ab.y);
        ab.x = fmaf(cb.w[2 * j + 1], u1.x, ab.x); ab.y = fmaf(cb.w[2 * j + 1], u1.y, ab.y);
    }
    ea = va ? aa : make_float2(0.f, 0.f);
    eb = vb ? ab : make_float2(0.f, 0.f);
}

// Corner values of the CENTRE point (stencil point 0) on the lane's two coarse levels (t and t + 4), kept across the 7 stencil
// evaluations of a sample: the +-eps points fall into the centre's cell ~85 % (levels 0-3) / ~50 % (levels 4-7) of the time and
// then need no gather at all, only new weights.  The arrays live in local memory if registers run out — a coalesced spill
// access costs the LSU 1/32 of a scattered gather, which is the resource these kernels are bound by.
struct CornerCache {
    uint32_t key[2][2];          // [coarse level slot][row]
    uint32_t val[2][2][8];
};

// pair_issue with an extra predicate: nothing is loaded when the cached corner values can be used
__device__ __forceinline__ void pair_issue_if(PairLoad& r, const __half2* t, uint32_t i0, uint32_t i1, bool need) {
    const uint32_t m = ((i0 ^ i1) == 1u) ? 1u : 0u;
    r.sel = m | ((i0 & 1u) << 1);
    r.lo = r.hi = r.a = r.b = 0u;
    const uint32_t pm = (need && m) ? 1u : 0u, ps = (need && !m) ? 1u : 0u;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %7, 0;\n\tsetp.ne.u32 q, %8, 0;\n\t"
                 "@p ld.global.nc.v2.u32 {%0, %1}, [%4];\n\t"
                 "@q ld.global.nc.u32 %2, [%5];\n\t"
                 "@q ld.global.nc.u32 %3, [%6];\n\t}"
                 : "+r"(r.lo), "+r"(r.hi), "+r"(r.a), "+r"(r.b)
                 : "l"(t + (i0 & ~1u)), "l"(t + i0), "l"(t + i1), "r"(pm), "r"(ps));
}
__device__ __forceinline__ void pair_raw(const PairLoad& r, uint32_t& u0, uint32_t& u1) {
    const bool merged = r.sel & 1u, odd = r.sel & 2u;
    u0 = merged ? (odd ? r.hi : r.lo) : r.a;
    u1 = merged ? (odd ? r.lo : r.hi) : r.b;
}

// encode_level_pair for a coarse level slot with the centre-cell cache: `centre` = this is stencil point 0 (fill the cache).
__device__ __forceinline__ void encode_level_pair_cached(const __half2* __restrict__ table, const LevelSmem& lv, const float pa[3], bool va,
                                                         const float pb[3], bool vb, bool smooth, float2& ea, float2& eb,
                                                         uint32_t ckey[2], uint32_t cval[2][8], bool centre) {
    Corners ca, cb;
    const uint32_t ka = level_corners(ca, lv, pa[0], pa[1], pa[2], smooth);
    const uint32_t kb = level_corners(cb, lv, pb[0], pb[1], pb[2], smooth);
    const bool need_a = centre || ka != ckey[0], need_b = centre || kb != ckey[1];
    const __half2* t = table + lv.offset;
    PairLoad la[4], lb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        pair_issue_if(la[j], t, ca.idx[2 * j], ca.idx[2 * j + 1], need_a);
        pair_issue_if(lb[j], t, cb.idx[2 * j], cb.idx[2 * j + 1], need_b);
    }
    if (centre) { ckey[0] = ka; ckey[1] = kb; }
    float2 aa = make_float2(0.f, 0.f), ab = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t a0, a1, b0, b1;
        pair_raw(la[j], a0, a1);
        pair_raw(lb[j], b0, b1);
        a0 = need_a ? a0 : cval[0][2 * j]; a1 = need_a ? a1 : cval[0][2 * j + 1];
        b0 = need_b ? b0 : cval[1][2 * j]; b1 = need_b ? b1 : cval[1][2 * j + 1];
        if (centre) { cval[0][2 * j] = a0; cval[0][2 * j + 1] = a1; cval[1][2 * j] = b0; cval[1][2 * j + 1] = b1; }
        const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&a0)), v1 = __half22float2(*reinterpret_cast<const __half2*>(&a1));
        const float2 u0 = __half22float2(*reinterpret_cast<const __half2*>(&b0)), u1 = __half22float2(*reinterpret_cast<const __half2*>(&b1));
        aa.x = fmaf(ca.w[2 * j], v0.x, aa.x); aa.y = fmaf(ca.w[2 * j], v0.y, aa.y);
        aa.x = fmaf(ca.w[2 * j + 1], v1.x, aa.x); aa.y = fmaf(ca.w[2 * j + 1], v1.y, aa.y);
        ab.x = fmaf(cb.w[2 * j], u0.x, ab.x); ab.y = fmaf(cb.w[2 * j], u0.y, ab.y);
        ab.x = fmaf(cb.w[2 * j + 1], u1.x, ab.x); ab.y = fmaf(cb.w[2 * j + 1], u1.y, ab.y);
    }
    ea = va ? aa : make_float2(0.f, 0.f);
    eb = vb ? ab : make_float2(0.f, 0.f);
}

// encode_rows with the centre-cell cache on the two coarse level slots (kt = 0); the fine slots (kt = 1) gather as usual.
__device__ __forceinline__ void encode_rows_cached(uint32_t a[2][4], const WeightsSmem& s, const FieldParams& p, int lane,
                                                   const float pa[3], bool va, const float pb[3], bool vb, CornerCache& cache, bool centre) {
    const int t = lane & 3;
    const bool smooth = p.interp_smoothstep != 0;
#pragma unroll
    for (int kt = 0; kt < 2; kt++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t level = kt * 8 + t + h * 4;
            float2 ea = make_float2(0.f, 0.f), eb = make_float2(0.f, 0.f);
            if (level < p.n_levels_active) {
                const LevelSmem lv = s.lv[level];
                if (kt == 0) encode_level_pair_cached(p.table, lv, pa, va, pb, vb, smooth, ea, eb, cache.key[h], cache.val[h], centre);
                else encode_level_pair(p.table, lv, pa, va, pb, vb, smooth, ea, eb);
            }
            a[kt][h * 2 + 0] = pack_half2(ea.x, ea.y);
            a[kt][h * 2 + 1] = pack_half2(eb.x, eb.y);
        }
    }
}

// The encoder of one stencil point for the two rows a lane owns, in A-fragment order:
//   a[kt][0] = (row g,   level 8kt + t)      a[kt][1] = (row g+8, level 8kt + t)
//   a[kt][2] = (row g,   level 8kt + t + 4)  a[kt][3] = (row g+8, level 8kt + t + 4)
// with g = lane>>2, t = lane&3 (two fp16 features per level = one 32-bit register).
__device__ __forceinline__ void encode_rows(uint32_t a[2][4], const WeightsSmem& s, const FieldParams& p, int lane,
                                            const float pa[3], bool va, const float pb[3], bool vb) {
    const int t = lane & 3;
    const bool smooth = p.interp_smoothstep != 0;
#pragma unroll
    for (int kt = 0; kt < 2; kt++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t level = kt * 8 + t + h * 4;
            float2 ea = make_float2(0.f, 0.f), eb = make_float2(0.f, 0.f);
            if (level < p.n_levels_active) {
                const LevelSmem lv = s.lv[level];
                encode_level_pair(p.table, lv, pa, va, pb, vb, smooth, ea, eb);
            }
            a[kt][h * 2 + 0] = pack_half2(ea.x, ea.y);
            a[kt][h * 2 + 1] = pack_half2(eb.x, eb.y);
        }
    }
}

// Split form of encode_rows for software pipelining (fused_field_bwd.cu): gather_issue() computes the corner indices of the
// lane's two FINE levels (8 + t, 12 + t: the hashed ones that miss L1) x 2 rows and issues those 32 gathers into `raw`
// (nothing waits on them); gather_finish() gathers the two coarse levels as usual, recomputes the trilinear weights of the
// fine ones (pure ALU) and folds the values that arrived meanwhile into the A fragments.  (Prefetching all four level
// groups needs 64 live registers and spills under the 128-register budget of the 16-warp CTA.)
__device__ __forceinline__ void gather_issue(__half2 raw[2][16], const WeightsSmem& s, const FieldParams& p, int lane,
                                             const float pa[3], const float pb[3]) {
    const int t = lane & 3;
    const bool smooth = p.interp_smoothstep != 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t level = 8 + t + h * 4;
        if (level < p.n_levels_active) {
            const LevelSmem lv = s.lv[level];
            Corners ca, cb;
            level_corners(ca, lv, pa[0], pa[1], pa[2], smooth);
            level_corners(cb, lv, pb[0], pb[1], pb[2], smooth);
            const __half2* tb = p.table + lv.offset;
#pragma unroll
            for (int k = 0; k < 8; k++) { raw[h][k] = __ldg(tb + ca.idx[k]); raw[h][8 + k] = __ldg(tb + cb.idx[k]); }
        }
    }
}

__device__ __forceinline__ void gather_finish(uint32_t a[2][4], const __half2 raw[2][16], const WeightsSmem& s, const FieldParams& p, int lane,
                                              const float pa[3], bool va, const float pb[3], bool vb) {
    const int t = lane & 3;
    const bool smooth = p.interp_smoothstep != 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {            // coarse levels t, t + 4: gathered now (mostly L1 hits)
        const uint32_t level = t + h * 4;
        float2 ea = make_float2(0.f, 0.f), eb = make_float2(0.f, 0.f);
        if (level < p.n_levels_active) encode_level_pair(p.table, s.lv[level], pa, va, pb, vb, smooth, ea, eb);
        a[0][h * 2 + 0] = pack_half2(ea.x, ea.y);
        a[0][h * 2 + 1] = pack_half2(eb.x, eb.y);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {            // fine levels: values were prefetched
        const uint32_t level = 8 + t + h * 4;
        float2 ea = make_float2(0.f, 0.f), eb = make_float2(0.f, 0.f);
        if (level < p.n_levels_active) {
            const LevelSmem lv = s.lv[level];
            Corners ca, cb;          // only the weights are used: the index arithmetic is dead code here
            level_corners(ca, lv, pa[0], pa[1], pa[2], smooth);
            level_corners(cb, lv, pb[0], pb[1], pb[2], smooth);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float2 v = __half22float2(raw[h][k]), u = __half22float2(raw[h][8 + k]);
                ea.x = fmaf(ca.w[k], v.x, ea.x); ea.y = fmaf(ca.w[k], v.y, ea.y);
                eb.x = fmaf(cb.w[k], u.x, eb.x); eb.y = fmaf(cb.w[k], u.y, eb.y);
            }
            if (!va) ea = make_float2(0.f, 0.f);
            if (!vb) eb = make_float2(0.f, 0.f);
        }
        a[1][h * 2 + 0] = pack_half2(ea.x, ea.y);
        a[1][h * 2 + 1] = pack_half2(eb.x, eb.y);
    }
}

// 32 -> 64 -> 64 -> 4 MLP for a 16-row tile held as A fragments.  Returns the layer-3 accumulator tile
// (cols 0..7; lane holds (row g, cols 2t,2t+1) in c[0],c[1] and (row g+8, ...) in c[2],c[3]).
// When KEEP is set the post-ReLU activations of both hidden layers are returned as A fragments.
template <bool KEEP>
__device__ __forceinline__ void mlp_forward(float out[4], const uint32_t a0[2][4], const WeightsSmem& s, int lane,
                                            uint32_t act1[4][4], uint32_t act2[4][4]) {
    const int t = lane & 3;
    uint32_t a1[4][4];
    // Every layer runs k-step outer / n-tile inner: 8 independent accumulator chains (a chain of dependent mma.sync stalls the
    // warp for the whole MMA latency).  One ldmatrix.x4 feeds two n-tiles of one k-step.
    {   // layer 1: 32 -> 64
        float c[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) { c[nt][0] = c[nt][2] = s.b1[nt * 8 + 2 * t]; c[nt][1] = c[nt][3] = s.b1[nt * 8 + 2 * t + 1]; }
#pragma unroll
        for (int kt = 0; kt < 2; kt++) {
#pragma unroll
            for (int np = 0; np < 4; np++) {
                uint32_t wb[4];
                ldsm_b_nn(wb, &s.w1[0][0], kW1Stride, np * 16, kt * 16, lane);
                mma16816(c[2 * np], a0[kt], wb[0], wb[1]);
                mma16816(c[2 * np + 1], a0[kt], wb[2], wb[3]);
            }
        }
        // fp16 output of the linear layer, ReLU, straight into the next layer's A fragment
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            a1[nt >> 1][(nt & 1) * 2 + 0] = pack_half2(fmaxf(c[nt][0], 0.f), fmaxf(c[nt][1], 0.f));
            a1[nt >> 1][(nt & 1) * 2 + 1] = pack_half2(fmaxf(c[nt][2], 0.f), fmaxf(c[nt][3], 0.f));
        }
    }
    uint32_t a2[4][4];
    {   // layer 2: 64 -> 64
        float c[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) { c[nt][0] = c[nt][2] = s.b2[nt * 8 + 2 * t]; c[nt][1] = c[nt][3] = s.b2[nt * 8 + 2 * t + 1]; }
#pragma unroll
        for (int kt = 0; kt < 4; kt++) {
#pragma unroll
            for (int np = 0; np < 4; np++) {
                uint32_t wb[4];
                ldsm_b_nn(wb, &s.w2[0][0], kW2Stride, np * 16, kt * 16, lane);
                mma16816(c[2 * np], a1[kt], wb[0], wb[1]);
                mma16816(c[2 * np + 1], a1[kt], wb[2], wb[3]);
            }
        }
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            a2[nt >> 1][(nt & 1) * 2 + 0] = pack_half2(fmaxf(c[nt][0], 0.f), fmaxf(c[nt][1], 0.f));
            a2[nt >> 1][(nt & 1) * 2 + 1] = pack_half2(fmaxf(c[nt][2], 0.f), fmaxf(c[nt][3], 0.f));
        }
    }
    // layer 3: 64 -> 4 (one n-tile, rows 4..7 of w3 are zero): a single chain of 4
    out[0] = out[2] = s.b3[2 * t];
    out[1] = out[3] = s.b3[2 * t + 1];
#pragma unroll
    for (int kp = 0; kp < 2; kp++) {
        uint32_t wb[4];
        ldsm_b2(wb, &s.w3[0][0], kW2Stride, kp * 32, lane);
        mma16816(out, a2[2 * kp], wb[0], wb[1]);
        mma16816(out, a2[2 * kp + 1], wb[2], wb[3]);
    }
    if (KEEP) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { act1[i][j] = a1[i][j]; act2[i][j] = a2[i][j]; }
    }
}

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// density_blob with the exp activation (renderer.py:339-349): blob_density * exp(-|x|^2 / (2 r^2))
__device__ __forceinline__ float blob(const FieldParams& p, const float x[3]) {
    const float d = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    return p.blob_density * __expf(-d / (2.f * p.blob_radius * p.blob_radius));
}

}  // namespace field
