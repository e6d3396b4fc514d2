// flash_attn_tc.cu — fused multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   o[b, i, h*d : (h+1)*d] = softmax_j(scale * q[b,i,h] . k[b,j,h]) v[b,j,h]        (ldm/modules/attention.py:170-193)
//
// for the long self-attention layers of the UNet (64x64 latents: 4096 x 4096 scores per head, d = 40; 32x32: d = 80), where the
// mma.sync kernel of flash_attn.cu sits at its legacy-pipe issue limit.  One CTA owns 128 queries of one (batch, head):
//   warp 4    TMA producer: the Q tile once, then K / V tiles of 128 keys through a 2-stage ring.  The tensor maps view q / k / v as
//             (d, heads, tokens, batch) so that a 64-wide box at (0, h, token0, b) is ONE head's slice, zero-filled beyond d (and beyond
//             the last token): no padding or transposition of the projection buffers
//   warp 5    MMA issuer: S = Q K^T (M128 x N128, K = d rounded up to 16) into TMEM columns [0,128); then, when the softmax warps have
//             published P, T = P V (M128 x N = d rounded up to 16, K = 128 keys; V is the MN-major B operand straight from its token-major
//             tile) into TMEM columns [128, 128 + d)
//   warps 0-3 softmax: thread r owns query row r (TMEM lane r) — no cross-thread reductions at all: two passes over S from TMEM (row
//             maximum, then exp2 / row sum / fp16 P written to shared memory in the 128B-swizzled K-major layout the MMA reads), then
//             O_row = alpha * O_row + T_row with O in registers; finally O / l -> fp16 -> global
// Shared memory 112 KB and 256 TMEM columns per CTA: two CTAs per SM, so one CTA's exponentials overlap the other's MMAs.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace {

constexpr int kBM = 128, kBN = 128, kThreads = 192;
constexpr int kTileBytes = 128 * 64 * 2;          // one [128 rows x 64 fp16] swizzled tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FA_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FA_DONE;\n\t"
        "bra FA_WAIT;\n\t"
        "FA_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// wait with a sleep between polls: the producer / issuer warps share their scheduler with a softmax warp, and a hot try_wait loop
// (measured: a third of all instructions the kernel issued) takes issue slots from it
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
    while (!mbar_try(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// [rows x 64 fp16] tile, 128-byte rows, 128B swizzle, 8-row groups 1024 bytes apart.  As a K-major operand (A, or B = K^T) the 64
// columns are the contraction dimension; as the MN-major B operand (V) the rows are the contraction dimension and SBO (1024 B) steps
// from one 8-row group to the next (cute/atom/mma_traits_sm100.hpp: make_umma_desc, canonical layouts).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct FaArgs {
    __half* o;
    int n, nkv, heads, d, ldo;
    float scale_log2;
};

// D16 = d rounded up to a multiple of 16 (<= 64)
template <int D16>
__global__ void __launch_bounds__(kThreads, 2)
k_flash_attn_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                const FaArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = smem;                         // 16 KB
    unsigned char* sK = smem + kTileBytes;            // 2 x 16 KB
    unsigned char* sV = smem + 3 * kTileBytes;        // 2 x 16 KB
    unsigned char* sP = smem + 5 * kTileBytes;        // 2 x 16 KB: keys 0..63 | keys 64..127
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
    uint64_t* q_full = bars;              // 1
    uint64_t* kv_full = bars + 1;         // 2
    uint64_t* kv_empty = bars + 3;        // 2
    uint64_t* s_full = bars + 5;          // S ready in TMEM
    uint64_t* p_full = bars + 6;          // P in smem, S consumed (128 arrivals)
    uint64_t* t_full = bars + 7;          // T ready in TMEM (P and the V stage are free again)
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 8);
    if (threadIdx.x == 0 && reinterpret_cast<unsigned char*>(bars) + 128 > smem_raw + (7 * kTileBytes + 128 + 896)) __trap();   // alignment slack exhausted

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBM, bh = blockIdx.y, b = bh / a.heads, h = bh - b * a.heads;
    const int n_tiles = (a.nkv + kBN - 1) / kBN;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    }
    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; i++) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(t_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, kTileBytes);
            tma_load_4d(&map_q, q_full, sQ, 0, h, q0, b);
            for (int j = 0; j < n_tiles; j++) {
                const int st = j & 1;
                mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
                tma_load_4d(&map_k, &kv_full[st], sK + st * kTileBytes, 0, h, j * kBN, b);
                tma_load_4d(&map_v, &kv_full[st], sV + st * kTileBytes, 0, h, j * kBN, b);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_s = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);                     // K-major A and B
        // B (= V) MN-major; N = the whole 64-wide swizzle atom (columns >= d are zero-filled by the TMA and never read back)
        constexpr uint32_t idesc_t = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        mbar_wait(q_full, 0);
        for (int j = 0; j < n_tiles; j++) {
            const int st = j & 1;
            mbar_wait(&kv_full[st], (j >> 1) & 1);
            // S_j is issued right behind T_{j-1} (different TMEM ranges; S_{j-1} was released by p_full of the previous trip), so it is
            // computed while the softmax warps still fold T_{j-1} into their output rows
            tc_fence_after();
            if (lane == 0) {
                const uint64_t dq = make_desc_sw128(smem_u32(sQ)), dk = make_desc_sw128(smem_u32(sK + st * kTileBytes));
#pragma unroll
                for (int k = 0; k < D16 / 16; k++) umma_f16(tmem_base, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_full, j & 1);                          // P_j written (generic -> async proxy fenced by the writers), S_j consumed
            tc_fence_after();
            if (lane == 0) {
                const uint64_t dv = make_desc_sw128(smem_u32(sV + st * kTileBytes));
#pragma unroll
                for (int k = 0; k < kBN / 16; k++) {
                    // A: P chunk (k / 4), 32 bytes per k-step inside the swizzled row; B: V rows 16 k .. 16 k + 15 = 2048 bytes further
                    const uint64_t dp = make_desc_sw128(smem_u32(sP + (k >> 2) * kTileBytes)) + (uint64_t)((k & 3) * 2);
                    umma_f16(tmem_base + 128, dp, dv + (uint64_t)(k * 128), idesc_t, k > 0 ? 1u : 0u);
                }
                umma_commit(&kv_empty[st]);                    // K_j / V_j free when these MMAs retire
                umma_commit(t_full);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax + output: thread r <-> query row r <-> TMEM lane r =====================
        const int r = warp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        float o_acc[D16];
#pragma unroll
        for (int i = 0; i < D16; i++) o_acc[i] = 0.f;
        float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
        unsigned char* p_row = sP + r * 128;
        const int sw = r & 7;
        for (int j = 0; j < n_tiles; j++) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int key0 = j * kBN;
            const bool ragged = key0 + kBN > a.nkv;
            // pass 1: row maximum.  The next chunk's TMEM read is issued before the current one is consumed (tcgen05.wait::ld waits for
            // every outstanding load, so the order is wait -> issue next -> use current).
            float mx = -INFINITY;
            uint32_t va[32], vb[32];
            tmem_ld32(t_row, va);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                tmem_ld_wait();
                uint32_t (&cur)[32] = (c & 1) ? vb : va;
                uint32_t (&nxt)[32] = (c & 1) ? va : vb;
                tmem_ld32(t_row + ((c + 1) & 3) * 32, nxt);         // chunk c + 1; after the last chunk: chunk 0 again, for pass 2
                if (ragged) {
#pragma unroll
                    for (int i = 0; i < 32; i++) mx = fmaxf(mx, key0 + c * 32 + i < a.nkv ? __uint_as_float(cur[i]) : -INFINITY);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])));
                }
            }
            const float m_new = fmaxf(m, mx);                 // in units of raw scores; exponent = (s - m_new) * scale_log2
            const float alpha = ex2((m - m_new) * a.scale_log2);     // 0 on the first tile (m = -inf)
            const float mb = m_new * a.scale_log2;
            float rs = 0.f;
            // fold the PREVIOUS tile's T = P V into the output row now: its MMA ran while pass 1 above was reading S, so the wait is short;
            // it also guarantees that the MMA has finished reading P before pass 2 overwrites it
            if (j > 0) {
                mbar_wait(t_full, (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < D16; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + 128 + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha_prev, __uint_as_float(v[i]));
                }
            }
            alpha_prev = alpha;
            // pass 2: p = exp2(s * scale_log2 - mb) evaluated two at a time in fp16 (P is fp16 for the MMA anyway: one cvt.f16x2 + one
            // MUFU.EX2.f16x2 per pair instead of two fp32 exponentials and a pack), row sum through short fp16x2 chains folded into fp32,
            // P into the swizzled A-operand layout.  va holds chunk 0 again (issued at the end of pass 1).
#pragma unroll
            for (int c = 0; c < 4; c++) {
                tmem_ld_wait();
                uint32_t (&cur)[32] = (c & 1) ? vb : va;
                uint32_t (&nxt)[32] = (c & 1) ? va : vb;
                if (c < 3) tmem_ld32(t_row + (c + 1) * 32, nxt);
                uint32_t pk[16];
                __half2 acc2[4] = {__float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f)};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    float x0 = fmaf(__uint_as_float(cur[2 * i]), a.scale_log2, -mb);
                    float x1 = fmaf(__uint_as_float(cur[2 * i + 1]), a.scale_log2, -mb);
                    if (ragged) {
                        if (key0 + c * 32 + 2 * i >= a.nkv) x0 = -INFINITY;
                        if (key0 + c * 32 + 2 * i + 1 >= a.nkv) x1 = -INFINITY;
                    }
                    uint32_t h2, p2;
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(x1), "f"(x0));          // hi = x1, lo = x0
                    asm("ex2.approx.f16x2 %0, %1;" : "=r"(p2) : "r"(h2));
                    pk[i] = p2;
                    acc2[i & 3] = __hadd2(acc2[i & 3], *reinterpret_cast<const __half2*>(&p2));
                }
#pragma unroll
                for (int q = 0; q < 4; q++) { const float2 f2 = __half22float2(acc2[q]); rs += f2.x + f2.y; }
                // columns c*32 .. c*32+31 of row r: chunk (c >> 1), 16-byte units u = (c & 1) * 4 .. +3, physical unit = u ^ (r & 7)
                unsigned char* base = p_row + (c >> 1) * kTileBytes;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int unit = ((c & 1) * 4 + u) ^ sw;
                    *reinterpret_cast<uint4*>(base + unit * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
            l = l * alpha + rs;
            m = m_new;
            // publish P to the tensor core (async proxy) and release S
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // the last tile's T
        mbar_wait(t_full, (n_tiles - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < D16; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_row + 128 + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i++) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha_prev, __uint_as_float(v[i]));
        }
        tc_fence_before();
        // ---- normalise and store
        const int qi = q0 + r;
        if (qi < a.n) {
            const float inv = 1.f / l;
            __half* op = a.o + ((size_t)b * a.n + qi) * a.ldo + h * a.d;
#pragma unroll
            for (int i = 0; i < D16; i += 2) {
                if (i < a.d) *reinterpret_cast<__half2*>(op + i) = __floats2half2_rn(o_acc[i] * inv, o_acc[i + 1] * inv);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}


// ------------------------------------------------------------------------------------------------------------------------------------
// Version 2 of the kernel (SDF_FLASH_TC_V=2): same tiles, same operand layouts, a shorter dependency chain per key tile.
//   * a softmax thread pulls its whole 128-score row of S into registers with ONE TMEM wait (four x32 loads in flight) and releases
//     the S columns at once (s_free), so the MMA warp computes S_{j+1} while the exponentials of tile j are still running;
//   * O accumulates in TMEM across key tiles (accumulate flag of the P V product) instead of travelling TMEM -> registers every tile;
//     the running maximum is only moved — and O rescaled in place, tcgen05.ld / multiply / tcgen05.st — when some row of the warp
//     sees its maximum grow by more than 2^8 (P stays <= 256 in fp16; the row sum carries the same stale maximum, so O / l is exact);
//   * K and V have separate two-stage rings: K_j is released by S_j (early), V_j by T_j.
// Exponentials are fp32 MUFU.EX2 (sm_100a's MUFU.EX2.F16 is not packed: the f16x2 form costs two MUFU slots as well), packed once.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr float kRescaleLog2 = 8.f;

template <int D16, int BN, int OCC>
__global__ void __launch_bounds__(kThreads, OCC)
k_flash_attn_tc2(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const FaArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = smem;                         // 16 KB
    constexpr int kKV = BN * 128;                     // one K or V stage: BN keys x 64 fp16
    constexpr int kPT = BN / 64;                      // P: one 16 KB swizzled tile per 64 keys
    constexpr int kTiles = kTileBytes + 4 * kKV + kPT * kTileBytes;
    unsigned char* sK = smem + kTileBytes;            // 2 stages
    unsigned char* sV = sK + 2 * kKV;                 // 2 stages
    unsigned char* sP = sV + 2 * kKV;                 // keys 0..63 | keys 64..127
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTiles);
    uint64_t* q_full = bars;              // 1
    uint64_t* k_full = bars + 1;          // 2
    uint64_t* k_empty = bars + 3;         // 2
    uint64_t* v_full = bars + 5;          // 2
    uint64_t* v_empty = bars + 7;         // 2
    uint64_t* s_full = bars + 9;          // S_j ready in TMEM
    uint64_t* s_free = bars + 10;         // S_j is in the softmax threads' registers (128 arrivals)
    uint64_t* p_full = bars + 11;         // P_j in smem, O rescaled if it had to be (128 arrivals)
    uint64_t* t_full = bars + 12;         // O += P_j V_j retired: P and the V stage are free again
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 13);
    if (threadIdx.x == 0 && reinterpret_cast<unsigned char*>(bars) + 128 > smem_raw + (kTiles + 128 + 896)) __trap();   // alignment slack exhausted

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBM, bh = blockIdx.y, b = bh / a.heads, h = bh - b * a.heads;
    const int n_tiles = (a.nkv + BN - 1) / BN;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    }
    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; i++) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128); mbar_init(t_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, kTileBytes);
            tma_load_4d(&map_q, q_full, sQ, 0, h, q0, b);
            for (int j = 0; j < n_tiles; j++) {
                const int st = j & 1;
                const uint32_t ph = ((j >> 1) & 1) ^ 1;
                mbar_wait_backoff(&k_empty[st], ph, 256);
                mbar_expect_tx(&k_full[st], kKV);
                tma_load_4d(&map_k, &k_full[st], sK + st * kKV, 0, h, j * BN, b);
                mbar_wait_backoff(&v_empty[st], ph, 256);
                mbar_expect_tx(&v_full[st], kKV);
                tma_load_4d(&map_v, &v_full[st], sV + st * kKV, 0, h, j * BN, b);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_s = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);                     // K-major A and B
        constexpr uint32_t idesc_t = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);         // B (= V) MN-major, N = 64
        const uint64_t dq = make_desc_sw128(smem_u32(sQ));
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        if (lane == 0) {
            const uint64_t dk = make_desc_sw128(smem_u32(sK));
#pragma unroll
            for (int k = 0; k < D16 / 16; k++) umma_f16(tmem_base, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&k_empty[0]);
            umma_commit(s_full);
        }
        __syncwarp();
        for (int j = 0; j < n_tiles; j++) {
            const int st = j & 1;
            if (j + 1 < n_tiles) {
                // S_{j+1}: as soon as K_{j+1} has landed and the softmax threads hold S_j in registers
                const int sn = (j + 1) & 1;
                mbar_wait_backoff(&k_full[sn], ((j + 1) >> 1) & 1, 32);
                mbar_wait_backoff(s_free, j & 1, 32);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t dk = make_desc_sw128(smem_u32(sK + sn * kKV));
#pragma unroll
                    for (int k = 0; k < D16 / 16; k++) umma_f16(tmem_base, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
                    umma_commit(&k_empty[sn]);
                    umma_commit(s_full);
                }
                __syncwarp();
            }
            mbar_wait_backoff(&v_full[st], (j >> 1) & 1, 32);
            mbar_wait_backoff(p_full, j & 1, 32);               // P_j written (generic -> async proxy fenced by the writers), O rescaled
            tc_fence_after();
            if (lane == 0) {
                const uint64_t dv = make_desc_sw128(smem_u32(sV + st * kKV));
#pragma unroll
                for (int k = 0; k < BN / 16; k++) {
                    const uint64_t dp = make_desc_sw128(smem_u32(sP + (k >> 2) * kTileBytes)) + (uint64_t)((k & 3) * 2);
                    umma_f16(tmem_base + BN, dp, dv + (uint64_t)(k * 128), idesc_t, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&v_empty[st]);
                umma_commit(t_full);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax: thread r <-> query row r <-> TMEM lane r =====================
        const int r = warp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        float m_used = -INFINITY, l = 0.f;
        unsigned char* p_row = sP + r * 128;
        const int sw = r & 7;
        for (int j = 0; j < n_tiles; j++) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t s[BN / 32][32];
#pragma unroll
            for (int c = 0; c < BN / 32; c++) tmem_ld32(t_row + c * 32, s[c]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);                               // the S columns may be overwritten by S_{j+1}
            const int key0 = j * BN;
            const bool ragged = key0 + BN > a.nkv;
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (ragged) {
#pragma unroll
                for (int c = 0; c < BN / 32; c++)
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        if (key0 + c * 32 + i >= a.nkv) s[c][i] = __float_as_uint(-INFINITY);
                    }
            }
#pragma unroll
            for (int c = 0; c < BN / 32; c++)
#pragma unroll
                for (int i = 0; i < 32; i += 2) mx4[c] = fmaxf(mx4[c], fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
            const float m_new = fmaxf(fmaxf(m_used, fmaxf(mx4[0], mx4[1])), fmaxf(mx4[2], mx4[3]));
            // T_{j-1} retired: O is complete up to tile j-1 and the P buffer may be rewritten
            if (j > 0) { mbar_wait(t_full, (j - 1) & 1); tc_fence_after(); }
            const bool need = (m_new - m_used) * a.scale_log2 > kRescaleLog2;        // first tile: m_used = -inf
            if (__any_sync(0xffffffffu, need)) {
                const float alpha = need ? ex2((m_used - m_new) * a.scale_log2) : 1.f;
                if (need) m_used = m_new;
                l *= alpha;
                if (j > 0) {
#pragma unroll
                    for (int c0 = 0; c0 < D16; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(t_row + BN + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st16(t_row + BN + c0, v);
                    }
                    tmem_st_wait();
                }
            }
            const float mb = m_used * a.scale_log2;
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < BN / 32; c++) {
                uint32_t pk[16];
                __half2 acc2[4] = {__float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f)};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float p0 = ex2(fmaf(__uint_as_float(s[c][2 * i]), a.scale_log2, -mb));
                    const float p1 = ex2(fmaf(__uint_as_float(s[c][2 * i + 1]), a.scale_log2, -mb));
                    uint32_t p2;
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(p1), "f"(p0));          // hi = p1, lo = p0
                    pk[i] = p2;
                    acc2[i & 3] = __hadd2(acc2[i & 3], *reinterpret_cast<const __half2*>(&p2));
                }
#pragma unroll
                for (int q = 0; q < 4; q++) { const float2 f2 = __half22float2(acc2[q]); rs += f2.x + f2.y; }
                // columns c*32 .. c*32+31 of row r: chunk (c >> 1), 16-byte units u = (c & 1) * 4 .. +3, physical unit = u ^ (r & 7)
                unsigned char* base = p_row + (c >> 1) * kTileBytes;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int unit = ((c & 1) * 4 + u) ^ sw;
                    *reinterpret_cast<uint4*>(base + unit * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
            l += rs;
            // publish P (and the rescaled O) to the tensor core
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            mbar_arrive(p_full);
        }
        mbar_wait(t_full, (n_tiles - 1) & 1);
        tc_fence_after();
        // ---- O / l -> fp16 -> global
        const int qi = q0 + r;
        const float inv = 1.f / l;
        __half* op = a.o + ((size_t)b * a.n + qi) * a.ldo + h * a.d;
#pragma unroll
        for (int c0 = 0; c0 < D16; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_row + BN + c0, v);
            tmem_ld_wait();
            if (qi < a.n) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    if (c0 + i < a.d) *reinterpret_cast<__half2*>(op + c0 + i) = __floats2half2_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
    }
}



typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fa() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// (d, heads, tokens, batch) view of a [B, tokens, ld] fp16 projection buffer; box = 64 x 1 x box_rows x 1, zero fill out of bounds
int make_head_map(CUtensorMap* map, const void* base, int d, int heads, int tokens, int B, int ld, int box_rows) {
    PFN_encodeTiled enc = get_encode_fa();
    if (!enc) { sdf_set_error("flash_attention(tc): cuTensorMapEncodeTiled unavailable"); return SDF_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)ld * 2, (cuuint64_t)tokens * ld * 2};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sdf_set_error("flash_attention(tc): cuTensorMapEncodeTiled failed (%d)", (int)r); return SDF_ERR_CUDA; }
    return SDF_OK;
}

// V = 1: k_flash_attn_tc (O in registers, S read twice, 128-key tiles, 2 CTAs / SM)
// V = 2: k_flash_attn_tc2<D16, 128, 2> (O in TMEM, S row in registers, lazy rescale)
// V = 4: k_flash_attn_tc2<D16, 64, 3>: the same with 64-key tiles — 65 KB of shared memory, 128 TMEM columns and <= 112 registers per
//        thread, so THREE CTAs (12 softmax warps) share an SM
template <int V> struct TcCfg { static constexpr int BN = (V == 4) ? 64 : 128, OCC = (V == 4) ? 3 : 2; };
template <int D16, int V>
int launch_tc(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const FaArgs& a, int B, cudaStream_t st) {
    constexpr int BN = TcCfg<V>::BN;
    constexpr int kSmem = kTileBytes + 4 * BN * 128 + (BN / 64) * kTileBytes + 128 + 896;       // tiles + barriers + alignment slack
    static bool attr_set[64] = {false};
    void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, FaArgs);
    if (V == 1) kern = k_flash_attn_tc<D16>; else kern = k_flash_attn_tc2<D16, BN, TcCfg<V>::OCC>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        SDF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.n + kBM - 1) / kBM, B * a.heads); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = sdf_pdl_enabled() ? 1 : 0;
    SDF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, a));
    return SDF_OK;
}

}  // namespace

// tcgen05 path of sdf_flash_attention (flash_attn.cu dispatches here): d <= 64 and a multiple of 8, rows 16-byte aligned.
// returns SDF_ERR_UNSUPPORTED when the shape is not covered (the caller then uses the mma.sync kernel).
int sdf_flash_attention_tc(const void* q, const void* k, const void* v, void* o, int B, int heads, int n, int nkv, int d, int ldq, int ldk, int ldo,
                           float scale, cudaStream_t st) {
    if (d > 64 || d % 8 != 0 || ldq % 8 != 0 || ldk % 8 != 0 || ldo % 2 != 0) return SDF_ERR_UNSUPPORTED;
    // SDF_FLASH_TC_V selects the kernel version (see launch_tc); measured at B2 x 8 heads x 4096 x d40 with tools/bench_attn.py: 1: 181 us, 2: 149 us
    static const int ver = [] { const char* e = getenv("SDF_FLASH_TC_V"); return (e && (e[0] == '1' || e[0] == '2' || e[0] == '4')) ? e[0] - '0' : 2; }();
    const int kv_rows = ver == 4 ? 64 : 128;
    CUtensorMap mq, mk, mv;
    int rc;
    if ((rc = make_head_map(&mq, q, d, heads, n, B, ldq, 128))) return rc;
    if ((rc = make_head_map(&mk, k, d, heads, nkv, B, ldk, kv_rows))) return rc;
    if ((rc = make_head_map(&mv, v, d, heads, nkv, B, ldk, kv_rows))) return rc;
    FaArgs a;
    a.o = (__half*)o; a.n = n; a.nkv = nkv; a.heads = heads; a.d = d; a.ldo = ldo; a.scale_log2 = scale * 1.4426950408889634f;
    const int d16 = (d + 15) / 16 * 16;
#define TC(DD) rc = (ver == 4) ? launch_tc<DD, 4>(mq, mk, mv, a, B, st) : (ver == 2) ? launch_tc<DD, 2>(mq, mk, mv, a, B, st) : launch_tc<DD, 1>(mq, mk, mv, a, B, st)
    switch (d16) {
        case 16: TC(16); break;
        case 32: TC(32); break;
        case 48: TC(48); break;
        default: TC(64); break;
    }
#undef TC
    return rc;
}
