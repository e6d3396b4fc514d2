// flash_attn.cu — fused multi-head attention forward for the UNet transformer blocks (ldm/modules/attention.py:170-193:
// softmax(q k^T / sqrt(d)) v), no score matrix in HBM.
//
// The tcgen05 plan path (csrc/sd_gemm.cu) materialises S = Q K^T in fp16: 537 MB per 64x64 self-attention layer, written,
// re-read by the softmax, written again and re-read by P V — 7.7 % + ~5 % of a step in the round-1 launch list.  This kernel
// keeps S / P in registers: one CTA (4 warps) owns 64 queries of one (batch, head), streams K / V tiles of 64 keys through
// double-buffered shared memory (cp.async), computes S with mma.sync m16n8k16 (fp16 in, fp32 accumulate), does the online
// softmax on the accumulator fragments, re-uses them as the A operand of P V (register chaining) and rescales O in place.
// q/k/v/o stay in the [tokens, heads*d] layout the projection GEMMs produce, so no transposes or padded copies exist.
// The legacy tensor path is used on purpose: the whole product is ~250 GFLOP per step and bound by exp/shuffle and smem
// traffic, not by MMA issue; a tcgen05/TMEM version is future work.
//
// Roofline: tensor (legacy mma.sync).  FLOPs per launch = 4 * B * heads * n * nkv * d.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int kBM = 64;      // queries per CTA (16 per warp)
constexpr int kBN = 64;      // keys per tile

__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int D>
struct Cfg {
    static constexpr int kKSteps = (D + 15) / 16;          // k-steps of Q K^T (d padded to 16)
    static constexpr int kDPad = kKSteps * 16;
    static constexpr int kStride = (D == 40) ? 56 : D + 8; // halfs; makes the K-fragment LDS bank-conflict free (28 / 44 / 84 words)
    static constexpr int kNTilesO = D / 8;                 // n-tiles of the output
    static_assert(kStride >= kDPad, "row stride must cover the padded head dim");
    static_assert(D % 8 == 0, "head dim must be a multiple of 8");
};

__device__ __forceinline__ void ldsm_x4(uint32_t r[4], const void* smem) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t r[4], const void* smem) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// q: [B, n, ldq] head h at columns [h*D, h*D + D); k, v: [B, nkv, ldk]; o: [B, n, ldo].
// MT = 16-row query tiles per warp: with MT = 2 every K / V fragment read from shared memory feeds two MMAs (the kernel is
// bound by shared-memory bandwidth + issue slots, not by the tensor pipe), at the price of ~190 registers (2 CTAs per SM).
template <int D, int MT>
__global__ void __launch_bounds__(128) k_flash_attn(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                                                    __half* __restrict__ o, int n, int nkv, int heads, int ldq, int ldk, int ldo, float scale_log2e) {
    using C = Cfg<D>;
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sK = reinterpret_cast<__half*>(smem_raw);                 // [2][kBN][kStride]
    __half* sV = sK + 2 * kBN * C::kStride;                           // [2][kBN][kStride]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
    const int q0 = blockIdx.x * (kBM * MT);
    const __half* qb = q + ((size_t)b * n) * ldq + h * D;
    const __half* kb = k + ((size_t)b * nkv) * ldk + h * D;
    const __half* vb = v + ((size_t)b * nkv) * ldk + h * D;

    // zero the padding columns [D, kStride) of both stages of K and V once (read by the last k-step when D % 16 != 0)
    {
        constexpr int per_row = (C::kStride - D) / 8;
        for (int i = tid; i < 4 * kBN * per_row; i += 128) {
            const int row = i / per_row, c = i % per_row;
            *reinterpret_cast<uint4*>(sK + (size_t)row * C::kStride + D + c * 8) = make_uint4(0, 0, 0, 0);
        }
    }

    auto load_tile = [&](int stage, int kv0) {
        constexpr int chunks = D / 8;                      // 16-byte chunks per row
        for (int i = tid; i < kBN * chunks; i += 128) {
            const int row = i / chunks, c = i % chunks;
            __half* dk = sK + ((size_t)stage * kBN + row) * C::kStride + c * 8;
            __half* dv = sV + ((size_t)stage * kBN + row) * C::kStride + c * 8;
            if (kv0 + row < nkv) {
                cp_async16(dk, kb + (size_t)(kv0 + row) * ldk + c * 8);
                cp_async16(dv, vb + (size_t)(kv0 + row) * ldk + c * 8);
            } else {
                *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
            }
        }
    };

    // ---- Q fragments (A operand), loaded straight from global memory; query tile mt of this warp starts at row_base(mt)
    auto row_base = [&](int mt) { return q0 + (warp * MT + mt) * 16; };
    uint32_t qf[MT][C::kKSteps][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        const int r0 = row_base(mt) + g, r1 = r0 + 8;
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ks++) {
            const int c0 = ks * 16 + 2 * t, c1 = c0 + 8;
            qf[mt][ks][0] = (r0 < n && c0 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * ldq + c0) : 0u;
            qf[mt][ks][1] = (r1 < n && c0 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * ldq + c0) : 0u;
            qf[mt][ks][2] = (r0 < n && c1 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * ldq + c1) : 0u;
            qf[mt][ks][3] = (r1 < n && c1 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * ldq + c1) : 0u;
        }
    }
    float oacc[MT][C::kNTilesO][4];
    float mrow[MT][2], lrow[MT][2];                                 // running max / sum for rows g and g+8 of each query tile
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
        for (int i = 0; i < C::kNTilesO; i++) oacc[mt][i][0] = oacc[mt][i][1] = oacc[mt][i][2] = oacc[mt][i][3] = 0.f;
        mrow[mt][0] = mrow[mt][1] = -INFINITY; lrow[mt][0] = lrow[mt][1] = 0.f;
    }

    const int n_tiles = (nkv + kBN - 1) / kBN;
    load_tile(0, 0);
    cp_async_commit();
    for (int it = 0; it < n_tiles; it++) {
        const int stage = it & 1;
        if (it + 1 < n_tiles) load_tile(stage ^ 1, (it + 1) * kBN);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const __half* tK = sK + (size_t)stage * kBN * C::kStride;
        const __half* tV = sV + (size_t)stage * kBN * C::kStride;

        // ---- S = Q K^T  (MT x 16 x 64 per warp): k-step outer, key-tile pairs inner; one ldmatrix.x4 = the B fragments of two
        //      key tiles, shared by the MT query tiles
        float s[MT][8][4];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ks++) {
#pragma unroll
            for (int np = 0; np < 4; np++) {
                uint32_t kf[4];
                ldsm_x4(kf, tK + (size_t)(np * 16 + (lane & 7) + (lane >> 4) * 8) * C::kStride + ks * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    mma16816(s[mt][2 * np], qf[mt][ks], kf[0], kf[1]);
                    mma16816(s[mt][2 * np + 1], qf[mt][ks], kf[2], kf[3]);
                }
            }
        }
        // ---- mask the key tail, online softmax (base-2 exponent with the scale folded in), P as A fragments
        const int kv0 = it * kBN;
        const bool partial_tile = kv0 + kBN > nkv;
        uint32_t pf[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                if (partial_tile) {                      // uniform: only the last key tile can be ragged
                    const int c = kv0 + nt * 8 + 2 * t;
                    if (c >= nkv) { s[mt][nt][0] = -INFINITY; s[mt][nt][2] = -INFINITY; }
                    if (c + 1 >= nkv) { s[mt][nt][1] = -INFINITY; s[mt][nt][3] = -INFINITY; }
                }
                mx0 = fmaxf(mx0, fmaxf(s[mt][nt][0], s[mt][nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[mt][nt][2], s[mt][nt][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(mrow[mt][0], mx0 * scale_log2e), mn1 = fmaxf(mrow[mt][1], mx1 * scale_log2e);
            const float corr0 = exp2f(mrow[mt][0] - mn0), corr1 = exp2f(mrow[mt][1] - mn1);     // exp2f(-inf) = 0 on the first tile
            mrow[mt][0] = mn0; mrow[mt][1] = mn1;
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const float p0 = exp2f(fmaf(s[mt][nt][0], scale_log2e, -mn0)), p1 = exp2f(fmaf(s[mt][nt][1], scale_log2e, -mn0));
                const float p2 = exp2f(fmaf(s[mt][nt][2], scale_log2e, -mn1)), p3 = exp2f(fmaf(s[mt][nt][3], scale_log2e, -mn1));
                rs0 += p0 + p1; rs1 += p2 + p3;
                const int kt = nt >> 1, hi = (nt & 1) * 2;
                pf[mt][kt][hi + 0] = pack_half2(p0, p1);
                pf[mt][kt][hi + 1] = pack_half2(p2, p3);
            }
            lrow[mt][0] = lrow[mt][0] * corr0 + rs0; lrow[mt][1] = lrow[mt][1] * corr1 + rs1;
#pragma unroll
            for (int i = 0; i < C::kNTilesO; i++) { oacc[mt][i][0] *= corr0; oacc[mt][i][1] *= corr0; oacc[mt][i][2] *= corr1; oacc[mt][i][3] *= corr1; }
        }
        // ---- O += P V   (V tile row-major [key][d]: transposed fragments through ldmatrix; x4 = two d-tiles of one key step)
#pragma unroll
        for (int kt = 0; kt < 4; kt++) {
#pragma unroll
            for (int ip = 0; ip < C::kNTilesO / 2; ip++) {
                uint32_t vf[4];
                ldsm_x4_trans(vf, tV + (size_t)(kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::kStride + ip * 16 + (lane >> 4) * 8);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    mma16816(oacc[mt][2 * ip], pf[mt][kt], vf[0], vf[1]);
                    mma16816(oacc[mt][2 * ip + 1], pf[mt][kt], vf[2], vf[3]);
                }
            }
            if (C::kNTilesO & 1) {
                constexpr int i = C::kNTilesO - 1;
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, tV + (size_t)(kt * 16 + (lane & 15)) * C::kStride + i * 8);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) mma16816(oacc[mt][i], pf[mt][kt], b0, b1);
            }
        }
        __syncthreads();      // everyone is done with this stage before it is refilled
    }
    cp_async_wait<0>();
    // ---- finalise: row sums across the lane quad, normalise, store
    __half* ob = o + ((size_t)b * n) * ldo + h * D;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
        float l0 = lrow[mt][0], l1 = lrow[mt][1];
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.f / l0, inv1 = 1.f / l1;
        const int r0 = row_base(mt) + g, r1 = r0 + 8;
#pragma unroll
        for (int i = 0; i < C::kNTilesO; i++) {
            const int c = i * 8 + 2 * t;
            if (r0 < n) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * ldo + c) = pack_half2(oacc[mt][i][0] * inv0, oacc[mt][i][1] * inv0);
            if (r1 < n) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * ldo + c) = pack_half2(oacc[mt][i][2] * inv1, oacc[mt][i][3] * inv1);
        }
    }
}

template <int D, int MT>
int launch_flash(const __half* q, const __half* k, const __half* v, __half* o, int B, int heads, int n, int nkv, int ldq, int ldk, int ldo,
                 float scale, cudaStream_t st) {
    using C = Cfg<D>;
    const int smem = 2 * 2 * kBN * C::kStride * (int)sizeof(__half);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev < 64 && !attr_set[dev]) {
        SDF_CHECK_CUDA(cudaFuncSetAttribute(k_flash_attn<D, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    dim3 grid((n + kBM * MT - 1) / (kBM * MT), B * heads);
    sdf_launch_pdl(k_flash_attn<D, MT>, grid, dim3(128), (size_t)smem, st, q, k, v, o, n, nkv, heads, ldq, ldk, ldo, scale * 1.4426950408889634f);
    return SDF_OK;
}

}  // namespace

int sdf_flash_attention_tc(const void* q, const void* k, const void* v, void* o, int B, int heads, int n, int nkv, int d, int ldq, int ldk, int ldo,
                           float scale, cudaStream_t st);      // flash_attn_tc.cu

// o[b, i, h*d + :] = softmax_j(scale * q[b,i,h] . k[b,j,h]) v[b,j,h]     (fp16, token-major; d in {40, 80, 160})
SDF_API int sdf_flash_attention(const void* q, const void* k, const void* v, void* o, int B, int heads, int n, int nkv, int d,
                                int ldq, int ldk, int ldo, float scale, void* stream) {
    if (B == 0 || n == 0) return SDF_OK;
    SDF_CHECK_ARG(q && k && v && o && nkv > 0, "flash_attention: bad arguments");
    SDF_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0,
                  "flash_attention: q/k/v must be 16-byte aligned with row strides multiple of 8");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    // long self-attention (the 64x64-latent layers: 4096 x 4096 scores per head) runs on tcgen05 / TMEM / TMA (flash_attn_tc.cu); short
    // sequences, cross-attention and d > 64 stay on the mma.sync kernel below.  SDF_FLASH_TC=0 disables the tcgen05 path.
    static const bool allow_tc = [] { const char* e = getenv("SDF_FLASH_TC"); return !(e && e[0] == '0'); }();
    if (allow_tc && d <= 64 && n >= 512 && nkv >= 512) {
        rc = sdf_flash_attention_tc(q, k, v, o, B, heads, n, nkv, d, ldq, ldk, ldo, scale, st);
        if (rc == SDF_OK) { SDF_CHECK_LAUNCH("flash_attention(tcgen05)"); return SDF_OK; }
        if (rc != SDF_ERR_UNSUPPORTED) return rc;
    }
#define FLASH(DD, MM) rc = launch_flash<DD, MM>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, B, heads, n, nkv, ldq, ldk, ldo, scale, st)
    // two query tiles per warp where the grid stays large enough to fill the GPU (the 64x64 / 32x32 self-attention layers)
    static const bool allow_mt2 = [] { const char* e = getenv("SDF_FLASH_MT2"); return !(e && e[0] == '0'); }();
    const bool mt2 = allow_mt2 && (long long)B * heads * ((n + 127) / 128) >= 2 * sdf_num_sms();
    switch (d) {
        case 40: if (mt2) FLASH(40, 2); else FLASH(40, 1); break;
        case 80: FLASH(80, 1); break;                      // two query tiles need 255 registers at d = 80: not worth it
        case 160: FLASH(160, 1); break;
        case 32: FLASH(32, 1); break;
        case 64: FLASH(64, 1); break;
        default: sdf_set_error("flash_attention: head dim %d not instantiated (32, 40, 64, 80, 160)", d); return SDF_ERR_UNSUPPORTED;
    }
#undef FLASH
    if (rc) return rc;
    SDF_CHECK_LAUNCH("flash_attention");
    return SDF_OK;
}
