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

// q: [B, n, ldq] head h at columns [h*D, h*D + D); k, v: [B, nkv, ldk]; o: [B, n, ldo].
template <int D>
__global__ void __launch_bounds__(128) k_flash_attn(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                                                    __half* __restrict__ o, int n, int nkv, int heads, int ldq, int ldk, int ldo, float scale_log2e) {
    using C = Cfg<D>;
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sK = reinterpret_cast<__half*>(smem_raw);                 // [2][kBN][kStride]
    __half* sV = sK + 2 * kBN * C::kStride;                           // [2][kBN][kStride]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
    const int q0 = blockIdx.x * kBM;
    const __half* qb = q + ((size_t)b * n) * ldq + h * D;
    const __half* kb = k + ((size_t)b * nkv) * ldk + h * D;
    const __half* vb = v + ((size_t)b * nkv) * ldk + h * D;

    // zero the padding columns [D, kStride) of both stages once (they are read by the last k-step when D % 16 != 0)
    for (int i = tid; i < 2 * 2 * kBN * (C::kStride - D) / 8; i += 128) {
        const int per_row = (C::kStride - D) / 8;
        const int row = i / per_row, c = i % per_row;
        *reinterpret_cast<uint4*>((row < 2 * kBN ? sK : sV - 2 * kBN * C::kStride) + (size_t)row * C::kStride + D + c * 8) = make_uint4(0, 0, 0, 0);
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

    // ---- Q fragments (A operand), loaded straight from global memory
    uint32_t qf[C::kKSteps][4];
    {
        const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ks++) {
            const int c0 = ks * 16 + 2 * t, c1 = c0 + 8;
            qf[ks][0] = (r0 < n && c0 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * ldq + c0) : 0u;
            qf[ks][1] = (r1 < n && c0 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * ldq + c0) : 0u;
            qf[ks][2] = (r0 < n && c1 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * ldq + c1) : 0u;
            qf[ks][3] = (r1 < n && c1 < D) ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * ldq + c1) : 0u;
        }
    }
    float oacc[C::kNTilesO][4];
#pragma unroll
    for (int i = 0; i < C::kNTilesO; i++) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // running max / sum for rows g and g+8

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

        // ---- S = Q K^T  (16 x 64 per warp)
        // k-step outer, key tile inner: 8 independent accumulator chains keep the tensor pipe busy (a single chain of
        // dependent mma.sync stalls the warp for the full MMA latency between instructions)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ks++) {
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const __half* krow = tK + (size_t)(nt * 8 + g) * C::kStride;
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 2 * t);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 2 * t + 8);
                mma16816(s[nt], qf[ks], b0, b1);
            }
        }
        // ---- mask the key tail, online softmax (base-2 exponent with the scale folded in)
        const int kv0 = it * kBN;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const int c = kv0 + nt * 8 + 2 * t;
            if (c >= nkv) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (c + 1 >= nkv) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0 * scale_log2e), mn1 = fmaxf(m1, mx1 * scale_log2e);
        const float corr0 = exp2f(m0 - mn0), corr1 = exp2f(m1 - mn1);     // exp2f(-inf) = 0 on the first tile
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const float p0 = exp2f(s[nt][0] * scale_log2e - mn0), p1 = exp2f(s[nt][1] * scale_log2e - mn0);
            const float p2 = exp2f(s[nt][2] * scale_log2e - mn1), p3 = exp2f(s[nt][3] * scale_log2e - mn1);
            rs0 += p0 + p1; rs1 += p2 + p3;
            const int kt = nt >> 1, hi = (nt & 1) * 2;
            pf[kt][hi + 0] = pack_half2(p0, p1);
            pf[kt][hi + 1] = pack_half2(p2, p3);
        }
        l0 = l0 * corr0 + rs0; l1 = l1 * corr1 + rs1;
#pragma unroll
        for (int i = 0; i < C::kNTilesO; i++) { oacc[i][0] *= corr0; oacc[i][1] *= corr0; oacc[i][2] *= corr1; oacc[i][3] *= corr1; }
        // ---- O += P V   (V tile row-major [key][d]: transposed fragments through ldmatrix)
#pragma unroll
        for (int kt = 0; kt < 4; kt++) {                 // key step outer: kNTilesO independent accumulator chains
#pragma unroll
            for (int i = 0; i < C::kNTilesO; i++) {
                uint32_t b0, b1;
                // lanes 0..7 address keys kt*16 + 0..7, lanes 8..15 keys kt*16 + 8..15 (x2: lanes 16..31 ignored but must be valid)
                const int krow = kt * 16 + (lane & 15);
                ldmatrix_x2_trans(b0, b1, tV + (size_t)krow * C::kStride + i * 8);
                mma16816(oacc[i], pf[kt], b0, b1);
            }
        }
        __syncthreads();      // everyone is done with this stage before it is refilled
    }
    cp_async_wait<0>();
    // ---- finalise: row sums across the lane quad, normalise, store
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __half* ob = o + ((size_t)b * n) * ldo + h * D;
#pragma unroll
    for (int i = 0; i < C::kNTilesO; i++) {
        const int c = i * 8 + 2 * t;
        if (r0 < n) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * ldo + c) = pack_half2(oacc[i][0] * inv0, oacc[i][1] * inv0);
        if (r1 < n) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * ldo + c) = pack_half2(oacc[i][2] * inv1, oacc[i][3] * inv1);
    }
}

template <int D>
int launch_flash(const __half* q, const __half* k, const __half* v, __half* o, int B, int heads, int n, int nkv, int ldq, int ldk, int ldo,
                 float scale, cudaStream_t st) {
    using C = Cfg<D>;
    const int smem = 2 * 2 * kBN * C::kStride * (int)sizeof(__half);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev < 64 && !attr_set[dev]) {
        SDF_CHECK_CUDA(cudaFuncSetAttribute(k_flash_attn<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    dim3 grid((n + kBM - 1) / kBM, B * heads);
    sdf_launch_pdl(k_flash_attn<D>, grid, dim3(128), (size_t)smem, st, q, k, v, o, n, nkv, heads, ldq, ldk, ldo, scale * 1.4426950408889634f);
    return SDF_OK;
}

}  // namespace

// o[b, i, h*d + :] = softmax_j(scale * q[b,i,h] . k[b,j,h]) v[b,j,h]     (fp16, token-major; d in {40, 80, 160})
SDF_API int sdf_flash_attention(const void* q, const void* k, const void* v, void* o, int B, int heads, int n, int nkv, int d,
                                int ldq, int ldk, int ldo, float scale, void* stream) {
    if (B == 0 || n == 0) return SDF_OK;
    SDF_CHECK_ARG(q && k && v && o && nkv > 0, "flash_attention: bad arguments");
    SDF_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0,
                  "flash_attention: q/k/v must be 16-byte aligned with row strides multiple of 8");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (d) {
        case 40: rc = launch_flash<40>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, B, heads, n, nkv, ldq, ldk, ldo, scale, st); break;
        case 80: rc = launch_flash<80>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, B, heads, n, nkv, ldq, ldk, ldo, scale, st); break;
        case 160: rc = launch_flash<160>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, B, heads, n, nkv, ldq, ldk, ldo, scale, st); break;
        case 32: rc = launch_flash<32>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, B, heads, n, nkv, ldq, ldk, ldo, scale, st); break;
        case 64: rc = launch_flash<64>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, B, heads, n, nkv, ldq, ldk, ldo, scale, st); break;
        default: sdf_set_error("flash_attention: head dim %d not instantiated (32, 40, 64, 80, 160)", d); return SDF_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    SDF_CHECK_LAUNCH("flash_attention");
    return SDF_OK;
}
