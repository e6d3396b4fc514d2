// sd_gemm.cu — tcgen05 implicit-GEMM convolution / linear kernel for the SD-1.5-shaped UNet and VAE encoder.
//
// One persistent, warp-specialised kernel serves every dense contraction on the guidance side
// (guidance/sd_utils.py:95-108 -> UNet ResBlock / attention projections / GEGLU, VAE conv + data-gradient):
//
//     D[m, n] = epilogue( sum_{tap, c}  A[pixel(m) + offset(tap), c] * Wt[n, tap*Cin + c] )
//
//   * activations are NHWC fp16; the A tile (128 output pixels x 64 channels) of every filter tap is fetched by
//     ONE 4-D TMA box load at the tap-shifted coordinate — image borders are zero-filled by the TMA unit, so a 3x3
//     convolution is 9 x (Cin/64) accumulating MMAs with no im2col buffer; a linear layer is the same with 1 tap;
//   * weights are packed [Cout, taps*Cin] fp16 (K-major) and fetched with 2-D TMA; both operands land in
//     128B-swizzled shared memory and feed tcgen05.mma (cta_group::1, M=128, N=BLOCK_N, K=16) through UMMA descriptors;
//   * accumulators live in TMEM (double-buffered, 2 x BLOCK_N columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1; epilogue = tcgen05.ld -> (+bias) (+per-image embedding) (+residual) (SiLU/GELU) -> fp16 store;
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quarter, 32-column chunks with
//     tcgen05.ld x32, 16-byte addend loads, 32-byte stores, optional GEGLU); 6-8-stage mbarrier pipeline sized to the smem budget;
//   * PAIR = true: 2-CTA clusters issue tcgen05.mma.cta_group::2 (M = 256 per MMA); each CTA stages its own 128 rows of A and
//     HALF of the weight tile, both CTAs' TMA loads complete on the leader's barrier, the leader's commits are multicast;
//   * split-K for the low-resolution UNet levels (M = 128..512): partials are reduced with red.global.add.v4.f32 into a
//     workspace and finished by a tiny epilogue kernel;
//   * programmatic dependent launch: the prologue (barrier init, TMEM allocation, tensor-map prefetch) runs before
//     griddepcontrol.wait, i.e. under the previous kernel's tail.
//
// Roofline: tensor pipe.  FLOPs per launch = 2 * M * N * taps * Cin.
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <mutex>
#include <vector>

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 fp16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 320;     // warp 0 producer, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kEpiWarps = 8;

enum Act { kActNone = 0, kActSilu = 1, kActGelu = 2, kActGeglu = 3 };    // GEGLU: columns come in 32-wide chunks [16 x | 16 gate], out = x * gelu(gate), N/2 wide

struct GemmArgs {
    int M, N, Cin, taps;             // Cin = K elements iterated per tap (multiple of 64; the TMA zero-fills past the real extent)
    int H, W, Nimg;                  // A-side geometry: x (pixels / tokens), y (rows / heads), img (images / batch)
    int tw, th, tn;                  // tile rectangle, tw*th*tn <= 128 rows
    int pad;                         // 1 for 3x3 (tap offsets -pad_lo .. 2 - pad_lo), 0 for 1x1
    int stride, pad_lo;              // strided 3x3: input pixel = output pixel * stride + tap - pad_lo (element-strided TMA box)
    int splitk;                      // >= 1
    int w_by, w_bimg;                // B operand indexed by the tile's y / img coordinate (batched products: attention)
    const float* bias;               // [N] fp32 or null
    const __half* temb; int temb_ld; // [Nimg, N] or null
    const __half* residual; long long r_sx, r_sy, r_simg;   // element strides of the residual, or null
    __half* out; long long o_sx, o_sy, o_simg;               // element strides of the output
    float* workspace;                // [M, N] fp32 when splitk > 1
    int act;
    float alpha;                     // scale applied to the accumulator before bias
    // GroupNorm statistics of up to two CONSUMERS of this output, accumulated by the epilogue (sum and sum of squares per (image, group),
    // fp32 atomics into [Nimg, 32, 2]): the consumer's separate statistics pass (a full re-read of the tensor) disappears
    float* gn_stats[2];
    int gn_cpg[2];                   // channels per group of the consumer's GroupNorm
    int gn_coff[2];                  // channel of the consumer's tensor that this product's column 0 lands on
};

// ------------------------------------------------------------------ PTX wrappers
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
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster share one M=256 MMA; rank 0 issues it.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier of the pair's leader CTA (address in the shared::cluster window)
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on the same barrier offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t v[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the fp16 output's resolution): one MUFU.RCP, one MUFU.EX2 and
// six FMAs instead of erff's branchy ~30-instruction sequence — the GEGLU projections (K = 320 / 640: five k-blocks per tile) are bound by
// their epilogue, where the gate's GELU was two thirds of the instructions.  Compile with -DSDF_EXACT_ERF to restore erff.
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float r = fmaf(-p * t, __expf(-ax * ax), 1.f);
    return copysignf(r, x);
}
__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == kActSilu) return v / (1.f + __expf(-v));
#ifdef SDF_EXACT_ERF
    if (act == kActGelu) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
#else
    if (act == kActGelu) return 0.5f * v * (1.f + erf_as(v * 0.70710678118654752f));
#endif
    return v;
}

constexpr int kGnMaxImg = 4;          // images a statistics-carrying plan may span (UNet: 2 CFG halves x views, VAE: views)

// v[c] = value of column c in this lane's row.  After the call v[0] of lane l holds the sum of column l over the 32 lanes: each step
// trades half of the remaining columns with the partner lane and adds the half it keeps (16 + 8 + 4 + 2 + 1 shuffles).
template <int HALF>
__device__ __forceinline__ void column_sums_step(float (&v)[32], int lane) {
    const bool up = (lane & HALF) != 0;
#pragma unroll
    for (int i = 0; i < HALF; i++) {
        const float send = up ? v[i] : v[i + HALF];
        const float keep = up ? v[i + HALF] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
    }
}
__device__ __forceinline__ void column_sums(float (&v)[32], int lane) {
    column_sums_step<16>(v, lane); column_sums_step<8>(v, lane); column_sums_step<4>(v, lane);
    column_sums_step<2>(v, lane); column_sums_step<1>(v, lane);
}

template <int BLOCK_N, bool PAIR>
struct SmemLayout {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBRows = PAIR ? BLOCK_N / 2 : BLOCK_N;      // a CTA pair splits the B tile between its two CTAs
    static constexpr int kBBytes = kBRows * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kMaxStages = (220 * 1024) / kStageBytes;
    static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
    static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16 + 2 * kGnMaxImg * 32 * 2 * 4;      // + the GroupNorm accumulators (STATS)
    static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;   // + alignment slack
    static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte swizzle-atom alignment");
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

template <int BLOCK_N, bool PAIR, bool STATS>
__global__ void __launch_bounds__(kNumThreads, 1)      // 10 warps -> 3 on one SM sub-partition -> 168 registers per thread at most
k_gemm(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
    using L = SmemLayout<BLOCK_N, PAIR>;
    constexpr int kStages = L::kStages;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* gn_acc = reinterpret_cast<float*>(tmem_base_smem + 4);        // [2 consumers][kGnMaxImg][32 groups][sum, sum of squares]
    if (STATS) {
        for (int i = threadIdx.x; i < 2 * kGnMaxImg * 32 * 2; i += blockDim.x) gn_acc[i] = 0.f;
    }

    pdl_trigger();        // the next kernel of the stream may be scheduled (it blocks in its own griddepcontrol.wait)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    // PAIR: rank 0 of the 2-CTA cluster is the leader (issues the M=256 MMAs, owns the full / tmem_empty barriers that count
    // both CTAs); every CTA owns its empty / tmem_full barriers, which the leader's commits reach by multicast.
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], PAIR ? 2 * kEpiWarps : kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();           // everything above overlapped the previous kernel's tail; its results are visible from here on

    // ---- tile bookkeeping (identical in every role)
    const int tiles_x = (g.W + g.tw - 1) / g.tw, tiles_y = (g.H + g.th - 1) / g.th, tiles_n_img = (g.Nimg + g.tn - 1) / g.tn;
    const int m_tiles = tiles_x * tiles_y * tiles_n_img;
    const int n_tiles = (g.N + BLOCK_N - 1) / BLOCK_N;
    const int kb_per_tap = g.Cin / kBlockK;
    const int kb_total = g.taps * kb_per_tap;
    const int kb_per_split = (kb_total + g.splitk - 1) / g.splitk;
    // PAIR: a work item is a pair of consecutive M tiles (this CTA takes 2*pair + rank; a tile past the end is all zero fill)
    const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;
    const int total_work = m_units * n_tiles * g.splitk;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = worker; w < total_work; w += n_workers) {
                const int split = w % g.splitk; const int t2 = w / g.splitk;
                const int nt = t2 % n_tiles, mt = PAIR ? 2 * (t2 / n_tiles) + (int)cta_rank : t2 / n_tiles;
                const int tx = mt % tiles_x, ty = (mt / tiles_x) % tiles_y, tz = mt / (tiles_x * tiles_y);
                const int x0 = tx * g.tw, y0 = ty * g.th, i0 = tz * g.tn, n0 = nt * BLOCK_N + (PAIR ? (int)cta_rank * L::kBRows : 0);
                const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
                const uint32_t a_bytes = (uint32_t)(g.tw * g.th * g.tn * kBlockK * 2);
                for (int kb = kb0; kb < kb1; kb++) {
                    const int tap = kb / kb_per_tap, cb = kb - tap * kb_per_tap;
                    const int dy = g.pad ? tap / 3 - g.pad_lo : 0, dx = g.pad ? tap % 3 - g.pad_lo : 0;
                    const int xa = x0 * g.stride + dx, ya = y0 * g.stride + dy;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * L::kStageBytes;
                    if (PAIR) {
                        // both CTAs' loads complete on the leader's barrier, which expects the bytes of the whole pair
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (a_bytes + (uint32_t)L::kBBytes));
                        const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
                        tma_load_4d_pair(&map_a, bar, sa, cb * kBlockK, xa, ya, i0);
                        tma_load_4d_pair(&map_b, bar, sa + L::kABytes, kb * kBlockK, n0, 0, 0);
                    } else {
                        mbar_expect_tx(&full_bar[stage], a_bytes + L::kBBytes);
                        tma_load_4d(&map_a, &full_bar[stage], sa, cb * kBlockK, xa, ya, i0);
                        tma_load_4d(&map_b, &full_bar[stage], sa + L::kABytes, kb * kBlockK, n0, g.w_by ? y0 : 0, g.w_bimg ? i0 : 0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
      if (cta_rank == 0) {
        // ===================== MMA issuer (leader CTA only in PAIR mode) =====================
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)((PAIR ? 2 * kBlockM : kBlockM) >> 4) << 24);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int w = worker; w < total_work; w += n_workers) {
            const int split = w % g.splitk;
            const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + L::kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; k++) {
                        // advance 32 bytes (16 fp16) inside the swizzled row: +2 in 16-byte units
                        if (PAIR) umma_f16_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    // frees the smem slot when the MMAs retire; on the last k-block the accumulator is ready for the epilogue
                    if (PAIR) { umma_commit_pair(&empty_bar[stage]); if (kb == kb1 - 1) umma_commit_pair(&tmem_full[acc]); }
                    else { umma_commit(&empty_bar[stage]); if (kb == kb1 - 1) umma_commit(&tmem_full[acc]); }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    } else {
        // ===================== epilogue (8 warps: TMEM lane quarter q = warp % 4, two warps interleave 32-column chunks) ==========
        const int q = warp & 3;                      // TMEM lanes 32q .. 32q+31
        const int half = (warp - 2) >> 2;            // 0 / 1: even / odd chunks
        const int r = q * 32 + lane;                 // row of the tile
        constexpr int kChunks = BLOCK_N / 32;        // 2, 4 or 5
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = worker; w < total_work; w += n_workers) {
            const int t2 = w / g.splitk;
            const int nt = t2 % n_tiles, mt = PAIR ? 2 * (t2 / n_tiles) + (int)cta_rank : t2 / n_tiles;
            const int tx = mt % tiles_x, ty = (mt / tiles_x) % tiles_y, tz = mt / (tiles_x * tiles_y);
            const int lx = r % g.tw, ly = (r / g.tw) % g.th, li = r / (g.tw * g.th);
            const int x = tx * g.tw + lx, y = ty * g.th + ly, img = tz * g.tn + li;
            const bool row_ok = (r < g.tw * g.th * g.tn) && (x < g.W) && (y < g.H) && (img < g.Nimg);
            const long long m = ((long long)img * g.H + y) * g.W + x;
            const long long o_off = (long long)img * g.o_simg + (long long)y * g.o_sy + (long long)x * g.o_sx;
            const int n0 = nt * BLOCK_N;

            // Addends of a full 32-column chunk (bias fp32, per-image embedding fp16, residual fp16) are fetched with 16-byte loads
            // BEFORE the accumulator is waited for, so their latency hides behind the main loop / the TMEM read.
            const __half* res_row = g.residual ? g.residual + (long long)img * g.r_simg + (long long)y * g.r_sy + (long long)x * g.r_sx : nullptr;
            const __half* temb_row = g.temb ? g.temb + (long long)img * g.temb_ld : nullptr;
            __half* out_row = g.out + o_off;
            const bool direct = (g.splitk == 1);
            const bool vec_ok = direct && row_ok && ((g.N & 7) == 0);
            const bool st32_ok = ((reinterpret_cast<uintptr_t>(out_row) & 31) == 0);
            uint4 cr[4];
            auto fetch_addends = [&](int nbase) {      // the residual row segment of a full chunk (64 bytes)
                if (!(vec_ok && res_row && nbase + 32 <= g.N)) return;
#pragma unroll
                for (int j = 0; j < 4; j++) cr[j] = *(reinterpret_cast<const uint4*>(res_row + nbase) + j);
            };
            if (half < kChunks) fetch_addends(n0 + half * 32);

            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
            uint32_t va[32], vb[32];
            if (half < kChunks) { tmem_ld32_nowait(taddr + half * 32, va); tmem_ld_wait(); }
            auto process_chunk = [&](const uint32_t (&cur)[32], int ch) {
                const int nbase = n0 + ch * 32;
                float f[32];
                bool have = false;
                if (row_ok && nbase < g.N) {
                    if (vec_ok && nbase + 32 <= g.N) {
                        // ---- fast path: whole chunk, 16-byte aligned operands
                        have = true;
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = __uint_as_float(cur[j]) * g.alpha;
                        if (g.bias) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + nbase) + j);
                                f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                            }
                        }
                        if (temb_row) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(temb_row + nbase) + j);
                                const __half2* h = reinterpret_cast<const __half2*>(&t4);
#pragma unroll
                                for (int k = 0; k < 4; k++) { const float2 t2 = __half22float2(h[k]); f[8 * j + 2 * k] += t2.x; f[8 * j + 2 * k + 1] += t2.y; }
                            }
                        }
                        if (res_row) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const __half2* h = reinterpret_cast<const __half2*>(&cr[j]);
#pragma unroll
                                for (int k = 0; k < 4; k++) { const float2 t2 = __half22float2(h[k]); f[8 * j + 2 * k] += t2.x; f[8 * j + 2 * k + 1] += t2.y; }
                            }
                        }
                        if (g.act == kActGeglu) {
                            // ldm/modules/attention.py:37-45 fused into the projection: the weight rows were interleaved at pack time so
                            // that this chunk holds 16 values and their 16 gates; the 16 products go to columns nbase/2 .. nbase/2 + 15
                            uint32_t pg[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const __half2 h2 = __floats2half2_rn(f[2 * j] * act_apply(f[16 + 2 * j], kActGelu), f[2 * j + 1] * act_apply(f[16 + 2 * j + 1], kActGelu));
                                pg[j] = *reinterpret_cast<const uint32_t*>(&h2);
                            }
                            __half* og = out_row + (nbase >> 1);
                            if (st32_ok) {
                                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(og), "r"(pg[0]), "r"(pg[1]), "r"(pg[2]), "r"(pg[3]),
                                             "r"(pg[4]), "r"(pg[5]), "r"(pg[6]), "r"(pg[7]) : "memory");
                            } else {
                                *reinterpret_cast<uint4*>(og) = make_uint4(pg[0], pg[1], pg[2], pg[3]);
                                *reinterpret_cast<uint4*>(og + 8) = make_uint4(pg[4], pg[5], pg[6], pg[7]);
                            }
                            if (ch + 2 < kChunks) fetch_addends(n0 + (ch + 2) * 32);
                            return;
                        }
                        if (g.act != kActNone) {
#pragma unroll
                            for (int j = 0; j < 32; j++) f[j] = act_apply(f[j], g.act);
                        }
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const __half2 h2 = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                            pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
                            if (STATS) {            // what the consumer will read (GroupNorm statistics below)
                                const float2 r2 = __half22float2(h2);
                                f[2 * j] = r2.x; f[2 * j + 1] = r2.y;
                            }
                        }
                        __half* o = out_row + nbase;
                        if (st32_ok) {       // two full 32-byte sectors per thread
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                                         "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 16), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]),
                                         "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15]) : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(o + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        }
                        if (ch + 2 < kChunks) fetch_addends(n0 + (ch + 2) * 32);      // next chunk's addends fly during the TMEM wait
                    } else {
                        have = false;
                        // ---- general path: split-K partial sums, ragged N, unaligned rows
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {
                            const int n = nbase + hh * 16;
                            if (n >= g.N) continue;
                            float f[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) f[j] = __uint_as_float(cur[hh * 16 + j]) * g.alpha;
                            if (!direct) {
                                float* ws = g.workspace + m * g.N + n;
                                if (n + 16 <= g.N && ((g.N & 3) == 0)) {
#pragma unroll
                                    for (int j = 0; j < 4; j++) atomicAdd(reinterpret_cast<float4*>(ws) + j, make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; j++) if (n + j < g.N) atomicAdd(ws + j, f[j]);
                                }
                                    continue;
                            }
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                if (n + j < g.N) {
                                    if (g.bias) f[j] += __ldg(g.bias + n + j);
                                    if (temb_row) f[j] += __half2float(temb_row[n + j]);
                                    if (res_row) f[j] += __half2float(res_row[n + j]);
                                    out_row[n + j] = __float2half_rn(act_apply(f[j], g.act));
                                }
                            }
                        }
                    }
                }
                // ---- GroupNorm statistics for the consumer(s) of this output (plans that request it are direct, N % 32 == 0, and their
                //      32-row quarters never straddle an image).  Column sums over the warp's 32 rows by a halving butterfly (31 shuffles
                //      for 32 columns: after the 5 steps lane l holds column l), then one shared-memory add per lane into the CTA's
                //      (image, group) accumulators; the CTA flushes them with global atomics once, after its last tile.
                if (STATS && nbase + 32 <= g.N) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) { f[j] = have ? f[j] : 0.f; sq[j] = f[j] * f[j]; }
                    column_sums(f, lane);
                    column_sums(sq, lane);
                    const uint32_t vm = __ballot_sync(0xffffffffu, have);
                    if (vm) {
                        const int img0 = __shfl_sync(0xffffffffu, img, __ffs(vm) - 1);
#pragma unroll
                        for (int slot = 0; slot < 2; slot++) {
                            if (!g.gn_stats[slot]) continue;
                            const int grp = (g.gn_coff[slot] + nbase + lane) / g.gn_cpg[slot];
                            float* acc_p = gn_acc + ((slot * kGnMaxImg + img0) * 32 + grp) * 2;
                            atomicAdd(acc_p, f[0]);
                            atomicAdd(acc_p + 1, sq[0]);
                        }
                    }
                }
            };
            // two chunks per trip so that the current / prefetched register sets are compile-time (no local-memory arrays):
            // the next chunk's TMEM read is in flight while the current one is finished and stored
#pragma unroll 1
            for (int ch = half; ch < kChunks; ch += 4) {
                if (ch + 2 < kChunks) tmem_ld32_nowait(taddr + (ch + 2) * 32, vb);
                process_chunk(va, ch);
                if (ch + 2 < kChunks) {
                    tmem_ld_wait();
                    if (ch + 4 < kChunks) tmem_ld32_nowait(taddr + (ch + 4) * 32, va);
                    process_chunk(vb, ch + 2);
                    if (ch + 4 < kChunks) tmem_ld_wait();
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));      // the leader's MMA warp waits for both CTAs
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (STATS) {
            // all 8 epilogue warps have added their last tile: flush the CTA's (image, group) partial sums, one global atomic each
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int et = threadIdx.x - 64;
            for (int i = et; i < 2 * kGnMaxImg * 32 * 2; i += 256) {
                const float v = gn_acc[i];
                const int slot = i / (kGnMaxImg * 64), rem = i - slot * (kGnMaxImg * 64), im = rem >> 6, k = rem & 63;
                if (v != 0.f && g.gn_stats[slot]) atomicAdd(g.gn_stats[slot] + (size_t)im * 64 + k, v);
            }
        }
    }

    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();      // PAIR: neither CTA may leave while its partner still uses its smem / TMEM / barriers
    if (warp == 2) {
        tcgen05_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// finishes a split-K product: workspace fp32 [M,N] -> epilogue -> fp16 out
__global__ void k_splitk_epilogue(const float* __restrict__ ws, GemmArgs g) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.M * g.N;
    if (i >= total) return;
    const long long m = i / g.N; const int n = (int)(i - m * g.N);
    const int x = (int)(m % g.W); const int y = (int)((m / g.W) % g.H); const int img = (int)(m / ((long long)g.W * g.H));
    float v = ws[i];
    if (g.bias) v += g.bias[n];
    if (g.temb) v += __half2float(g.temb[(long long)img * g.temb_ld + n]);
    if (g.residual) v += __half2float(g.residual[(long long)img * g.r_simg + (long long)y * g.r_sy + (long long)x * g.r_sx + n]);
    v = act_apply(v, g.act);
    g.out[(long long)img * g.o_simg + (long long)y * g.o_sy + (long long)x * g.o_sx + n] = __float2half_rn(v);
}

// ------------------------------------------------------------------ host side: plans
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

struct GemmPlan {
    CUtensorMap map_a, map_b;
    GemmArgs args;
    int block_n;
    int pair;        // 1: 2-CTA clusters (cta_group::2), grid is even
    int grid;
};

std::mutex g_plan_mu;
std::vector<GemmPlan*> g_plans;

template <int BN, bool PAIR, bool STATS>
int launch_gemm_v(const GemmPlan& p, cudaStream_t st) {
    using L = SmemLayout<BN, PAIR>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        SDF_CHECK_CUDA(cudaFuncSetAttribute(k_gemm<BN, PAIR, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        attr_set[dev] = true;
    }
    // split-K plans follow their workspace memset: programmatic launch needs a kernel as stream predecessor
    const bool pdl = sdf_pdl_enabled() && p.args.splitk == 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.grid); cfg.blockDim = dim3(kNumThreads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (PAIR) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        na++;
    }
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    SDF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_gemm<BN, PAIR, STATS>, p.map_a, p.map_b, p.args));
    return SDF_OK;
}

// the statistics-carrying epilogue is a separate instantiation: plans without a GroupNorm consumer keep the leaner one
template <int BN, bool PAIR>
int launch_gemm(const GemmPlan& p, cudaStream_t st) {
    if (p.args.gn_stats[0] || p.args.gn_stats[1]) return launch_gemm_v<BN, PAIR, true>(p, st);
    return launch_gemm_v<BN, PAIR, false>(p, st);
}

}  // namespace

// Creates a reusable plan for
//   out[img, y, x, n] = act(alpha * sum_{tap, c} a[img, y + dy(tap), x + dx(tap), c] * wt[(img, y,) n, tap*Cin + c]
//                           + bias[n] + temb[img, n] + residual[img, y, x, n])
// All strides are in ELEMENTS (fp16) and must be multiples of 8 (16-byte TMA alignment).
//   a        : element (img, y, x, c) at a[img*a_simg + y*a_sy + x*a_sx + c]; channels >= a_c_valid read as zero
//   wt       : row n, K index k at wt[(img*w_simg + y*w_sy) + n*w_ld + k]; k >= w_k_valid and rows >= n_rows_w read as zero.
//              w_sy / w_simg = 0 -> one weight matrix shared by all tiles (convolution / linear);
//              non-zero -> batched product (attention: y = head, img = batch); then tiles never span y / img.
//   Cin      : K elements iterated per tap, multiple of 64 (>= the real extent).  taps = 1, or 9 (3x3, stride 1, zero pad 1,
//              tap = ky*3 + kx, packed weights [n][tap][Cin]).
//   linear   : H = 1, Nimg = 1, W = rows.
//   bias fp32 [N], temb fp16 [Nimg, temb_ld], residual fp16 (strides r_*) optional (NULL).
//   act      : 0 none, 1 SiLU, 2 GELU(erf), 3 GEGLU (weight rows interleaved in 32-row chunks [16 value | 16 gate]; output N/2 wide).  splitk > 1 needs workspace fp32 [Nimg*H*W, N].  block_n in {64, 128, 160}.
//   cta_pair : 1 -> 2-CTA clusters (tcgen05 cta_group::2, M = 256 per MMA, each CTA stages half of the weight tile);
//              block_n in {128, 160, 256}; not for batched products.
// Returns a handle >= 0 or a negative error code.
SDF_API int sdf_gemm_plan_create_strided(const void* a, long long a_sx, long long a_sy, long long a_simg, int a_c_valid,
                                 const void* wt, long long w_ld, long long w_sy, long long w_simg, int w_k_valid, int n_rows_w,
                                 int Nimg, int H, int W, int Cin, int taps, int N,
                                 void* out, long long o_sx, long long o_sy, long long o_simg,
                                 const float* bias, const void* temb, int temb_ld,
                                 const void* residual, long long r_sx, long long r_sy, long long r_simg,
                                 int act, float alpha, int splitk, float* workspace, int block_n, int cta_pair, int stride, int pad_lo) {
    SDF_CHECK_ARG(stride == 1 || (stride == 2 && taps == 9), "gemm_plan: stride must be 1, or 2 for a 3x3 convolution");
    SDF_CHECK_ARG(pad_lo == 0 || pad_lo == 1, "gemm_plan: pad_lo must be 0 or 1");
    SDF_CHECK_ARG(a && wt && out, "gemm_plan: null pointer");
    SDF_CHECK_ARG(Cin > 0 && Cin % kBlockK == 0, "gemm_plan: Cin must be a positive multiple of 64");
    SDF_CHECK_ARG(taps == 1 || taps == 9, "gemm_plan: taps must be 1 or 9");
    SDF_CHECK_ARG(a_c_valid > 0 && a_c_valid <= Cin && w_k_valid > 0 && w_k_valid <= taps * Cin, "gemm_plan: bad valid extents");
    SDF_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)wt & 15) == 0 && ((uintptr_t)out & 15) == 0, "gemm_plan: pointers must be 16-byte aligned");
    SDF_CHECK_ARG(a_sx % 8 == 0 && a_sy % 8 == 0 && a_simg % 8 == 0 && w_ld % 8 == 0 && w_sy % 8 == 0 && w_simg % 8 == 0,
                  "gemm_plan: operand strides must be multiples of 8 elements");
    SDF_CHECK_ARG(o_sx % 8 == 0 && o_sy % 8 == 0 && o_simg % 8 == 0, "gemm_plan: output strides must be multiples of 8 elements");
    SDF_CHECK_ARG(!residual || (((uintptr_t)residual & 15) == 0 && r_sx % 8 == 0 && r_sy % 8 == 0 && r_simg % 8 == 0), "gemm_plan: residual alignment");
    SDF_CHECK_ARG(cta_pair ? (block_n == 128 || block_n == 160 || block_n == 256) : (block_n == 64 || block_n == 128 || block_n == 160),
                  "gemm_plan: block_n must be 64, 128 or 160 (128, 160 or 256 with cta_pair)");
    SDF_CHECK_ARG(!cta_pair || (w_sy == 0 && w_simg == 0), "gemm_plan: cta_pair needs one weight matrix shared by all tiles (not a batched product)");
    SDF_CHECK_ARG(N > 0 && n_rows_w > 0 && Nimg > 0 && H > 0 && W > 0, "gemm_plan: bad sizes");
    SDF_CHECK_ARG(splitk >= 1 && (splitk == 1 || workspace), "gemm_plan: split-K needs a workspace");
    SDF_CHECK_ARG(act != kActGeglu || (N % 32 == 0 && splitk == 1 && !residual && !temb), "gemm_plan: GEGLU epilogue needs N %% 32 == 0, no split-K, no residual / embedding");
    PFN_encodeTiled enc = get_encode();
    if (!enc) { sdf_set_error("gemm_plan: cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return SDF_ERR_CUDA; }

    GemmPlan* p = new GemmPlan();
    GemmArgs& g = p->args;
    g.M = Nimg * H * W; g.N = N; g.Cin = Cin; g.taps = taps; g.H = H; g.W = W; g.Nimg = Nimg;
    const bool batched = (w_sy != 0) || (w_simg != 0);
    // tile rectangle of up to 128 rows: as wide as possible, then (unless batched) rows, then images
    g.tw = W >= kBlockM ? kBlockM : W;
    g.th = 1; g.tn = 1;
    if (!batched && g.tw < kBlockM) {
        g.th = std::max(1, std::min(H, kBlockM / g.tw));
        g.tn = std::max(1, std::min(Nimg, kBlockM / (g.tw * g.th)));
    }
    g.pad = taps == 9 ? 1 : 0;
    g.stride = stride; g.pad_lo = pad_lo;
    {   // no split may own an empty K range: shrink splitk to ceil(kb / ceil(kb / splitk))
        const int kb_total = taps * (Cin / kBlockK);
        if (splitk > kb_total) splitk = kb_total;
        const int per = (kb_total + splitk - 1) / splitk;
        splitk = (kb_total + per - 1) / per;
    }
    g.splitk = splitk; g.w_by = w_sy != 0; g.w_bimg = w_simg != 0;
    g.bias = bias; g.temb = (const __half*)temb; g.temb_ld = temb_ld;
    g.residual = (const __half*)residual; g.r_sx = r_sx; g.r_sy = r_sy; g.r_simg = r_simg;
    g.out = (__half*)out; g.o_sx = o_sx; g.o_sy = o_sy; g.o_simg = o_simg; g.workspace = workspace;
    g.act = act; g.alpha = alpha;
    g.gn_stats[0] = g.gn_stats[1] = nullptr; g.gn_cpg[0] = g.gn_cpg[1] = 1; g.gn_coff[0] = g.gn_coff[1] = 0;
    p->block_n = block_n;
    p->pair = cta_pair ? 1 : 0;

    auto stride_or = [](long long s, long long fallback) { return (cuuint64_t)((s != 0 ? s : fallback) * 2); };
    {   // A: 4-D map (c, x, y, img), 128B swizzle, zero OOB fill
        // strided convolution: the map covers the INPUT (stride x the output extent); a box that traverses tw * stride elements with
        // element stride `stride` lands tw pixels — the same dense [tn][th][tw][64] smem tile as the unit-stride case
        cuuint64_t dims[4] = {(cuuint64_t)a_c_valid, (cuuint64_t)W * stride, (cuuint64_t)H * stride, (cuuint64_t)Nimg};
        const long long fx = a_sx, fy = a_sy ? a_sy : a_sx * W * stride, fi = a_simg ? a_simg : fy * H * stride;
        cuuint64_t strides[3] = {(cuuint64_t)fx * 2, (cuuint64_t)fy * 2, (cuuint64_t)fi * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(g.tw * stride), (cuuint32_t)(g.th * stride), (cuuint32_t)g.tn};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = enc(&p->map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete p; sdf_set_error("gemm_plan: cuTensorMapEncodeTiled(A) failed (%d)", (int)r); return SDF_ERR_CUDA; }
    }
    {   // B: 4-D map (k, n, y, img); the last two dims have extent 1 unless the product is batched
        cuuint64_t dims[4] = {(cuuint64_t)w_k_valid, (cuuint64_t)n_rows_w, (cuuint64_t)(g.w_by ? H : 1), (cuuint64_t)(g.w_bimg ? Nimg : 1)};
        const long long rows_bytes_el = w_ld * (long long)n_rows_w;
        cuuint64_t strides[3] = {(cuuint64_t)w_ld * 2, stride_or(w_sy, rows_bytes_el), stride_or(w_simg, rows_bytes_el * (g.w_by ? H : 1))};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(cta_pair ? block_n / 2 : block_n), 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p->map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(wt), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete p; sdf_set_error("gemm_plan: cuTensorMapEncodeTiled(B) failed (%d)", (int)r); return SDF_ERR_CUDA; }
    }
    const int tiles_x = (W + g.tw - 1) / g.tw, tiles_y = (H + g.th - 1) / g.th, tiles_i = (Nimg + g.tn - 1) / g.tn;
    const long long m_tiles = (long long)tiles_x * tiles_y * tiles_i;
    const long long work = (cta_pair ? (m_tiles + 1) / 2 : m_tiles) * ((N + block_n - 1) / block_n) * splitk;
    if (cta_pair) p->grid = 2 * (int)std::min<long long>(work, sdf_num_sms() / 2);
    else p->grid = work < sdf_num_sms() ? (int)work : sdf_num_sms();
    std::lock_guard<std::mutex> lk(g_plan_mu);
    g_plans.push_back(p);
    return (int)g_plans.size() - 1;
}

SDF_API int sdf_gemm_plan_create(const void* a, long long a_sx, long long a_sy, long long a_simg, int a_c_valid,
                                 const void* wt, long long w_ld, long long w_sy, long long w_simg, int w_k_valid, int n_rows_w,
                                 int Nimg, int H, int W, int Cin, int taps, int N,
                                 void* out, long long o_sx, long long o_sy, long long o_simg,
                                 const float* bias, const void* temb, int temb_ld,
                                 const void* residual, long long r_sx, long long r_sy, long long r_simg,
                                 int act, float alpha, int splitk, float* workspace, int block_n, int cta_pair) {
    return sdf_gemm_plan_create_strided(a, a_sx, a_sy, a_simg, a_c_valid, wt, w_ld, w_sy, w_simg, w_k_valid, n_rows_w, Nimg, H, W, Cin, taps, N,
                                        out, o_sx, o_sy, o_simg, bias, temb, temb_ld, residual, r_sx, r_sy, r_simg, act, alpha, splitk, workspace,
                                        block_n, cta_pair, 1, 1);
}

// Ask a plan's epilogue to accumulate the GroupNorm statistics of a consumer of its output: stats fp32 [Nimg, 32, 2] (sum, sum of squares;
// zeroed by the caller before the plan runs), `channels_per_group` of the consumer's GroupNorm(32, C), `channel_offset` = the consumer's
// channel that column 0 of this product lands on (non-zero when the output is one part of a concatenation).  Two consumers per plan.
// Requirements (else SDF_ERR_UNSUPPORTED and the caller keeps its separate statistics pass): no split-K, N % 32 == 0, no GEGLU, 16-byte
// aligned output rows, tiles whose 32-row quarters stay inside one image.
SDF_API int sdf_gemm_plan_set_gn_stats(int plan, int slot, float* stats, int channels_per_group, int channel_offset) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    SDF_CHECK_ARG(plan >= 0 && plan < (int)g_plans.size() && g_plans[plan], "gemm_plan_set_gn_stats: bad plan handle");
    SDF_CHECK_ARG(slot >= 0 && slot < 2 && stats && channels_per_group > 0 && channel_offset >= 0, "gemm_plan_set_gn_stats: bad arguments");
    GemmArgs& g = g_plans[plan]->args;
    const bool ok = g.splitk == 1 && (g.N % 32) == 0 && g.act != kActGeglu && ((g.tw * g.th) % 32) == 0 && (g.o_sx % 8) == 0 && g.Nimg <= kGnMaxImg &&
                    ((reinterpret_cast<uintptr_t>(g.out) & 15) == 0);
    if (!ok) { sdf_set_error("gemm_plan_set_gn_stats: plan shape cannot carry statistics (split-K / ragged N / GEGLU / tile geometry)"); return SDF_ERR_UNSUPPORTED; }
    g.gn_stats[slot] = stats; g.gn_cpg[slot] = channels_per_group; g.gn_coff[slot] = channel_offset;
    return SDF_OK;
}

SDF_API int sdf_gemm_run(int plan, void* stream) {
    GemmPlan* p;
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        SDF_CHECK_ARG(plan >= 0 && plan < (int)g_plans.size() && g_plans[plan], "gemm_run: bad plan handle");
        p = g_plans[plan];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const GemmArgs& g = p->args;
    if (g.splitk > 1) SDF_CHECK_CUDA(cudaMemsetAsync(g.workspace, 0, (size_t)g.M * g.N * sizeof(float), st));
    int rc;
    if (p->pair) {
        if (p->block_n == 128) rc = launch_gemm<128, true>(*p, st);
        else if (p->block_n == 160) rc = launch_gemm<160, true>(*p, st);
        else rc = launch_gemm<256, true>(*p, st);
    } else {
        if (p->block_n == 64) rc = launch_gemm<64, false>(*p, st);
        else if (p->block_n == 128) rc = launch_gemm<128, false>(*p, st);
        else rc = launch_gemm<160, false>(*p, st);
    }
    if (rc) return rc;
    SDF_CHECK_LAUNCH("gemm");
    if (g.splitk > 1) {
        const long long total = (long long)g.M * g.N;
        sdf_launch_pdl(k_splitk_epilogue, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, st, (const float*)g.workspace, g);
        SDF_CHECK_LAUNCH("gemm(split-K epilogue)");
    }
    return SDF_OK;
}

SDF_API int sdf_gemm_plan_destroy(int plan) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    SDF_CHECK_ARG(plan >= 0 && plan < (int)g_plans.size(), "gemm_plan_destroy: bad plan handle");
    delete g_plans[plan];
    g_plans[plan] = nullptr;
    return SDF_OK;
}
