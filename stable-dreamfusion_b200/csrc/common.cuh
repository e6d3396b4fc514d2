// common.cuh — shared helpers for the sm_100a kernels behind the C-ABI (include/sdf_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define SDF_API extern "C" __attribute__((visibility("default")))

// error codes returned by every sdf_* entry point
enum { SDF_OK = 0, SDF_ERR_ARG = -1, SDF_ERR_CUDA = -2, SDF_ERR_UNSUPPORTED = -3 };

void sdf_set_error(const char* fmt, ...);

#define SDF_CHECK_ARG(cond, ...)                                   \
    do { if (!(cond)) { sdf_set_error(__VA_ARGS__); return SDF_ERR_ARG; } } while (0)

#define SDF_CHECK_LAUNCH(name)                                                      \
    do { cudaError_t e_ = cudaPeekAtLastError();                                    \
         if (e_ != cudaSuccess) { sdf_set_error("%s: %s", name, cudaGetErrorString(e_)); \
                                  (void)cudaGetLastError(); return SDF_ERR_CUDA; } } while (0)

#define SDF_CHECK_CUDA(expr)                                                        \
    do { cudaError_t e_ = (expr);                                                   \
         if (e_ != cudaSuccess) { sdf_set_error("%s: %s", #expr, cudaGetErrorString(e_)); \
                                  return SDF_ERR_CUDA; } } while (0)

static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static inline uint64_t cdiv64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// SM count of the current device (148 on B200), queried once: grids are sized in multiples of it
int sdf_num_sms();

// ---- programmatic dependent launch (PDL).  The SD launch lists are ~950 small dependent kernels per step; with PDL the next
// kernel's CTAs are scheduled as soon as the previous grid's CTAs have all STARTED (every kernel triggers at its first
// instruction), run their prologue (barrier init, TMEM allocation, tensor-map prefetch, index math) and then block in
// griddepcontrol.wait until the previous grid has completed and flushed — launch latency and prologues leave the critical path.
// Contract for a PDL-aware kernel: pdl_prologue() (or trigger + wait) before the first global-memory access.
// SDF_PDL=0 in the environment launches everything with full stream serialisation (the wait is then a no-op).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }
bool sdf_pdl_enabled();

// Launch `kernel` so that it may overlap the tail of its stream predecessor (which must be a kernel: callers use a plain
// <<<>>> launch right after a memset).
template <typename... KArgs, typename... Args>
inline cudaError_t sdf_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = sdf_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
