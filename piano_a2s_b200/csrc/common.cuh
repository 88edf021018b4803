// Shared helpers for the piano-a2s B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PA2S_API extern "C" __attribute__((visibility("default")))

// Every kernel launch in this library goes through PA2S_LAUNCH_CHECK so that `pa2s_launch_count()`
// reports how many of OUR kernels ran (bench.py's `gpu_launches`).
extern unsigned long long g_pa2s_launches;
#define PA2S_COUNT_LAUNCH() (++g_pa2s_launches)
#define PA2S_CHECK_LAST()                                   \
    do {                                                    \
        PA2S_COUNT_LAUNCH();                                \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)
#define PA2S_TRY(x)                                         \
    do {                                                    \
        cudaError_t e__ = (x);                              \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

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
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
