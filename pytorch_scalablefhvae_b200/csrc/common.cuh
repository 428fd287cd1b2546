// Shared helpers for libfhvae_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/fhvae_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfhvae_b200 is written for sm_100a only"
#endif

namespace fhvae {

void set_error(const char* fmt, ...);
void count_launches(int n);   // process-wide kernel-launch counter (bench.py gpu_launches)
bool deterministic_mode();    // fhvae_set_deterministic: split-K launches keep a fixed summation order

#define FHVAE_CHECK_ARG(cond, ...)                     \
    do {                                               \
        if (!(cond)) {                                 \
            ::fhvae::set_error(__VA_ARGS__);           \
            return FHVAE_EINVAL;                       \
        }                                              \
    } while (0)

#define FHVAE_CHECK_SUP(cond, ...)                     \
    do {                                               \
        if (!(cond)) {                                 \
            ::fhvae::set_error(__VA_ARGS__);           \
            return FHVAE_ENOSUP;                       \
        }                                              \
    } while (0)

// after a kernel launch: report launch-configuration errors without synchronising
#define FHVAE_LAUNCH_CHECK(name)                                                        \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        ::fhvae::count_launches(1);                                                     \
        if (e__ != cudaSuccess) {                                                       \
            ::fhvae::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__;                                                            \
        }                                                                               \
    } while (0)

constexpr float kLog2Pi = 1.8378770664093453f;
constexpr float kPz2Logvar = -1.3862943611198906f;   // log(0.5^2), simple_fhvae.py:88
constexpr float kInvS2 = 4.0f;                        // 1 / exp(kPz2Logvar)
constexpr int kNumSM = 148;

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
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU-based (ex2.approx + rcp.approx) gate nonlinearities for the persistent recurrence: absolute
// error ~2e-7 (a few ulp of the [0,1] / [-1,1] range), far inside the 1e-4 parity budget.
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace fhvae
