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

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl may become resident while its predecessor
// on the same stream is still draining; it must execute pdl_wait() before it reads anything the predecessor wrote (or
// writes anything the predecessor reads).  What runs before pdl_wait() -- barrier / TMEM setup, loads of data that is
// older than the predecessor (weights, packed operand images) -- overlaps the predecessor's tail; the ~2 us of
// launch latency between dependent graph nodes disappears.  Without the launch attribute both instructions are no-ops.
// RULE for every kernel launched this way: nothing a predecessor produced may be read through ld.global.nc (__ldg, or
// plain loads through `const __restrict__` kernel parameters, which the compiler turns into ld.global.nc) -- such loads
// carry no ordering and were observed hoisted above griddepcontrol.wait (stale Q rows in the wavefront LSTM).  These
// kernels read with __ldcg / plain loads and take no const __restrict__ pointers to produced data.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
enum { PDL_WAVE = 1, PDL_HEADS = 2, PDL_GEMM = 4, PDL_ELBO = 8, PDL_MISC = 16 };
bool pdl_enabled(int family);   // env FHVAE_PDL = bit mask of kernel families (default: all)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int family, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled(family) ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace fhvae
