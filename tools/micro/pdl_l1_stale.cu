// Which loads of a kernel launched with programmatic stream serialization can return STALE L1 lines?
//   iteration i:  writer (normal launch) fills A with i;  middle (normal launch, triggers early) fills B with i;
//   reader (PDL) loads A before griddepcontrol.wait (A's writer is complete: `middle` is fully ordered after it) and
//   B after the wait, each through ld.global.nc (__ldg), plain ld.global and ld.global.cg.
// All reader CTAs read the same 4 KB every iteration, so every SM's L1 holds last iteration's lines.
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int N = 1024;
__global__ void writer(int* A, int v) { for (int j = threadIdx.x; j < N; j += blockDim.x) A[j] = v; }
__global__ void middle(int* B, int v, long long cycles) {
    asm volatile("griddepcontrol.launch_dependents;");
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    for (int j = threadIdx.x; j < N; j += blockDim.x) B[j] = v;
}
__device__ __forceinline__ int ld_plain(const int* p) { int v; asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__global__ void reader(const int* A, const int* B, int v, int* bad) {
    int pre_nc = 0, pre_pl = 0, pre_cg = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        pre_nc += __ldg(A + j) != v;
        pre_pl += ld_plain(A + j) != v;
        pre_cg += __ldcg(A + j) != v;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int post_nc = 0, post_pl = 0, post_cg = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        post_nc += __ldg(B + j) != v;
        post_pl += ld_plain(B + j) != v;
        post_cg += __ldcg(B + j) != v;
    }
    if (pre_nc) atomicAdd(bad + 0, pre_nc);
    if (pre_pl) atomicAdd(bad + 1, pre_pl);
    if (pre_cg) atomicAdd(bad + 2, pre_cg);
    if (post_nc) atomicAdd(bad + 3, post_nc);
    if (post_pl) atomicAdd(bad + 4, post_pl);
    if (post_cg) atomicAdd(bad + 5, post_cg);
}
static cudaError_t launch_reader(cudaStream_t st, const int* A, const int* B, int v, int* bad, int pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148 * 2); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, reader, A, B, v, bad);
}
int main() {
    int *A, *B, *bad;
    CK(cudaMalloc(&A, N * 4)); CK(cudaMalloc(&B, N * 4)); CK(cudaMalloc(&bad, 6 * 4));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int pdl = 0; pdl < 2; ++pdl)
        for (int slow = 0; slow < 2; ++slow) {
            CK(cudaMemset(bad, 0, 24));
            for (int i = 1; i <= 200; ++i) {
                writer<<<1, 256, 0, st>>>(A, i);
                middle<<<1, 256, 0, st>>>(B, i, slow ? 20000 : 0);
                CK(launch_reader(st, A, B, i, bad, pdl));
            }
            CK(cudaStreamSynchronize(st));
            int h[6];
            CK(cudaMemcpy(h, bad, 24, cudaMemcpyDeviceToHost));
            printf("pdl=%d slow_middle=%d  stale words: pre-wait nc %d plain %d cg %d | post-wait nc %d plain %d cg %d\n", pdl, slow,
                   h[0], h[1], h[2], h[3], h[4], h[5]);
        }
    return 0;
}
