// Does griddepcontrol.wait in a kernel launched with programmatic stream serialization still order the kernel after the
// preceding kernel of its stream when a cudaStreamWaitEvent (cross-stream join) sits between the two launches?
// Cases: eager / captured, with and without the intervening join.  Prints 1 per case when the ordering held.
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void slow_writer(int* flag, long long cycles) {
    asm volatile("griddepcontrol.launch_dependents;");
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 0 && blockIdx.x == 0) *flag = 1;
}
__global__ void side_kernel(int* x) { if (threadIdx.x == 0) *x = 7; }
__global__ void reader(const int* flag, int* seen) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) *seen = *(volatile const int*)flag;
}
static cudaError_t launch_reader(cudaStream_t st, const int* flag, int* seen, int pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, reader, flag, seen);
}
int main() {
    int *flag, *seen, *x;
    CK(cudaMalloc(&flag, 4)); CK(cudaMalloc(&seen, 4)); CK(cudaMalloc(&x, 4));
    cudaStream_t main_s, side_s;
    CK(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&side_s, cudaStreamNonBlocking));
    cudaEvent_t fork_e, join_e;
    CK(cudaEventCreateWithFlags(&fork_e, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&join_e, cudaEventDisableTiming));
    for (int captured = 0; captured < 2; ++captured)
        for (int join = 0; join < 3; ++join)
            for (int pdl = 0; pdl < 2; ++pdl) {
                int ok = 0;
                for (int rep = 0; rep < 20; ++rep) {
                    CK(cudaMemset(flag, 0, 4)); CK(cudaMemset(seen, 0xff, 4));
                    CK(cudaDeviceSynchronize());
                    cudaGraph_t g; cudaGraphExec_t ge;
                    if (captured) CK(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal));
                    if (join) {                                  // side work forked BEFORE the slow kernel: done long before the join
                        CK(cudaEventRecord(fork_e, main_s));
                        CK(cudaStreamWaitEvent(side_s, fork_e, 0));
                        side_kernel<<<1, 32, 0, side_s>>>(x);
                        CK(cudaEventRecord(join_e, side_s));
                    }
                    slow_writer<<<4, 32, 0, main_s>>>(flag, 200000);
                    if (join == 2) {                              // a second fork between the two kernels (event RECORD on main)
                        CK(cudaEventRecord(fork_e, main_s));
                    }
                    if (join) CK(cudaStreamWaitEvent(main_s, join_e, 0));
                    CK(launch_reader(main_s, flag, seen, pdl));
                    if (captured) {
                        CK(cudaStreamEndCapture(main_s, &g));
                        CK(cudaGraphInstantiate(&ge, g, 0));
                        CK(cudaGraphLaunch(ge, main_s));
                    }
                    CK(cudaStreamSynchronize(main_s)); CK(cudaDeviceSynchronize());
                    int h = -1;
                    CK(cudaMemcpy(&h, seen, 4, cudaMemcpyDeviceToHost));
                    ok += (h == 1);
                    if (captured) { cudaGraphExecDestroy(ge); cudaGraphDestroy(g); }
                }
                printf("captured=%d join=%d pdl=%d : ordered %d/20\n", captured, join, pdl, ok);
            }
    return 0;
}
