// fp32 FFMA grouped GEMM (FHVAE_MODE_F32_SIMT): the exact-arithmetic reference kernel that every
// tensor-core kernel in this library is checked against on the GPU, and the fall-through for
// shapes the tcgen05 kernels do not cover (tiny heads).  64x64x16 tiles, 256 threads, 4x4 / thread.
#include "common.cuh"

namespace fhvae {

constexpr int TM = 64, TN = 64, TK = 16, PAD = 4;

struct GemmBatchParams {
    fhvae_gemm_problem p[FHVAE_GEMM_MAX_BATCH];
    int tile_start[FHVAE_GEMM_MAX_BATCH + 1];
    int tiles_n[FHVAE_GEMM_MAX_BATCH];
    int n;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const __grid_constant__ GemmBatchParams bp) {
    __shared__ __align__(16) float As[TK][TM + PAD];
    __shared__ __align__(16) float Bs[TK][TN + PAD];

    int pi = 0;
    while (pi + 1 < bp.n && (int)blockIdx.x >= bp.tile_start[pi + 1]) ++pi;
    const fhvae_gemm_problem& P = bp.p[pi];
    const int tile = blockIdx.x - bp.tile_start[pi];
    const int m0 = (tile / bp.tiles_n[pi]) * TM;
    const int n0 = (tile % bp.tiles_n[pi]) * TN;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int M = P.M, N = P.N, K = P.K;
    const bool a_kfast = (P.sa_k == 1);
    const bool b_nfast = (P.sb_n == 1);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            int m, k;
            if (a_kfast) { k = e & 15; m = e >> 4; } else { m = e & 63; k = e >> 6; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < K) ? __ldg(P.A + (int64_t)gm * P.sa_m + (int64_t)gk * P.sa_k) : 0.f;
            int n, kb;
            if (b_nfast) { n = e & 63; kb = e >> 6; } else { kb = e & 15; n = e >> 4; }
            const int gn = n0 + n, gkb = k0 + kb;
            Bs[kb][n] = (gn < N && gkb < K) ? __ldg(P.B + (int64_t)gkb * P.sb_k + (int64_t)gn * P.sb_n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float* c = P.C + (int64_t)gm * P.ldc + gn;
            float v = acc[i][j];
            if (P.bias) v += __ldg(P.bias + gn);
            if (P.beta != 0.f) v += P.beta * (*c);
            if (P.relu) v = fmaxf(v, 0.f);
            *c = v;
        }
    }
}

int gemm_batch_simt(const fhvae_gemm_problem* problems, int n, cudaStream_t st) {
    GemmBatchParams bp;
    memset(&bp, 0, sizeof(bp));
    int total = 0;
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const fhvae_gemm_problem& p = problems[i];
        if (p.M == 0 || p.N == 0) continue;
        bp.p[k] = p;
        bp.tile_start[k] = total;
        bp.tiles_n[k] = cdiv(p.N, TN);
        total += cdiv(p.M, TM) * bp.tiles_n[k];
        ++k;
    }
    bp.tile_start[k] = total;
    bp.n = k;
    if (total == 0) return 0;
    gemm_simt_kernel<<<total, 256, 0, st>>>(bp);
    FHVAE_LAUNCH_CHECK("gemm_simt");
    return 0;
}

}  // namespace fhvae
