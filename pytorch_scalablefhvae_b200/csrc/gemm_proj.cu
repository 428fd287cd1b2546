// Projection GEMM  C[m,n] = sum_k A[m,k] W[n,k] (+ bias[n])  -- nn.Linear forward (simple_fhvae.py:130-134, 208-212) for the
// SHORT-K, LARGE-M products on the critical path of the LSTM model: the layer-0 input projections of x over all T*B
// rows (K = F = 80, N = 4H = 1024: a 21 MB fp32 output) and the decoder's Gaussian head (K = H = 256, N = 2F = 160).
//
// gemm_tc.cu serves them with fp32 operands converted on the way into shared memory (register-staged producer warps)
// and an epilogue whose threads each own one output ROW (32 rows x 16 B per warp store: half-used sectors); it took
// 22 us for the x projection, 7 % tensor pipe, far from the ~4 us its 21 MB of output cost at HBM speed.  Here:
//   * operands are bf16 hi/lo planes [plane][row][K] (K contiguous = "K-major").  A K-block of 32 elements x 128 rows is ONE
//     cp.async.bulk.tensor box of 64 B x 128 rows = 8 KB with the 64-byte swizzle: it lands in the canonical K-major
//     SWIZZLE_64B UMMA layout (8-row x 64-byte atoms, SBO = 512 B; the second K = 16 step of a block is +32 B on the
//     descriptor's start address): no registers, no conversion, no LSU on the way in.  (First version: no-swizzle layout
//     from boxes of 16 B x 128 rows -- 40 boxes of 128 sixteen-byte rows per tile kept the TMA unit busy for longer than
//     the whole gemm_tc kernel took: 26 vs 17 us.)
//   * CTA tile 128 x 128, BK = 32 (one box per plane and operand: 32 KB per stage), up to 3 stages in flight -> 96 KB,
//     two CTAs per SM; warp 0 = TMA producer, warp 1 = tcgen05 issuer (lo*hi + hi*lo + hi*hi in the bf16x3 mode),
//     warps 2-5 = epilogue;
//   * epilogue: tcgen05.ld (thread = accumulator row) -> padded shared-memory tile -> each warp writes whole 128-byte
//     row segments (bias added on the way): every store instruction is one full line.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace fhvae {

using namespace tc;

constexpr int PJ_BM = 128, PJ_BN = 128, PJ_BK = 32, PJ_NS = 3;
constexpr int PJ_PLANE = 128 * PJ_BK * 2;                     // one TMA box: 32 elements (64 B) x 128 rows = 8 KB: one operand,
                                                              // one bf16 part, one stage
constexpr int PJ_STAGE = 4 * PJ_PLANE;                        // [A_hi][A_lo][B_hi][B_lo] = 32 KB
constexpr int PJ_STG_LD = 33;                                 // epilogue staging: 128 rows x 32 columns, padded
constexpr int PJ_SMEM = PJ_NS * PJ_STAGE + 1024 + 256;        // + alignment slack + barriers   (staging reuses stage 0)
constexpr int PJ_THREADS = 192;
constexpr int PJ_MAX_BATCH = 8;

struct PjProblem {
    CUtensorMap ta, tb;
    float* C;
    const float* bias;
    long long ldc;
    int M, N, K;
    int tile_start, tiles_n;
    int pad[2];
};
struct PjBatch {
    PjProblem p[PJ_MAX_BATCH];
    int n, passes;
};

__device__ __forceinline__ void pj_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pj_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint64_t pj_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // LBO: not used by swizzled K-major layouts
    d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;   // SBO: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
    return d;
}

__global__ void __launch_bounds__(PJ_THREADS, 2) proj_tma_kernel(const __grid_constant__ PjBatch pb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + PJ_NS * PJ_STAGE);
    uint64_t* empty = full + PJ_NS;
    uint64_t* accd = empty + PJ_NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accd + 1);

    int pi = 0;
    while (pi + 1 < pb.n && (int)blockIdx.x >= pb.p[pi + 1].tile_start) ++pi;
    const PjProblem& P = pb.p[pi];
    const int t = blockIdx.x - P.tile_start;
    const int m0 = (t / P.tiles_n) * PJ_BM, n0 = (t % P.tiles_n) * PJ_BN;
    const int nkb = (P.K + PJ_BK - 1) / PJ_BK;
    const int nrem = min(PJ_BN, P.N - n0);
    const int nt = (nrem + 15) & ~15;                        // MMA N (multiple of 16)
    const int planes = pb.passes == 3 ? 2 : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 1) tmem_alloc<PJ_BN>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < PJ_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accd, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % PJ_NS;
                if (it >= PJ_NS) mbar_wait(&empty[s], ((it / PJ_NS) - 1) & 1);
                const int k0 = it * PJ_BK;
                pj_expect_tx(&full[s], (uint32_t)(planes * 2 * PJ_PLANE));    // (a box past K is zero-filled, still full size)
                const uint32_t st = smem_u32(smem + s * PJ_STAGE);
                for (int pl = 0; pl < planes; ++pl) {
                    pj_tma_load_3d(st + pl * PJ_PLANE, &P.ta, k0, m0, pl, &full[s]);
                    pj_tma_load_3d(st + (2 + pl) * PJ_PLANE, &P.tb, k0, n0, pl, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(PJ_BM, nt);                // A and B K-major
            for (int it = 0; it < nkb; ++it) {
                const int s = it % PJ_NS;
                mbar_wait(&full[s], (it / PJ_NS) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * PJ_STAGE);
                const int nk16 = (min(PJ_BK, P.K - it * PJ_BK) + 15) >> 4;    // K = 16 steps of this block (K % 16 == 0)
                for (int j = 0; j < nk16; ++j) {
                    const uint32_t ko = (uint32_t)j * 32;                     // 16 bf16 further along the 64-byte swizzled rows
                    const uint64_t dah = pj_desc_sw64(st + ko), dal = pj_desc_sw64(st + PJ_PLANE + ko);
                    const uint64_t dbh = pj_desc_sw64(st + 2 * PJ_PLANE + ko), dbl = pj_desc_sw64(st + 3 * PJ_PLANE + ko);
                    const uint32_t acc0 = (it > 0 || j > 0) ? 1u : 0u;
                    if (planes == 2) {
                        umma_bf16(tmem_d, dal, dbh, idesc, acc0);
                        umma_bf16(tmem_d, dah, dbl, idesc, 1u);
                        umma_bf16(tmem_d, dah, dbh, idesc, 1u);
                    } else {
                        umma_bf16(tmem_d, dah, dbh, idesc, acc0);
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accd);
        }
    } else {
        // ================= epilogue =================
        mbar_wait(accd, 0);
        tc_fence_after();
        const int q = warp & 3;                              // TMEM lane quarter this warp may read = its 32 rows
        float* stg = reinterpret_cast<float*>(smem) + q * 32 * PJ_STG_LD;   // stage 0 is free: every MMA has completed
        for (int c0 = 0; c0 < nt; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * PJ_STG_LD + j] = v[j];
            __syncwarp();
            const int n = n0 + c0 + lane;
            const float bv = (P.bias && c0 + lane < nrem) ? __ldg(P.bias + n) : 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int m = m0 + q * 32 + r;
                if (m < P.M && c0 + lane < nrem) P.C[(long long)m * P.ldc + n] = stg[r * PJ_STG_LD + lane] + bv;
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<PJ_BN>(tmem_d);
}

typedef CUresult (*PjEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PjEncodeTiledFn pj_get_encode() {
    static PjEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PjEncodeTiledFn>(p);
    }
    return fn;
}
// planes tensor [2][rows][K] of bf16 (row stride ld, plane stride ps, in elements) -> 3-D map, box 32 x 128 x 1, 64B swizzle
static int pj_make_map(CUtensorMap* map, const void* base, int K, int rows, long long ld, long long ps) {
    PjEncodeTiledFn enc = pj_get_encode();
    if (!enc) { set_error("proj_planes: cuTensorMapEncodeTiled is not available from the driver"); return FHVAE_ENOSUP; }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ps * 2};
    cuuint32_t box[3] = {(cuuint32_t)PJ_BK, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("proj_planes: cuTensorMapEncodeTiled failed (%d)", (int)r); return FHVAE_EINVAL; }
    return 0;
}

}  // namespace fhvae

using namespace fhvae;

extern "C" int fhvae_proj_planes_batch(const fhvae_proj_problem* problems, int n, int mode, void* stream) {
    FHVAE_CHECK_ARG(problems && n > 0 && n <= FHVAE_PROJ_MAX_BATCH, "proj_planes: need 1..%d problems", FHVAE_PROJ_MAX_BATCH);
    FHVAE_CHECK_ARG(mode == FHVAE_MODE_BF16X3 || mode == FHVAE_MODE_BF16, "proj_planes: mode must be BF16X3 or BF16");
    static_assert(FHVAE_PROJ_MAX_BATCH == PJ_MAX_BATCH, "batch size");
    static int attr_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_dev[dev]) {
        cudaError_t e = cudaFuncSetAttribute(proj_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PJ_SMEM);
        if (e != cudaSuccess) { set_error("proj_planes: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_dev[dev] = 1;
    }
    alignas(64) PjBatch pb;
    memset(&pb, 0, sizeof(pb));
    pb.passes = mode == FHVAE_MODE_BF16X3 ? 3 : 1;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        const fhvae_proj_problem& p = problems[i];
        FHVAE_CHECK_ARG(p.A && p.W && p.C && p.M > 0 && p.N > 0 && p.K > 0, "proj_planes: problem %d: bad pointer or size", i);
        // K % 16 == 0: a K = 16 MMA step never reads a chunk that was not loaded
        FHVAE_CHECK_ARG(p.K % 16 == 0 && p.lda % 8 == 0 && p.ldw % 8 == 0 && p.a_plane_stride % 8 == 0 && p.w_plane_stride % 8 == 0 &&
                        ((uintptr_t)p.A & 15) == 0 && ((uintptr_t)p.W & 15) == 0,
                        "proj_planes: problem %d: K %% 16, 16-byte aligned planes and strides multiple of 8 elements required", i);
        PjProblem& q = pb.p[pb.n];
        int r = pj_make_map(&q.ta, p.A, p.K, p.M, p.lda, p.a_plane_stride);
        if (r) return r;
        r = pj_make_map(&q.tb, p.W, p.K, p.N, p.ldw, p.w_plane_stride);
        if (r) return r;
        q.C = p.C; q.bias = p.bias; q.ldc = p.ldc; q.M = p.M; q.N = p.N; q.K = p.K;
        q.tiles_n = cdiv(p.N, PJ_BN);
        q.tile_start = total;
        total += cdiv(p.M, PJ_BM) * q.tiles_n;
        ++pb.n;
    }
    proj_tma_kernel<<<total, PJ_THREADS, PJ_SMEM, as_stream(stream)>>>(pb);
    FHVAE_LAUNCH_CHECK("proj_planes");
    return 0;
}
