// Weight-gradient GEMM  C[m,n] = sum_k A[k,m] B[k,n]  (K = T*B rows, both operands row-contiguous) on tcgen05,
// fed by TMA from pre-split bf16 planes.
//
// The fp32-parity mode computes x*y as hi*hi + hi*lo + lo*hi with x = hi + lo (bf16 each).  gemm_tc.cu does that
// split on the fly (global fp32 -> registers -> cvt -> st.shared): for the long-K weight gradients that producer
// is bound by the LSU / L1 wavefront rate at ~13-27 B/clk/SM, 4-5x under what the tensor pipe needs.  Here the
// operands already sit in HBM as two bf16 planes [plane][k][m] (written by fhvae_split_planes_batch, or directly by
// the kernel that produces the tensor), so a K-block is a handful of cp.async.bulk.tensor boxes that land in
// shared memory already in the UMMA layout -- no registers, no conversions, no LSU:
//   * operand tile = "MN-major, 128-byte swizzle" canonical layout: a TMA box of 64 elements (128 B) x BK rows
//     is exactly one swizzle-atom column (8-row groups 1024 B apart => SBO = 1024; the next 64 elements of M/N
//     are the next box => LBO = box size).  Instruction descriptor: a_major = b_major = MN.
//   * CTA tile 128 x 256 (N of a recurrent weight gradient = H = 256: the 4H-wide dgates operand is read once),
//     accumulator = 256 TMEM columns, BK = 32, 4-stage mbarrier ring (48 KB per stage), 1 CTA / SM.
//   * warp 0 = TMA producer, warp 1 = tcgen05 issuer (3 MMAs 128x256x16 per k16 step: lo*hi, hi*lo, hi*hi),
//     warps 2-5 = epilogue (tcgen05.ld -> red.global.add.v4.f32 into the pre-zeroed C; split-K partials).
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace fhvae {

using namespace tc;

constexpr int WG_BM = 128, WG_BN = 256, WG_BK = 32, WG_NS = 4;
constexpr int WG_BOX = 64 * WG_BK * 2;                       // one TMA box: 64 elements x 32 rows of bf16 = 4 KB
constexpr int WG_A_PLANE = (WG_BM / 64) * WG_BOX;            // 8 KB
constexpr int WG_B_PLANE = (WG_BN / 64) * WG_BOX;            // 16 KB
constexpr int WG_STAGE = 2 * (WG_A_PLANE + WG_B_PLANE);      // [A_hi][A_lo][B_hi][B_lo] = 48 KB
constexpr int WG_SMEM = WG_NS * WG_STAGE + 1024 + 256;       // + alignment slack + barriers
constexpr int WG_THREADS = 192;
constexpr uint32_t WG_LBO = WG_BOX, WG_SBO = 1024;

struct WgProblem {
    CUtensorMap ta, tb;
    float* C;
    long long ldc;
    int M, N, K;
    int c_vec;
    int tile_start, tiles_m, tiles_n, ksplit, kb_per_split;
    int store;                  // ksplit == 1: the epilogue stores (no pre-zero pass, no atomics)
    int pad[2];
};
struct WgBatch {
    WgProblem p[FHVAE_WGRAD_MAX_BATCH];
    int n, passes;
    int ns;                     // stages of the ring actually used: 4 (1 CTA/SM) for long K, 2 (two CTAs/SM share the
                                // SM: one tile's epilogue overlaps the other's loads) when every CTA has <= 4 K-blocks
};

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((WG_LBO >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((WG_SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(WG_THREADS, NS == 2 ? 2 : 1) wgrad_tma_kernel(const __grid_constant__ WgBatch wb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * WG_STAGE);
    uint64_t* empty = full + WG_NS;
    uint64_t* accd = empty + WG_NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accd + 1);

    int pi = 0;
    while (pi + 1 < wb.n && (int)blockIdx.x >= wb.p[pi + 1].tile_start) ++pi;
    const WgProblem& P = wb.p[pi];
    int t = blockIdx.x - P.tile_start;
    const int tiles_mn = P.tiles_m * P.tiles_n;
    const int split = t / tiles_mn;
    t -= split * tiles_mn;
    const int m0 = (t / P.tiles_n) * WG_BM, n0 = (t % P.tiles_n) * WG_BN;
    const int nkb = (P.K + WG_BK - 1) / WG_BK;
    const int kb0 = split * P.kb_per_split;
    const int nit = min(nkb, kb0 + P.kb_per_split) - kb0;
    const int mrem = min(WG_BM, P.M - m0), nrem = min(WG_BN, P.N - n0);
    const int nbox_a = (mrem + 63) >> 6, nbox_b = (nrem + 63) >> 6;
    const int nt = (nrem + 15) & ~15;                        // MMA N (multiple of 16)
    const int planes = wb.passes == 3 ? 2 : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 1) tmem_alloc<WG_BN>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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
            const uint32_t bytes = (uint32_t)(planes * (nbox_a + nbox_b) * WG_BOX);
            for (int it = 0; it < nit; ++it) {
                const int s = it % NS;
                if (it >= NS) mbar_wait(&empty[s], ((it / NS) - 1) & 1);
                mbar_expect_tx(&full[s], bytes);
                const uint32_t st = smem_u32(smem + s * WG_STAGE);
                const int k = (kb0 + it) * WG_BK;
                for (int pl = 0; pl < planes; ++pl) {
                    for (int j = 0; j < nbox_a; ++j)
                        tma_load_3d(st + pl * WG_A_PLANE + j * WG_BOX, &P.ta, m0 + 64 * j, k, pl, &full[s]);
                    for (int j = 0; j < nbox_b; ++j)
                        tma_load_3d(st + 2 * WG_A_PLANE + pl * WG_B_PLANE + j * WG_BOX, &P.tb, n0 + 64 * j, k, pl, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(WG_BM, nt) | (1u << 15) | (1u << 16);     // A and B MN-major
            for (int it = 0; it < nit; ++it) {
                const int s = it % NS;
                mbar_wait(&full[s], (it / NS) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * WG_STAGE);
#pragma unroll
                for (int j = 0; j < WG_BK / 16; ++j) {
                    const uint32_t ko = (uint32_t)j * 2 * WG_SBO;             // 16 rows = two 8-row swizzle groups
                    const uint64_t dah = make_desc_sw128(st + ko), dal = make_desc_sw128(st + WG_A_PLANE + ko);
                    const uint64_t dbh = make_desc_sw128(st + 2 * WG_A_PLANE + ko);
                    const uint64_t dbl = make_desc_sw128(st + 2 * WG_A_PLANE + WG_B_PLANE + ko);
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
        // ================= epilogue: TMEM -> registers -> global (store, or red.add for split-K partials) ==========
        // A thread owns one accumulator row and writes 16-byte pieces of it straight from registers, chunk by chunk
        // behind each tcgen05.ld.  (Staging the tile through shared memory for 512-byte coalesced rows was measured
        // SLOWER -- 27 vs 18.6 us on the 21 MB projection output: the extra pass serialises load, barrier and store.)
        mbar_wait(accd, 0);
        tc_fence_after();
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const int m = m0 + q * 32 + lane;
        float* crow = P.C + (long long)m * P.ldc + n0;
        for (int c0 = 0; c0 < nt; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (m < P.M) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = c0 + 4 * j;
                    if (P.c_vec && n + 4 <= nrem) {
                        if (P.store)
                            *reinterpret_cast<float4*>(crow + n) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        else
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + n), "f"(v[4 * j]),
                                         "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (n + i < nrem) {
                                if (P.store) crow[n + i] = v[4 * j + i];
                                else atomicAdd(crow + n + i, v[4 * j + i]);
                            }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<WG_BN>(tmem_d);
}

// zero the outputs before the split-K partials land: one CTA per 32 rows x 256 columns
__global__ void __launch_bounds__(256) wgrad_zero_kernel(const __grid_constant__ WgBatch wb) {
    const int bt = blockIdx.x >> 2, sub = blockIdx.x & 3;
    int pi = 0, start = 0;
    while (pi + 1 < wb.n && bt >= start + wb.p[pi].tiles_m * wb.p[pi].tiles_n) { start += wb.p[pi].tiles_m * wb.p[pi].tiles_n; ++pi; }
    const WgProblem& P = wb.p[pi];
    if (P.store) return;
    const int t = bt - start;
    const int m0 = (t / P.tiles_n) * WG_BM + sub * 32, n0 = (t % P.tiles_n) * WG_BN;
    const int cq = threadIdx.x & 63, r0 = threadIdx.x >> 6;       // 64 float4 columns x 4 rows per pass
    const int n = n0 + cq * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + r0 + i * 4;
        if (m >= P.M || m >= m0 + 32 || n >= P.N) continue;
        float* c = P.C + (long long)m * P.ldc + n;
        if (P.c_vec && n + 4 <= P.N) *reinterpret_cast<float4*>(c) = make_float4(0.f, 0.f, 0.f, 0.f);
        else
            for (int j = 0; j < 4 && n + j < P.N; ++j) c[j] = 0.f;
    }
}

// ---- fp32 -> bf16 hi/lo planes (elementwise; 8 floats per thread: 2 LDG.128 -> 2 STG.128)
struct SplitBatch {
    fhvae_split_problem p[FHVAE_SPLIT_MAX_BATCH];
    long long start[FHVAE_SPLIT_MAX_BATCH + 1];              // first 8-float item of every problem
    int n;
};
__global__ void __launch_bounds__(256) split_planes_kernel(const __grid_constant__ SplitBatch sb) {
    const long long total = sb.start[sb.n];
    for (long long item = (long long)blockIdx.x * 256 + threadIdx.x; item < total; item += (long long)gridDim.x * 256) {
        int pi = 0;
        while (pi + 1 < sb.n && item >= sb.start[pi + 1]) ++pi;
        const fhvae_split_problem& P = sb.p[pi];
        const long long e = item - sb.start[pi];
        const int c8 = P.cols >> 3;
        const long long r = e / c8;
        const int c = (int)(e - r * c8) * 8;
        const float4 x = __ldg(reinterpret_cast<const float4*>(P.src + r * P.ld_src + c));
        const float4 y = __ldg(reinterpret_cast<const float4*>(P.src + r * P.ld_src + c) + 1);
        const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        uint4 hi, lo;
        split_bf16(v, hi, lo);
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(P.dst) + r * P.ld_dst + c;
        *reinterpret_cast<uint4*>(d) = hi;
        *reinterpret_cast<uint4*>(d + P.plane_stride) = lo;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// planes tensor [2][K][W] of bf16 (row stride ld, plane stride ps, in elements) -> 3-D map, box 64 x BK x 1, 128B swizzle
static int make_map(CUtensorMap* map, const void* base, int W, int K, long long ld, long long ps) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("wgrad_planes: cuTensorMapEncodeTiled is not available from the driver"); return FHVAE_ENOSUP; }
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)K, 2};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ps * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)WG_BK, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad_planes: cuTensorMapEncodeTiled failed (%d)", (int)r); return FHVAE_EINVAL; }
    return 0;
}

}  // namespace fhvae

using namespace fhvae;

extern "C" int fhvae_wgrad_planes_batch(const fhvae_wgrad_problem* problems, int n, int mode, void* stream) {
    FHVAE_CHECK_ARG(problems && n > 0 && n <= FHVAE_WGRAD_MAX_BATCH, "wgrad_planes: need 1..%d problems", FHVAE_WGRAD_MAX_BATCH);
    FHVAE_CHECK_ARG(mode == FHVAE_MODE_BF16X3 || mode == FHVAE_MODE_BF16, "wgrad_planes: mode must be BF16X3 or BF16");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tma_kernel<WG_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(wgrad_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * WG_STAGE + 1024 + 256);
        if (e != cudaSuccess) { set_error("wgrad_planes: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    alignas(64) WgBatch wb;
    memset(&wb, 0, sizeof(wb));
    wb.passes = mode == FHVAE_MODE_BF16X3 ? 3 : 1;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const fhvae_wgrad_problem& p = problems[i];
        FHVAE_CHECK_ARG(p.A && p.B && p.C && p.M > 0 && p.N > 0 && p.K > 0, "wgrad_planes: problem %d: bad pointer or size", i);
        FHVAE_CHECK_ARG(p.lda % 8 == 0 && p.ldb % 8 == 0 && p.a_plane_stride % 8 == 0 && p.b_plane_stride % 8 == 0 &&
                        ((uintptr_t)p.A & 15) == 0 && ((uintptr_t)p.B & 15) == 0,
                        "wgrad_planes: problem %d: planes must be 16-byte aligned with strides multiple of 8 elements", i);
        tiles += cdiv(p.M, WG_BM) * cdiv(p.N, WG_BN);
    }
    // split K so that the launch is one wave of 1-CTA/SM tiles; at least 4 K-blocks per split.  One split count for all
    // problems (148 / 32 tiles = 4 for a stack's four gradients = 128 CTAs).  Measured alternatives, whole step at config 1:
    // splits capped at 3 / 2 / 1: +16 / +27 / +58 us; a cost-balanced split filling all 148 SMs (5,5,5,3): +6 us for the
    // step's last launch only, +15 us everywhere (more CTAs = more prologues and split-K partials per flop).
    const int want = tiles > 0 ? (kNumSM / tiles > 0 ? kNumSM / tiles : 1) : 1;
    int total = 0, ztotal = 0, nzero = 0, max_nit = 0;
    for (int i = 0; i < n; ++i) {
        const fhvae_wgrad_problem& p = problems[i];
        WgProblem& q = wb.p[wb.n];
        int r = make_map(&q.ta, p.A, p.M, p.K, p.lda, p.a_plane_stride);
        if (r) return r;
        r = make_map(&q.tb, p.B, p.N, p.K, p.ldb, p.b_plane_stride);
        if (r) return r;
        q.C = p.C; q.ldc = p.ldc; q.M = p.M; q.N = p.N; q.K = p.K;
        q.c_vec = (p.ldc % 4 == 0 && ((uintptr_t)p.C & 15) == 0);
        q.tiles_m = cdiv(p.M, WG_BM); q.tiles_n = cdiv(p.N, WG_BN);
        const int nkb = cdiv(p.K, WG_BK);
        int ks = want;
        if (ks > nkb / 4) ks = nkb / 4 > 0 ? nkb / 4 : 1;
        // deterministic mode: at most TWO partials meet in the pre-zeroed C -- (0 + a) + b == (0 + b) + a exactly,
        // so the arrival order of the red.global.add no longer matters
        if (deterministic_mode() && ks > 2) ks = 2;
        q.kb_per_split = cdiv(nkb, ks);
        q.ksplit = cdiv(nkb, q.kb_per_split);
        q.store = q.ksplit == 1;
        if (q.kb_per_split > max_nit) max_nit = q.kb_per_split;
        q.tile_start = total;
        total += q.tiles_m * q.tiles_n * q.ksplit;
        ztotal += q.tiles_m * q.tiles_n;
        nzero += q.store ? 0 : 1;
        ++wb.n;
    }
    cudaStream_t st = as_stream(stream);
    if (nzero > 0) {
        wgrad_zero_kernel<<<ztotal * 4, 256, 0, st>>>(wb);
        FHVAE_LAUNCH_CHECK("wgrad_zero");
    }
    wb.ns = max_nit <= 4 ? 2 : WG_NS;
    if (wb.ns == 2) wgrad_tma_kernel<2><<<total, WG_THREADS, 2 * WG_STAGE + 1024 + 256, st>>>(wb);
    else wgrad_tma_kernel<WG_NS><<<total, WG_THREADS, WG_SMEM, st>>>(wb);
    FHVAE_LAUNCH_CHECK("wgrad_tma");
    return 0;
}

extern "C" int fhvae_split_planes_batch(const fhvae_split_problem* problems, int n, void* stream) {
    FHVAE_CHECK_ARG(problems && n > 0 && n <= FHVAE_SPLIT_MAX_BATCH, "split_planes: need 1..%d problems", FHVAE_SPLIT_MAX_BATCH);
    SplitBatch sb;
    long long total = 0;
    for (int i = 0; i < n; ++i) {
        const fhvae_split_problem& p = problems[i];
        FHVAE_CHECK_ARG(p.src && p.dst && p.rows > 0 && p.cols > 0 && p.cols % 8 == 0 && p.ld_src % 4 == 0 &&
                        p.ld_dst % 8 == 0 && p.plane_stride % 8 == 0 && ((uintptr_t)p.src & 15) == 0 && ((uintptr_t)p.dst & 15) == 0,
                        "split_planes: problem %d: cols %% 8, 16-byte aligned rows and planes required", i);
        sb.p[i] = p;
        sb.start[i] = total;
        total += (long long)p.rows * (p.cols >> 3);
    }
    sb.start[n] = total;
    sb.n = n;
    const int grid = (int)((total + 255) / 256 < 4 * kNumSM * 8 ? (total + 255) / 256 : 4 * kNumSM * 8);
    split_planes_kernel<<<grid, 256, 0, as_stream(stream)>>>(sb);
    FHVAE_LAUNCH_CHECK("split_planes");
    return 0;
}
