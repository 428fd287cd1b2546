// Grouped GEMM on the 5th-gen tensor cores (tcgen05.mma kind::f16, accumulators in TMEM).
//
//   C[m,n] = sum_k A(m,k) B(k,n) (+bias, +beta*C, ReLU)      fp32 in HBM on both sides.
//
// Operands stay fp32 in HBM (they are master weights / saved activations that the fp32-parity
// contract needs); each CTA converts its tiles to bf16 on the way into shared memory:
//   FHVAE_MODE_BF16    one pass      hi(A) * hi(B)
//   FHVAE_MODE_BF16X3  three passes  lo*hi + hi*lo + hi*hi with x = hi + lo (+2^-18 x), fp32 accumulate
//                      => ~1e-5 relative, the "fp32" tolerance of the north star, at 1/3 of the bf16 rate.
// Shared-memory tiles use the canonical K-major no-swizzle UMMA layout (8x16B core matrices): element
// (r,k) of a 128 x 64 tile at byte (k/8)*2048 + r*16 + (k%8)*2, i.e. LBO = 2048, SBO = 128.
// 128x128 output tile per CTA, K stepped in blocks of BK (64 bf16 / 32 bf16x3: 64 KB of smem either way,
// two CTAs per SM) through a 2-stage ring: all 256 threads stage (global fp32 -> registers, prefetched one
// block ahead -> bf16 hi/lo -> st.shared, fence.proxy.async), one thread issues the MMAs, tcgen05.commit
// on an mbarrier releases the stage.  Long-K problems (weight gradients: K = T*B) are split along K
// across CTAs and reduced with fp32 atomics into a pre-zeroed C.
// Epilogue: tcgen05.ld 32x32b.x32 -> registers -> bias/beta/ReLU -> global.
#include "common.cuh"
#include "tc_common.cuh"

namespace fhvae {

using namespace tc;

constexpr int BM = 128, BN = 128;
constexpr int TCT = 256;                        // threads per CTA
constexpr uint32_t LBO = BM * 16, SBO = 128;
template <bool X3>
struct Cfg {
    static constexpr int BK = X3 ? 32 : 64;
    static constexpr int TILE_BYTES = BM * BK * 2;          // one bf16 operand tile
    static constexpr int OPS = X3 ? 4 : 2;                  // A_hi [A_lo] B_hi [B_lo]
    static constexpr int STAGE_BYTES = OPS * TILE_BYTES;    // 32 KB
    static constexpr int IT = BM * (BK / 8) / TCT;          // (row, k-chunk) items per thread per operand
    static constexpr int SMEM = 2 * STAGE_BYTES + 64;
};

struct TcProblem {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    int M, N, K;
    int relu;
    long long a_rs, a_ks;       // A(m,k) = A[m*a_rs + k*a_ks]
    long long b_rs, b_ks;       // B(k,n) = B[n*b_rs + k*b_ks]   (row = n)
    long long ldc;
    float beta;
    int a_vec, b_vec, c_vec;    // 16-byte vector access allowed
    int ksplit, kb_per_split;   // K blocks (of BK) per split
    int tile_start, tiles_n, tiles_mn;
};
struct TcBatch {
    TcProblem p[FHVAE_GEMM_MAX_BATCH];
    int n;
};

// One operand tile = 128 rows x BK: thread -> items (row = item & 127, k-chunk = item >> 7).
// load_tile: global fp32 -> registers (issued one k-block ahead);  store_tile: registers -> bf16 smem.
template <int IT>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, long long rs, long long ks, int vec,
                                          int row0, int rows, int k0, int kend, float (&v)[IT][8]) {
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int item = threadIdx.x + i * TCT;
        const int r = item & (BM - 1), kc = item >> 7;
        const int gr = row0 + r, gk = k0 + kc * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        if (gr < rows && gk < kend) {
            if (ks == 1) {
                const float* p = src + (long long)gr * rs + gk;
                if (vec && gk + 8 <= kend) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
                    v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
                    v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (gk + j < kend) v[i][j] = __ldg(p + j);
                }
            } else {   // rows contiguous: coalesced across the warp for every k
                const float* p = src + (long long)gk * ks + gr;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (gk + j < kend) v[i][j] = __ldg(p + (long long)j * ks);
            }
        }
    }
}
template <bool X3, int IT>
__device__ __forceinline__ void store_tile(const float (&v)[IT][8], uint8_t* s_hi, uint8_t* s_lo) {
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int item = threadIdx.x + i * TCT;
        const int r = item & (BM - 1), kc = item >> 7;
        const uint32_t off = (uint32_t)(kc * BM + r) * 16;
        if (X3) {
            uint4 hi, lo;
            split_bf16(v[i], hi, lo);
            *reinterpret_cast<uint4*>(s_hi + off) = hi;
            *reinterpret_cast<uint4*>(s_lo + off) = lo;
        } else {
            *reinterpret_cast<uint4*>(s_hi + off) = make_uint4(pack_bf16(v[i][0], v[i][1]), pack_bf16(v[i][2], v[i][3]),
                                                               pack_bf16(v[i][4], v[i][5]), pack_bf16(v[i][6], v[i][7]));
        }
    }
}

template <bool X3>
__global__ void __launch_bounds__(TCT, 2) gemm_tc_kernel(const __grid_constant__ TcBatch tb) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using C = Cfg<X3>;
    constexpr int BK = C::BK, TILE_BYTES = C::TILE_BYTES, STAGE_BYTES = C::STAGE_BYTES, IT = C::IT;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 2 * STAGE_BYTES);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

    int pi = 0;
    while (pi + 1 < tb.n && (int)blockIdx.x >= tb.p[pi + 1].tile_start) ++pi;
    const TcProblem& P = tb.p[pi];
    int t = blockIdx.x - P.tile_start;
    const int split = t / P.tiles_mn;
    t -= split * P.tiles_mn;
    const int m0 = (t / P.tiles_n) * BM, n0 = (t % P.tiles_n) * BN;
    const int nkb_total = (P.K + Cfg<X3>::BK - 1) / Cfg<X3>::BK;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(nkb_total, kb0 + P.kb_per_split);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc<BN>(tmem_slot);
    if (tid == 32) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);

    float va[IT][8], vb[IT][8];
    if (kb0 < kb1) {
        load_tile<IT>(P.A, P.a_rs, P.a_ks, P.a_vec, m0, P.M, kb0 * BK, P.K, va);
        load_tile<IT>(P.B, P.b_rs, P.b_ks, P.b_vec, n0, P.N, kb0 * BK, P.K, vb);
    }
    for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, s = it & 1;
        uint8_t* st = smem + s * STAGE_BYTES;
        if (it >= 2) mbar_wait(&mbar[s], ((it >> 1) - 1) & 1);     // MMAs that read this stage are done
        const int k0 = kb * BK;
        store_tile<X3, IT>(va, st, st + TILE_BYTES);
        store_tile<X3, IT>(vb, st + (X3 ? 2 : 1) * TILE_BYTES, st + 3 * TILE_BYTES);
        if (kb + 1 < kb1) {                                        // next block's loads fly during sync + MMA
            load_tile<IT>(P.A, P.a_rs, P.a_ks, P.a_vec, m0, P.M, k0 + BK, P.K, va);
            load_tile<IT>(P.B, P.b_rs, P.b_ks, P.b_vec, n0, P.N, k0 + BK, P.K, vb);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TILE_BYTES;
            const uint32_t b_hi = a_hi + (X3 ? 2 : 1) * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
            const int kleft = min(BK, P.K - k0);
            const int n16 = (kleft + 15) >> 4;
            for (int j = 0; j < n16; ++j) {
                const uint32_t ko = (uint32_t)j * 2 * LBO;
                const uint64_t dah = make_smem_desc(a_hi + ko, LBO, SBO), dbh = make_smem_desc(b_hi + ko, LBO, SBO);
                const uint32_t acc0 = (it > 0 || j > 0) ? 1u : 0u;
                if (X3) {
                    const uint64_t dal = make_smem_desc(a_lo + ko, LBO, SBO), dbl = make_smem_desc(b_lo + ko, LBO, SBO);
                    umma_bf16(tmem_d, dal, dbh, idesc, acc0);
                    umma_bf16(tmem_d, dah, dbl, idesc, 1u);
                    umma_bf16(tmem_d, dah, dbh, idesc, 1u);
                } else {
                    umma_bf16(tmem_d, dah, dbh, idesc, acc0);
                }
            }
            umma_commit(&mbar[s]);
        }
    }
    // drain: the last commit covers every MMA issued before it
    const int nit = kb1 - kb0;
    if (nit > 0) {
        const int last = nit - 1;
        mbar_wait(&mbar[last & 1], (last >> 1) & 1);
    }
    tc_fence_after();

    // epilogue: warp w reads TMEM lanes 32*(w&3).., columns 64*(w>>2)..+63 (thread = row), transposes each
    // 32x32 block through shared memory (the operand stages are free now) and writes full 128-byte lines
    const int q = warp & 3, half = warp >> 2;
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const bool atomic = P.ksplit > 1;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
        const int col0 = half * 64 + cc * 32;
        if (n0 + col0 >= P.N) break;
        float v[32];
        if (nit > 0) {
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = v[j];
        __syncwarp();
        const int gn = n0 + col0 + lane;                      // lane = column now
        const bool nok = gn < P.N;
        const float bv = (nok && P.bias && (!atomic || split == 0)) ? __ldg(P.bias + gn) : 0.f;
        const int mrow0 = m0 + q * 32;
        const int rmax = min(32, P.M - mrow0);
        if (nok) {
            float* cp = P.C + (long long)mrow0 * P.ldc + gn;
            if (atomic) {
                for (int r = 0; r < rmax; ++r) atomicAdd(cp + (long long)r * P.ldc, tr[r * 33 + lane] + bv);
            } else {
                for (int r = 0; r < rmax; ++r) {
                    float x = tr[r * 33 + lane] + bv;
                    if (P.beta != 0.f) x += P.beta * cp[(long long)r * P.ldc];
                    if (P.relu) x = fmaxf(x, 0.f);
                    cp[(long long)r * P.ldc] = x;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<BN>(tmem_d);
}

// zero the C tiles of split-K problems (beta == 0) before the atomics land
__global__ void __launch_bounds__(256) gemm_tc_zero_kernel(const __grid_constant__ TcBatch tb) {
    int pi = 0;
    while (pi + 1 < tb.n && (int)blockIdx.x >= tb.p[pi + 1].tile_start) ++pi;
    const TcProblem& P = tb.p[pi];
    const int t = blockIdx.x - P.tile_start;
    const int m0 = (t / P.tiles_n) * BM, n0 = (t % P.tiles_n) * BN;
    for (int e = threadIdx.x; e < BM * BN; e += 256) {
        const int m = m0 + e / BN, n = n0 + e % BN;
        if (m < P.M && n < P.N) P.C[(long long)m * P.ldc + n] = 0.f;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int gemm_batch_simt(const fhvae_gemm_problem* problems, int n, cudaStream_t st);

int gemm_batch_tc(const fhvae_gemm_problem* problems, int n, int mode, cudaStream_t st) {
    static bool attr_set = false;
    const bool x3 = (mode == FHVAE_MODE_BF16X3);
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM);
        cudaError_t e2 = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return (int)(e1 != cudaSuccess ? e1 : e2);
        }
        attr_set = true;
    }
    const int bk = x3 ? Cfg<true>::BK : Cfg<false>::BK;
    TcBatch tb, zb;
    memset(&tb, 0, sizeof(tb));
    memset(&zb, 0, sizeof(zb));
    fhvae_gemm_problem small[FHVAE_GEMM_MAX_BATCH];
    int nsmall = 0, total = 0, ztotal = 0, unsplit_tiles = 0;
    for (int i = 0; i < n; ++i)
        if (problems[i].M > 0 && problems[i].N > 0 && problems[i].K >= 16)
            unsplit_tiles += cdiv(problems[i].M, BM) * cdiv(problems[i].N, BN);
    // split K so that the launch offers ~2 CTAs per SM, never below 128 of K per split
    const int want_split = unsplit_tiles > 0 ? cdiv(2 * kNumSM, unsplit_tiles) : 1;
    for (int i = 0; i < n; ++i) {
        const fhvae_gemm_problem& p = problems[i];
        if (p.M == 0 || p.N == 0) continue;
        // tiny contractions (K < 16) gain nothing from the tensor pipe: exact fp32 path
        if (p.K < 16) { small[nsmall++] = p; continue; }
        TcProblem& q = tb.p[tb.n];
        q.A = p.A; q.B = p.B; q.C = p.C; q.bias = p.bias;
        q.M = p.M; q.N = p.N; q.K = p.K; q.relu = p.relu; q.beta = p.beta; q.ldc = p.ldc;
        q.a_rs = p.sa_m; q.a_ks = p.sa_k; q.b_rs = p.sb_n; q.b_ks = p.sb_k;
        q.a_vec = (p.sa_k == 1 && p.sa_m % 4 == 0 && aligned16(p.A));
        q.b_vec = (p.sb_k == 1 && p.sb_n % 4 == 0 && aligned16(p.B));
        q.c_vec = (p.ldc % 4 == 0 && aligned16(p.C) && (p.bias == nullptr || aligned16(p.bias)));
        const int nkb = cdiv(p.K, bk);
        int ks = 1;
        if (!p.relu && (p.beta == 0.f || p.beta == 1.f)) {
            const int max_split = p.K / 128;                 // >= 128 of K per split
            ks = want_split < max_split ? want_split : max_split;
            if (ks > 32) ks = 32;
            if (ks < 1) ks = 1;
        }
        q.kb_per_split = cdiv(nkb, ks);
        q.ksplit = cdiv(nkb, q.kb_per_split);
        q.tiles_n = cdiv(p.N, BN);
        q.tiles_mn = cdiv(p.M, BM) * q.tiles_n;
        q.tile_start = total;
        total += q.tiles_mn * q.ksplit;
        if (q.ksplit > 1 && p.beta == 0.f) {
            TcProblem& z = zb.p[zb.n];
            z = q;
            z.tile_start = ztotal;
            ztotal += q.tiles_mn;
            ++zb.n;
        }
        ++tb.n;
    }
    if (ztotal > 0) {
        gemm_tc_zero_kernel<<<ztotal, 256, 0, st>>>(zb);
        FHVAE_LAUNCH_CHECK("gemm_tc_zero");
    }
    if (total > 0) {
        if (x3) gemm_tc_kernel<true><<<total, TCT, Cfg<true>::SMEM, st>>>(tb);
        else    gemm_tc_kernel<false><<<total, TCT, Cfg<false>::SMEM, st>>>(tb);
        FHVAE_LAUNCH_CHECK("gemm_tc");
    }
    if (nsmall > 0) return gemm_batch_simt(small, nsmall, st);
    return 0;
}

}  // namespace fhvae
