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
// 128x128 output tile per CTA, K stepped in blocks of BK (64 bf16 / 32 bf16x3 = 32 KB per stage) through a
// 3-stage mbarrier ring (96 KB, two CTAs per SM), warp-specialised: 8 producer warps stage the operands
// (global fp32 -> a register ring that keeps DEPTH K-blocks of loads in flight per thread -> bf16 hi/lo ->
// st.shared, fence.proxy.async, one arrive per warp on full[s]); a 9th warp's elected lane waits full[s],
// issues the MMAs and tcgen05.commit's onto empty[s]; no CTA-wide barrier inside the K loop.  The producers
// then run the epilogue.  (Measured: with one K-block in flight per thread a stage took 2500-5000 cycles =
// one exposed L2/HBM round trip; tools/gemm_timeline.py.)  Long-K problems (weight gradients: K = T*B) are split along K
// across CTAs and reduced with fp32 atomics into a pre-zeroed C.
// Epilogue: tcgen05.ld 32x32b.x32 -> registers -> bias/beta/ReLU -> global.
#include "common.cuh"
#include "tc_common.cuh"

namespace fhvae {

using namespace tc;

constexpr int BM = 128, BN = 128;
constexpr int TCT = 256;                        // producer / epilogue threads per CTA
constexpr int TCT_ALL = TCT + 32;               // + the MMA-issuing warp
constexpr int NSTAGE = 3;                       // 96 KB: two CTAs per SM (296 slots for the 320-tile projections)
constexpr uint32_t LBO = BM * 16 + 32, SBO = 128;   // K-chunk planes padded by 32 B: 2-way instead of 4-way
                                                    // bank conflicts for the producers' 16-byte stores
template <bool X3>
struct Cfg;

// ---- coalesced path for K-contiguous operands (16-byte aligned): a warp-wide LDG.128 reads whole 128-byte
// lines (8 or 16 lanes per row) instead of 32 different rows -- the lane<->row mapping capped the L1 tag
// stage at one 32-byte sector per cycle (~3.7 TB/s chip-wide, tools/gemm_timeline.py).  A lane pair then
// swaps one float4 each (4 shuffles) so that every thread owns whole 8-element K chunks for the 16-byte
// UMMA-layout stores.
template <int BK, int NV>
__device__ __forceinline__ void load_tile_c(const float* __restrict__ src, long long rs, int row0, int rows,
                                            int k0, int kend, float4 (&v)[NV]) {
    constexpr int LPR = BK / 4;                          // lanes per row
    constexpr int RPW = 32 / LPR;                        // rows per warp-instruction
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane % LPR, rg = lane / LPR;
    const int gk = k0 + j * 4;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int gr = row0 + warp * (NV * RPW) + i * RPW + rg;
        v[i] = (gr < rows && gk + 4 <= kend) ? __ldcg(reinterpret_cast<const float4*>(src + (long long)gr * rs + gk))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <bool X3, int BK, int NV>
__device__ __forceinline__ void store_tile_c(const float4 (&v)[NV], uint8_t* s_hi, uint8_t* s_lo) {
    constexpr int LPR = BK / 4, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane % LPR, rg = lane / LPR;
    const bool odd = j & 1;
#pragma unroll
    for (int i = 0; i < NV; i += 2) {
        // even lane keeps row i (gets the partner's upper half), odd lane keeps row i+1 (gets the lower half)
        const float4 send = odd ? v[i] : v[i + 1];
        float4 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
        recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
        const float4 lo4 = odd ? recv : v[i], hi4 = odd ? v[i + 1] : recv;
        const float c[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        const int r = warp * (NV * RPW) + (i + (odd ? 1 : 0)) * RPW + rg, kc = j >> 1;
        const uint32_t off = (uint32_t)kc * LBO + (uint32_t)r * 16;
        if (X3) {
            uint4 hi, lo;
            split_bf16(c, hi, lo);
            *reinterpret_cast<uint4*>(s_hi + off) = hi;
            *reinterpret_cast<uint4*>(s_lo + off) = lo;
        } else {
            *reinterpret_cast<uint4*>(s_hi + off) = make_uint4(pack_bf16(c[0], c[1]), pack_bf16(c[2], c[3]),
                                                               pack_bf16(c[4], c[5]), pack_bf16(c[6], c[7]));
        }
    }
}

// ---- MN-contiguous operands (weight gradients: A(m,k) = G[k*ld + m], B(k,n) = X[k*ld + n]; data gradients:
// B(k,n) = W[k*ld + n]).  The tile goes into the canonical *MN-major* no-swizzle UMMA layout: a 2 KB plane per
// group of 8 k, core matrix = 8 k x 8 mn (16 B per k), element (mn, k) at (k/8)*LBO + (mn/8)*128 + (k%8)*16 +
// (mn%8)*2 -- same LBO/SBO numbers as the K-major planes, only the instruction descriptor's major bit differs.
// A global 32-byte run of 8 consecutive mn is one 16-byte smem store per bf16 part: no transposition, and
// 2 LDG.128 per item instead of 8 strided LDG.32.  Lane = (k%8) + 8*(mn-chunk%4): a warp reads 8 rows x 128 B
// and writes 4 whole core matrices (conflict-free).
template <int BK, int NV>
__device__ __forceinline__ void load_tile_mn(const float* __restrict__ src, long long ks, bool vec, int row0, int rows,
                                             int k0, int kend, float4 (&v)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NI = NV / 2;                              // items (8 mn x 1 k) per thread
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int idx = warp * NI + i, kg = idx >> 2, mg = idx & 3;
        const int k = k0 + kg * 8 + (lane & 7), m = row0 + (mg * 4 + (lane >> 3)) * 8;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
        if (k < kend && m < rows) {
            const float* p = src + (long long)k * ks + m;
            if (vec && m + 8 <= rows) {
                x = __ldcg(reinterpret_cast<const float4*>(p));
                y = __ldcg(reinterpret_cast<const float4*>(p) + 1);
            } else {
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = (m + j < rows) ? __ldcg(p + j) : 0.f;
                x = make_float4(t[0], t[1], t[2], t[3]);
                y = make_float4(t[4], t[5], t[6], t[7]);
            }
        }
        v[2 * i] = x;
        v[2 * i + 1] = y;
    }
}
template <bool X3, int BK, int NV>
__device__ __forceinline__ void store_tile_mn(const float4 (&v)[NV], uint8_t* s_hi, uint8_t* s_lo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NI = NV / 2;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int idx = warp * NI + i, kg = idx >> 2, mg = idx & 3;
        const float c[8] = {v[2 * i].x, v[2 * i].y, v[2 * i].z, v[2 * i].w, v[2 * i + 1].x, v[2 * i + 1].y, v[2 * i + 1].z, v[2 * i + 1].w};
        const uint32_t off = (uint32_t)kg * LBO + (uint32_t)(mg * 4 + (lane >> 3)) * 128 + (uint32_t)(lane & 7) * 16;
        if (X3) {
            uint4 hi, lo;
            split_bf16(c, hi, lo);
            *reinterpret_cast<uint4*>(s_hi + off) = hi;
            *reinterpret_cast<uint4*>(s_lo + off) = lo;
        } else {
            *reinterpret_cast<uint4*>(s_hi + off) = make_uint4(pack_bf16(c[0], c[1]), pack_bf16(c[2], c[3]),
                                                               pack_bf16(c[4], c[5]), pack_bf16(c[6], c[7]));
        }
    }
}

template <bool X3>
struct Cfg {
    static constexpr int BK = X3 ? 32 : 64;
    static constexpr int TILE_BYTES = (BK / 8) * (int)LBO;  // one bf16 operand tile (K-chunk planes of LBO bytes)
    static constexpr int OPS = X3 ? 4 : 2;                  // A_hi [A_lo] B_hi [B_lo]
    static constexpr int STAGE_BYTES = OPS * TILE_BYTES;    // 32 KB
    static constexpr int IT = BM * (BK / 8) / TCT;          // (row, k-chunk) items per thread per operand
    static constexpr int NV = 2 * IT;                       // float4 loads per thread per operand
    static constexpr int DEPTH = X3 ? 2 : 1;                // K-blocks of global loads in flight per thread
    static constexpr int SMEM = NSTAGE * STAGE_BYTES + 128;
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
    int a_mn, b_mn;             // operand is M/N-contiguous: staged in the MN-major UMMA layout
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
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
                    const float4 b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
                    v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
                    v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (gk + j < kend) v[i][j] = __ldcg(p + j);
                }
            } else {   // rows contiguous: coalesced across the warp for every k
                const float* p = src + (long long)gk * ks + gr;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (gk + j < kend) v[i][j] = __ldcg(p + (long long)j * ks);
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
        const uint32_t off = (uint32_t)kc * LBO + (uint32_t)r * 16;
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

#ifdef FHVAE_TIMELINE
__device__ long long g_gemm_tl[64];
#define GTL(slot) do { if (blockIdx.x == 0 && (threadIdx.x == 0)) g_gemm_tl[slot] = clock64(); } while (0)
#define GTLM(slot) do { if (blockIdx.x == 0) g_gemm_tl[slot] = clock64(); } while (0)
extern "C" int fhvae_debug_gemm_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_gemm_tl, sizeof(g_gemm_tl)); }
#else
#define GTL(slot) do { } while (0)
#define GTLM(slot) do { } while (0)
#endif

template <bool X3>
__global__ void __launch_bounds__(TCT_ALL, 2) gemm_tc_kernel(const __grid_constant__ TcBatch tb) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using C = Cfg<X3>;
    constexpr int BK = C::BK, TILE_BYTES = C::TILE_BYTES, STAGE_BYTES = C::STAGE_BYTES, IT = C::IT;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);   // [NSTAGE] producers -> MMA
    uint64_t* empty = full + NSTAGE;                                             // [NSTAGE] MMA done -> producers
    uint64_t* accd = empty + NSTAGE;                                             // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accd + 1);

    int pi = 0;
    while (pi + 1 < tb.n && (int)blockIdx.x >= tb.p[pi + 1].tile_start) ++pi;
    const TcProblem& P = tb.p[pi];
    int t = blockIdx.x - P.tile_start;
    const int split = t / P.tiles_mn;
    t -= split * P.tiles_mn;
    const int m0 = (t / P.tiles_n) * BM, n0 = (t % P.tiles_n) * BN;
    const int nkb_total = (P.K + BK - 1) / BK;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(nkb_total, kb0 + P.kb_per_split);
    const int nit = kb1 - kb0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    GTL(0);
    pdl_launch_dependents();

    if (warp == TCT / 32) tmem_alloc<BN>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], TCT / 32); mbar_init(&empty[i], 1); }
        mbar_init(accd, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc_bf16(BM, BN) | ((uint32_t)P.a_mn << 15) | ((uint32_t)P.b_mn << 16);
    pdl_wait();                 // barrier / TMEM setup above overlaps the preceding launch's tail
    GTL(1);

    if (warp == TCT / 32) {
        // ================= MMA issuer: one elected lane =================
        if (elect_one()) {
            for (int it = 0; it < nit; ++it) {
                const int s = it % NSTAGE;
                mbar_wait(&full[s], (it / NSTAGE) & 1);
                if (it < 12) GTLM(32 + 2 * it);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + TILE_BYTES;
                const uint32_t b_hi = a_hi + (X3 ? 2 : 1) * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
                const int kleft = min(BK, P.K - (kb0 + it) * BK);
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
                umma_commit(&empty[s]);            // stage s reusable once these MMAs have read it
                if (it < 12) GTLM(33 + 2 * it);
            }
            umma_commit(accd);                     // covers every MMA issued above
        }
        __syncwarp();
    } else {
        // ================= producers (8 warps), then epilogue =================
        constexpr int DEPTH = C::DEPTH, NV = C::NV;
        // register ring: DEPTH K-blocks of loads in flight per thread.  A float4[NV] slot holds either NV
        // coalesced float4 loads (K-contiguous, aligned operand) or IT items of 8 floats (any other operand).
        float4 ra[DEPTH][NV], rb[DEPTH][NV];
        const bool ca = P.a_vec && !P.a_mn, cb = P.b_vec && !P.b_mn;   // CTA-uniform
        const bool ma = P.a_mn, mb = P.b_mn;
        auto LOAD_A = [&](int k0, float4 (&slot)[NV]) {
            if (ma) load_tile_mn<BK, NV>(P.A, P.a_ks, P.a_vec, m0, P.M, k0, P.K, slot);
            else if (ca) load_tile_c<BK, NV>(P.A, P.a_rs, m0, P.M, k0, P.K, slot);
            else load_tile<IT>(P.A, P.a_rs, P.a_ks, 0, m0, P.M, k0, P.K, reinterpret_cast<float (&)[IT][8]>(slot));
        };
        auto LOAD_B = [&](int k0, float4 (&slot)[NV]) {
            if (mb) load_tile_mn<BK, NV>(P.B, P.b_ks, P.b_vec, n0, P.N, k0, P.K, slot);
            else if (cb) load_tile_c<BK, NV>(P.B, P.b_rs, n0, P.N, k0, P.K, slot);
            else load_tile<IT>(P.B, P.b_rs, P.b_ks, 0, n0, P.N, k0, P.K, reinterpret_cast<float (&)[IT][8]>(slot));
        };
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
            if (d < nit) { LOAD_A((kb0 + d) * BK, ra[d]); LOAD_B((kb0 + d) * BK, rb[d]); }
        for (int it0 = 0; it0 < nit; it0 += DEPTH) {
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                const int it = it0 + d;
                if (it < nit) {
                    const int s = it % NSTAGE;
                    uint8_t* st = smem + s * STAGE_BYTES;
                    if (it >= NSTAGE) mbar_wait(&empty[s], ((it / NSTAGE) - 1) & 1);
                    uint8_t* sb = st + (X3 ? 2 : 1) * TILE_BYTES;
                    if (ma) store_tile_mn<X3, BK, NV>(ra[d], st, st + TILE_BYTES);
                    else if (ca) store_tile_c<X3, BK, NV>(ra[d], st, st + TILE_BYTES);
                    else store_tile<X3, IT>(reinterpret_cast<float (&)[IT][8]>(ra[d]), st, st + TILE_BYTES);
                    if (mb) store_tile_mn<X3, BK, NV>(rb[d], sb, st + 3 * TILE_BYTES);
                    else if (cb) store_tile_c<X3, BK, NV>(rb[d], sb, st + 3 * TILE_BYTES);
                    else store_tile<X3, IT>(reinterpret_cast<float (&)[IT][8]>(rb[d]), sb, st + 3 * TILE_BYTES);
                    if (it + DEPTH < nit) {        // refill this ring slot: DEPTH K-blocks stay in flight
                        const int k0 = (kb0 + it + DEPTH) * BK;
                        LOAD_A(k0, ra[d]);
                        LOAD_B(k0, rb[d]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    if (it < 12) GTL(2 + it);
                }
            }
        }
        if (nit > 0) mbar_wait(accd, 0);
        tc_fence_after();
        GTL(20);

        // epilogue: warp w owns the 32x32 block (TMEM lanes 32*(w&3).., columns 32*(w>>2)..): thread = row,
        // transposed through shared memory (the operand stages are free now) into full 128-byte line stores
#pragma unroll 1
        for (int cb = warp >> 2; cb < BN / 32; cb += TCT / 128) {
            const int q = warp & 3, col0 = cb * 32;
            float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
            const bool atomic = P.ksplit > 1;
            __syncwarp();
            if (n0 + col0 < P.N) {
                float v[32];
                if (nit > 0) {
                    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = v[j];
                __syncwarp();
                const int gn = n0 + col0 + lane;                  // lane = column now
                const bool nok = gn < P.N;
                const float bv = (nok && P.bias && (!atomic || split == 0)) ? __ldcg(P.bias + gn) : 0.f;
                const int mrow0 = m0 + q * 32;
                const int rmax = min(32, P.M - mrow0);
                if (nok) {
                    float* cp = P.C + (long long)mrow0 * P.ldc + gn;
                    if (atomic && P.c_vec && n0 + col0 + 32 <= P.N) {
                        // split-K partials: one red.global.add.v4.f32 per 4 columns (4x fewer L2 atomic operations
                        // than scalar atomics: the L2 atomic units were the tail of every weight-gradient GEMM)
                        const int cgp = lane & 7, rr = lane >> 3;
                        float* cq = P.C + (long long)mrow0 * P.ldc + n0 + col0 + cgp * 4;
                        const float4 b4 = (P.bias && split == 0) ? __ldcg(reinterpret_cast<const float4*>(P.bias + n0 + col0 + cgp * 4))
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rr + 4 * i;
                            if (r < rmax) {
                                const float* t4 = tr + r * 33 + cgp * 4;
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cq + (long long)r * P.ldc),
                                             "f"(t4[0] + b4.x), "f"(t4[1] + b4.y), "f"(t4[2] + b4.z), "f"(t4[3] + b4.w) : "memory");
                            }
                        }
                    } else if (atomic) {
#pragma unroll 8
                        for (int r = 0; r < rmax; ++r) atomicAdd(cp + (long long)r * P.ldc, tr[r * 33 + lane] + bv);
                    } else if (P.beta == 0.f) {
#pragma unroll 8
                        for (int r = 0; r < rmax; ++r) {
                            float x = tr[r * 33 + lane] + bv;
                            if (P.relu) x = fmaxf(x, 0.f);
                            cp[(long long)r * P.ldc] = x;
                        }
                    } else {
#pragma unroll 8
                        for (int r = 0; r < rmax; ++r) {
                            float x = tr[r * 33 + lane] + bv + P.beta * cp[(long long)r * P.ldc];
                            if (P.relu) x = fmaxf(x, 0.f);
                            cp[(long long)r * P.ldc] = x;
                        }
                    }
                }
            }
        }
    GTL(21);
    }   // producers / epilogue
    tc_fence_before();
    __syncthreads();
    if (warp == TCT / 32) tmem_dealloc<BN>(tmem_d);
}

// zero the C tiles of split-K problems (beta == 0) before the atomics land: 4 CTAs per 128x128 tile
__global__ void __launch_bounds__(256) gemm_tc_zero_kernel(const __grid_constant__ TcBatch tb) {
    const int bt = blockIdx.x >> 2, sub = blockIdx.x & 3;
    int pi = 0;
    while (pi + 1 < tb.n && bt >= tb.p[pi + 1].tile_start) ++pi;
    const TcProblem& P = tb.p[pi];
    const int t = bt - P.tile_start;
    const int m0 = (t / P.tiles_n) * BM + sub * 32, n0 = (t % P.tiles_n) * BN;
    const int cq = threadIdx.x & 31, r0 = threadIdx.x >> 5;           // 32 float4 columns x 8 rows per pass
    const int n = n0 + cq * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + r0 + i * 8;
        if (m >= P.M || n >= P.N) continue;
        float* c = P.C + (long long)m * P.ldc + n;
        if (P.c_vec && n + 4 <= P.N) *reinterpret_cast<float4*>(c) = make_float4(0.f, 0.f, 0.f, 0.f);
        else
            for (int j = 0; j < 4 && n + j < P.N; ++j) c[j] = 0.f;
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
    // (floor, not ceil: 304 CTAs on 296 slots is two waves -- measured 45 us vs 33 us for the same weight gradient)
    const int want_split = unsplit_tiles > 0 ? ((2 * kNumSM) / unsplit_tiles > 0 ? (2 * kNumSM) / unsplit_tiles : 1) : 1;
    for (int i = 0; i < n; ++i) {
        const fhvae_gemm_problem& p = problems[i];
        if (p.M == 0 || p.N == 0) continue;
        // tiny contractions (K < 16) gain nothing from the tensor pipe: exact fp32 path
        if (p.K < 16) { small[nsmall++] = p; continue; }
        TcProblem& q = tb.p[tb.n];
        q.A = p.A; q.B = p.B; q.C = p.C; q.bias = p.bias;
        q.M = p.M; q.N = p.N; q.K = p.K; q.relu = p.relu; q.beta = p.beta; q.ldc = p.ldc;
        q.a_rs = p.sa_m; q.a_ks = p.sa_k; q.b_rs = p.sb_n; q.b_ks = p.sb_k;
        q.a_mn = (p.sa_m == 1 && p.sa_k != 1);
        q.b_mn = (p.sb_n == 1 && p.sb_k != 1);
        q.a_vec = q.a_mn ? (p.sa_k % 4 == 0 && aligned16(p.A)) : (p.sa_k == 1 && p.sa_m % 4 == 0 && p.K % 4 == 0 && aligned16(p.A));
        q.b_vec = q.b_mn ? (p.sb_k % 4 == 0 && aligned16(p.B)) : (p.sb_k == 1 && p.sb_n % 4 == 0 && p.K % 4 == 0 && aligned16(p.B));
        q.c_vec = (p.ldc % 4 == 0 && aligned16(p.C) && (p.bias == nullptr || aligned16(p.bias)));
        const int nkb = cdiv(p.K, bk);
        int ks = 1;
        if (!p.relu && (p.beta == 0.f || p.beta == 1.f)) {
            const int max_split = p.K / 128;                 // >= 128 of K per split (256: the decoder head unsplit, +2 us/step)
            ks = want_split < max_split ? want_split : max_split;
            if (ks > 32) ks = 32;
            if (ks < 1) ks = 1;
            // deterministic mode: <= 2 partials into a zeroed C commute exactly; accumulating launches (beta == 1) do not split
            if (deterministic_mode()) ks = p.beta == 0.f ? (ks > 2 ? 2 : ks) : 1;
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
        gemm_tc_zero_kernel<<<ztotal * 4, 256, 0, st>>>(zb);
        FHVAE_LAUNCH_CHECK("gemm_tc_zero");
    }
    if (total > 0) {
        if (x3) launch_pdl(PDL_GEMM, gemm_tc_kernel<true>, dim3(total), dim3(TCT_ALL), Cfg<true>::SMEM, st, tb);
        else    launch_pdl(PDL_GEMM, gemm_tc_kernel<false>, dim3(total), dim3(TCT_ALL), Cfg<false>::SMEM, st, tb);
        FHVAE_LAUNCH_CHECK("gemm_tc");
    }
    if (nsmall > 0) return gemm_batch_simt(small, nsmall, st);
    return 0;
}

}  // namespace fhvae
